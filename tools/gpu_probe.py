#!/usr/bin/env python3
"""Quick device-side timing probe (development aid; bench.py is the contract).  Prints IMAD / modmul
throughput, NTT and MSM timings with inputs resident in HBM."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import codec  # noqa: E402
import pcd_b200  # noqa: E402
from pcd_b200 import synthetic  # noqa: E402


def timed(ctx, fn, reps=3):
    fn()
    ctx.sync()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ctx.sync()
        ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3


def main():
    out = {}
    ctx = pcd_b200.Context(0)
    dev = torch.device("cuda:0")
    for mode, name in ((0, "imad_wide"), (3, "imad_wide_carry_chain"), (1, "modmul_r4"), (2, "modmul_q4")):
        ops, ms = ctx.bench_imad(mode, 4000 if mode == 0 else 2000)
        out[name] = {"ops_per_s": ops, "ms": ms}
        print(name, "%.3e ops/s  %.3f ms" % (ops, ms), flush=True)
    for log_n in [int(x) for x in os.environ.get("PROBE_NTT", "16,20,22,24").split(",") if x]:
        x = torch.from_numpy(codec.random_field_elems(1 << min(log_n, 20), 0, 5).view(np.int64)).to(dev)
        if log_n > 20:
            x = x.repeat(1 << (log_n - 20), 1)
        ms = timed(ctx, lambda: ctx.ntt_dev(0, x.data_ptr(), log_n, False, True))
        gbs = 2 * 40 * (1 << log_n) / (ms * 1e-3) / 1e9
        mm = (1 << (log_n - 1)) * log_n / (ms * 1e-3)
        out["ntt_2^%d" % log_n] = {"ms": ms, "GBps": gbs, "modmul_per_s": mm}
        print("ntt r4 2^%d: %.3f ms  %.1f GB/s  %.2e butterflies/s" % (log_n, ms, gbs, mm), flush=True)
    for curve, sizes in ((0, os.environ.get("PROBE_MSM", "16,18,20")), (1, os.environ.get("PROBE_MSM_G2", "16,18"))):
        for log_n in [int(x) for x in sizes.split(",") if x]:
            n = 1 << log_n
            pts = synthetic.random_points_dev(ctx, curve, n, seed=3)
            sc = torch.from_numpy(codec.random_field_elems(n, 0, 9).view(np.int64)).to(dev)
            res = torch.zeros(64, dtype=torch.int64, device=dev)
            for c in [int(v) for v in os.environ.get("PROBE_C", "0").split(",")]:
                ctx.set_msm_window(c)
                ms = timed(ctx, lambda: ctx.msm_dev(curve, pts.data_ptr(), sc.data_ptr(), n, res.data_ptr()))
                out["msm_c%d_2^%d_w%d" % (curve, log_n, c)] = {"ms": ms, "Mpts": n / ms / 1e3}
                print("msm curve %d 2^%d c=%d: %.3f ms  %.2f Mpts/s" % (curve, log_n, c, ms, n / ms / 1e3), flush=True)
            ctx.set_msm_window(0)
            if curve == 0:
                b = pcd_b200.Bases(ctx, curve, pts.cpu().numpy().view(np.uint64), precompute=True)
                ms = timed(ctx, lambda: b.msm_dev(sc.data_ptr(), n, res.data_ptr()))
                out["msm_pre_c%d_2^%d" % (curve, log_n)] = {"ms": ms, "Mpts": n / ms / 1e3}
                print("msm precomputed curve %d 2^%d: %.3f ms  %.2f Mpts/s" % (curve, log_n, ms, n / ms / 1e3), flush=True)
                b.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
