#!/usr/bin/env python3
"""bench.py -- the Groth16 proving step of arkworks-rs/pcd (IC::MainSNARK::prove of ECCyclePCD::prove,
/root/reference/src/ec_cycle_pcd/mod.rs:171) on B200, through libpcdgpu.so.

One "step" = one Groth16 proof on MNT4-298 for a synthetic satisfiable R1CS whose evaluation domain
is 2^LOG_N (default 2^20: 2^20 - 2 constraints, 2^20 variables): the CSR witness map with its seven
NTTs, four G1 MSMs and one G2 MSM over the resident proving key, and the proof assembly.  With N > 1
every GPU proves its own independent instance (independent PCD nodes: no data-path collective,
weak scaling).  The JSON line also carries the two kernel figures BASELINE.json names: G1 MSM at
2^20 points (Mpts/s) and the largest radix-2 NTT measured (GB/s), each with its roofline fraction.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--log-n 20] [--impl reference]

`--impl reference`: the CPU arm.  The reference (Rust, un-vendored arkworks crates) cannot be built
here, so this times oracle/c (the C++ restatement of the same algorithms in the shape arkworks runs
them: 5x64 CIOS, arkworks' Pippenger with one task per window, per-stage parallel FFT) on all host
threads, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODMUL_IMADS = 210            # 10-limb Montgomery product
MADD_MODMULS = {1: 10, 2: 28, 3: 58}  # XYZZ mixed add 8M + 2S in Fq / Fq2 (M=3,S=2) / Fq3 (M=6,S=5)
METRIC = "groth16_proofs_per_sec"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# CPU arm (oracle/c): bounded sample of the same workload
# ---------------------------------------------------------------------------------------------------
def cpu_sample_setup(log_n_sample, seed=5):
    """Instance of the same synthetic family at 2^log_n_sample for the CPU arm.  The key's points are
    4096 random group elements tiled over the queries: Pippenger's running time does not depend on
    which points it adds, and building 5 * 2^17 distinct points on the CPU would dominate the run."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import c_oracle as co
    from pcd_b200 import synthetic

    class NoGpu:  # the generator only needs fixed_base_mul for the key; give it the oracle's
        def fixed_base_mul(self, curve, g, k):
            k = np.ascontiguousarray(k, dtype=np.uint64).reshape(-1, 5)
            if k.shape[0] <= 4096:
                return co.fixed_base_mul(curve, g, k)
            base = co.fixed_base_mul(curve, g, k[:4096])
            reps = (k.shape[0] + 4095) // 4096
            return np.tile(base, (reps, 1))[:k.shape[0]].copy()

    inst = synthetic.make_groth16_instance(NoGpu(), 0, log_n_sample, seed=seed)
    return co, inst


def cpu_prove_once(co, inst, threads, r, s):
    t0 = time.perf_counter()
    co.groth16_prove(0, inst["pk"], inst["A"], inst["B"], inst["C"], inst["m"], inst["num_inputs"],
                     inst["num_witness"], inst["z"], r, s, threads=threads)
    return time.perf_counter() - t0


def run_reference(args, rank, world):
    if rank != 0:
        return
    log_n = args.log_n
    sample_log = min(log_n, args.cpu_sample_log_n)
    co, inst = cpu_sample_setup(sample_log)
    threads = co.hw_threads()
    rng = np.random.Generator(np.random.Philox(99))
    draw = lambda: np.concatenate([rng.integers(0, 2 ** 64, 4, dtype=np.uint64), np.zeros(1, np.uint64)])
    for _ in range(args.warmup):
        cpu_prove_once(co, inst, threads, draw(), draw())
    t = 0.0
    for _ in range(args.steps):
        t += cpu_prove_once(co, inst, threads, draw(), draw())
    frac = 2.0 ** (sample_log - log_n)
    value = frac * args.steps / t
    sample = ("one Groth16 proof (MNT4-298) at 2^%d constraints per step = 2^%d of the 2^%d workload; value scaled "
              "linearly by that fraction" % (sample_log, sample_log - log_n, log_n))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "proofs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps / frac,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64 (5x64-bit limbs)",
        "data": "synthetic",
        "config": {"workload": "groth16_mnt4_298_domain_2^%d" % log_n, "cpu_impl": "oracle/c (C++ restatement of "
                   "arkworks' prover; the Rust reference cannot be built here)"},
        "cpu_baseline": {"value": value, "unit": "proofs/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
def run_gpu(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    import pcd_b200
    from pcd_b200 import synthetic

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    ctx = pcd_b200.Context(local_rank)
    stream = torch.cuda.Stream(device=dev)  # a real (non-default) stream shared by torch's events and the library
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    log = (lambda m: print("[bench rank %d] %s" % (rank, m), file=sys.stderr, flush=True)) if rank == 0 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # the integer roof: independent accumulate-form IMAD.WIDE.U32 (fmaheavy pipe, 32 lanes/clk/SM on B200 -- the
    # first version of this microbenchmark let ptxas hoist the products and measured 64-bit ADDS instead)
    imad_indep, _ = ctx.bench_imad(0, 4000)
    imad_chain, _ = ctx.bench_imad(3, 4000)  # the same multiply-adds as carry chains (.X form): same pipe, same rate
    imad_peak = max(imad_indep, imad_chain)

    # ---- workload -----------------------------------------------------------------------------------
    log_n = args.log_n
    inst = synthetic.make_groth16_instance(ctx, pcd_b200.MNT4_298, log_n, seed=20261017 + 1000 * rank, verbose=log)
    g = pcd_b200.Groth16(ctx, pcd_b200.MNT4_298)
    pk = pcd_b200.ProvingKey(pairing=0, **inst["pk"])
    cm = pcd_b200.ConstraintMatrices(0, inst["num_inputs"], inst["num_witness"], inst["A"], inst["B"], inst["C"])
    if os.environ.get("PCD_MSM_WINDOW"):  # development aid: override the window size of the resident tables
        ctx.set_msm_window(int(os.environ["PCD_MSM_WINDOW"]))
    idx = g.index(pk, cm, precompute=not args.no_precompute)
    ctx.set_msm_window(0)
    ctx.sync()
    if log:
        log("key resident on the GPU (precompute=%s)" % (not args.no_precompute))
    nvars = inst["num_inputs"] + inst["num_witness"]
    z_host = torch.from_numpy(inst["z"].view(np.int64)).pin_memory()
    z_dev = z_host.to(dev)
    p = inst["p"]
    rng = np.random.Generator(np.random.Philox(7 + rank))

    def draw():
        v = int.from_bytes(rng.bytes(40), "little") % p
        return v, np.array([(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(5)], dtype=np.uint64)

    # correctness gate, outside the timed region: the proof must equal [known discrete logs] * G
    r_i, r_l = draw()
    s_i, s_l = draw()
    proof = g.create_proof_dev(idx, z_dev.data_ptr(), r_l, s_l)
    expect = synthetic.expected_proof(ctx, inst, r_i, s_i)
    if not np.array_equal(proof.affine_limbs(), expect):
        raise SystemExit("bench: GPU proof does not match its known discrete logarithms -- refusing to time it")
    if log:
        log("proof at 2^%d verified against the trapdoor" % log_n)

    rs = [(draw()[1], draw()[1]) for _ in range(args.warmup + 2 * args.steps)]
    # Proofs in flight: independent proofs (PCD nodes of one tree depth) are issued from `inflight` host
    # threads, each with its own context (streams + scratch) over the SAME resident key and matrices, so
    # one proof's single-warp tail (window combination, proof assembly) overlaps the next one's MSMs.
    nfl = max(1, args.inflight)
    lanes = [(ctx, g, stream)]
    for _ in range(nfl - 1):
        c2 = pcd_b200.Context(local_rank)
        s2 = torch.cuda.Stream(device=dev)
        c2.set_stream(s2.cuda_stream)
        lanes.append((c2, pcd_b200.Groth16(c2, pcd_b200.MNT4_298), s2))

    def run_steps(first, count, host):
        """prove rs[first : first + count], round-robin over the in-flight contexts; returns elapsed ms
        measured with CUDA events on the launching streams (earliest start to latest end)"""
        starts = [torch.cuda.Event(enable_timing=True) for _ in lanes]
        ends = [torch.cuda.Event(enable_timing=True) for _ in lanes]
        errs = []

        def worker(k):
            try:
                torch.cuda.set_device(local_rank)
                c_, g_, s_ = lanes[k]
                starts[k].record(s_)
                for i in range(first + k, first + count, len(lanes)):
                    r_l, s_l = rs[i]
                    if host:
                        out = np.zeros(40, dtype=np.uint64)
                        c_._check(c_.lib.pcdgpu_groth16_prove(c_.h, idx.pk, idx.r1cs, z_host.data_ptr(), r_l.ctypes.data,
                                                              s_l.ctypes.data, out.ctypes.data))
                    else:
                        g_.create_proof_dev(idx, z_dev.data_ptr(), r_l, s_l)
                ends[k].record(s_)
            except Exception as e:  # noqa: BLE001
                errs.append(e)

        ths = [threading.Thread(target=worker, args=(k,)) for k in range(len(lanes))]
        t0 = time.perf_counter()
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        torch.cuda.synchronize()
        wall = 1e3 * (time.perf_counter() - t0)
        if errs:
            raise errs[0]
        dev_ms = max(a.elapsed_time(b) for a in starts for b in ends)
        return dev_ms, wall

    if args.no_concurrency:
        for c_, _, _ in lanes:
            c_.set_concurrency(False)
    for k in range(len(lanes)):
        for i in range(args.warmup):
            lanes[k][1].create_proof_dev(idx, z_dev.data_ptr(), *rs[i])

    # ---- timed region 1: inputs resident in HBM -------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    import ctypes
    NC = 8
    ms = (ctypes.c_double * NC)()
    units = (ctypes.c_double * NC)()
    spans = (ctypes.c_uint64 * NC)()
    launches = ctypes.c_uint64()
    ctx.lib.pcdgpu_profile_enable(ctx.h, 0)  # no event spans in the timed region; resets the launch counter
    for c_, _, _ in lanes[1:]:
        c_.lib.pcdgpu_profile_enable(c_.h, 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    dev_ms, _ = run_steps(args.warmup, args.steps, host=False)
    barrier()
    ms_dev = max_over_ranks(dev_ms)
    ctx._check(ctx.lib.pcdgpu_profile_read(ctx.h, ms, units, spans, ctypes.byref(launches)))
    n_launches = int(launches.value)
    for c_, _, _ in lanes[1:]:
        l2 = ctypes.c_uint64()
        c_._check(c_.lib.pcdgpu_profile_read(c_.h, ms, units, spans, ctypes.byref(l2)))
        n_launches += int(l2.value)
    # per-kernel pass (same proofs again): the MSMs of a proof normally overlap on five streams, which
    # makes per-kernel durations meaningless, so this pass serialises them and records CUDA-event spans
    ctx.set_concurrency(False)
    g.create_proof_dev(idx, z_dev.data_ptr(), *rs[0])  # grows lane 0's scratch outside the spans
    ctx.lib.pcdgpu_profile_enable(ctx.h, 1)
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record(stream)
    for i in range(args.steps):
        g.create_proof_dev(idx, z_dev.data_ptr(), *rs[args.warmup + i])
    e3.record(stream)
    torch.cuda.synchronize()
    ms_serial = e2.elapsed_time(e3)
    launches2 = ctypes.c_uint64()
    ctx._check(ctx.lib.pcdgpu_profile_read(ctx.h, ms, units, spans, ctypes.byref(launches2)))
    ctx.lib.pcdgpu_profile_enable(ctx.h, 0)
    ctx.set_concurrency(not args.no_concurrency)
    prof = {"ms": list(ms), "units": list(units), "spans": list(spans)}

    # ---- timed region 2: end to end through the C ABI with host buffers ---------------------------------
    run_steps(0, len(lanes), host=True)  # untimed: grows the host path's staging scratch on every in-flight context
    barrier()
    dev_ms, wall_ms = run_steps(args.warmup + args.steps, args.steps, host=True)
    barrier()
    ms_e2e = max_over_ranks(max(dev_ms, wall_ms))
    clocks = sampler.stop()

    # ---- kernel figures: G1 MSM at 2^20 points and the largest NTT --------------------------------------
    extra = kernel_figures(args, ctx, dev, stream, imad_peak, rank, world)
    if rank == 0 and not args.no_pcd_step:
        idx.close()
        extra["pcd_step"] = pcd_step_figure(args, ctx, dev, stream, log)
    if rank == 0 and not args.no_gm17:
        extra["gm17"] = gm17_figure(args, ctx, dev, stream, log)
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample_log = min(log_n, args.cpu_sample_log_n)
        co, cinst = cpu_sample_setup(sample_log)
        threads = co.hw_threads()
        cpu_prove_once(co, cinst, threads, rs[0][0], rs[0][1])
        reps, t = 0, 0.0
        while t < 10.0 and reps < 8:
            t += cpu_prove_once(co, cinst, threads, rs[reps][0], rs[reps][1])
            reps += 1
        frac = 2.0 ** (sample_log - log_n)
        cpu_baseline = {"value": frac * reps / t, "unit": "proofs/s", "cores": threads, "kind": "port",
                        "sample": "%d Groth16 proofs at 2^%d constraints (2^%d of the workload, scaled linearly), "
                                  "oracle/c on all host threads" % (reps, sample_log, sample_log - log_n)}
    if rank != 0:
        return
    # ---- roofline of the dominant kernel class ----------------------------------------------------------
    names = ["msm_digits_sort", "msm_accumulate_g1", "msm_accumulate_g2", "msm_reduce", "msm_horner", "ntt",
             "spmv_qap", "assemble"]
    total_ms = sum(prof["ms"]) or 1.0
    shares = {names[i]: round(prof["ms"][i] / total_ms, 4) for i in range(NC)}
    dom = max((1, 2), key=lambda i: prof["ms"][i])
    deg = 1 if dom == 1 else 2
    imads = prof["units"][dom] * MADD_MODMULS[deg] * MODMUL_IMADS
    achieved = imads / (prof["ms"][dom] * 1e-3) / 1e12 if prof["ms"][dom] > 0 else 0.0
    hbm_peak, hbm_src = load_peaks()
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            traffic = json.load(f).get(names[dom], {}).get("bytes_per_launch")
    except Exception:
        pass
    roofline = {
        "kernel": names[dom], "bound": "imad", "achieved": achieved, "peak": imad_peak / 1e12, "unit": "TIMAD/s",
        "frac": achieved / (imad_peak / 1e12), "traffic": traffic,
        "traffic_note": "DRAM bytes of one launch from the committed ncu capture (2^20 points, uniform scalars); "
                        "algorithmic: 80 B per gathered point + 4 B per entry",
        "peak_source": "measured live: IMAD.WIDE.U32 issue rate of the fmaheavy pipe (32 lanes/clk/SM), max of the "
                       "independent accumulate form (%.2f T/s) and the carry-chain form (%.2f T/s)" % (imad_indep / 1e12, imad_chain / 1e12),
        "launch_ms_avg": prof["ms"][dom] / max(prof["spans"][dom], 1),
        "work": "bucket entries x %d Montgomery products (XYZZ mixed add 8M+2S) x %d IMAD" % (MADD_MODMULS[deg], MODMUL_IMADS),
    }
    value = world * args.steps / (ms_dev * 1e-3)
    e2e_value = world * args.steps / (ms_e2e * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": "proofs/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32 (10x32-bit limbs, Montgomery, integer only)", "data": "synthetic",
        "config": {"workload": "groth16_mnt4_298_domain_2^%d" % log_n, "constraints": inst["m"], "variables": nvars,
                   "msm_lengths": {"h": (1 << log_n) - 1, "l": inst["num_witness"], "a": nvars - 1, "b_g1": nvars - 1,
                                   "b_g2": nvars - 1},
                   "per_gpu": "independent instance per GPU (PCD nodes), no collective",
                   "precomputed_window_tables": not args.no_precompute, "proofs_in_flight": nfl,
                   "l2": "per-step inputs (proving-key tables, several GB) exceed the 126 MB L2"},
        "e2e": {"value": e2e_value, "unit": "proofs/s", "h2d_bytes_per_step": nvars * 40 + 80,
                "d2h_bytes_per_step": 320, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": n_launches,
        "clocks": clocks,
        "roofline": roofline,
        "kernel_time_shares": shares,
        "kernel_ms_per_step": {names[i]: round(prof["ms"][i] / args.steps, 3) for i in range(NC)},
        "serialized_ms_per_step": ms_serial / args.steps,
        "kernel_timing_note": "kernel_* and roofline come from a second pass over the same proofs with the five MSM "
                              "streams serialised (pcdgpu_set_concurrency(0)); value / ms_per_step are the overlapped run",
        "imad_peak_measured": {"independent_TIMAD_s": imad_indep / 1e12, "carry_chain_TIMAD_s": imad_chain / 1e12},
        "hbm_peak_GBps": {"value": hbm_peak, "source": hbm_src},
    }
    line.update(extra)
    if cpu_baseline:
        line["cpu_baseline"] = cpu_baseline
    emit(line)


def pcd_step_figure(args, ctx, dev, stream, log):
    """One PCD step as ECCyclePCD::prove issues it (mod.rs:171,179): the main proof on MNT4-298, then --
    strictly after it, because the helper circuit's witness contains the main proof -- the helper proof on
    MNT6-298 (G2 over Fq3).  Synthetic circuits of PCD-like size: main domain 2^18, helper
    2^16; constraint synthesis and the CRH (CPU work above the SNARK seam) are not part of the figure."""
    import torch

    import pcd_b200
    from pcd_b200 import synthetic
    sides = []
    for pairing, log_n in ((pcd_b200.MNT4_298, args.pcd_main_log_n), (pcd_b200.MNT6_298, args.pcd_help_log_n)):
        inst = synthetic.make_groth16_instance(ctx, pairing, log_n, seed=77 + pairing)
        g = pcd_b200.Groth16(ctx, pairing)
        idx = g.index(pcd_b200.ProvingKey(pairing=pairing, **inst["pk"]),
                      pcd_b200.ConstraintMatrices(pairing, inst["num_inputs"], inst["num_witness"], inst["A"], inst["B"],
                                                  inst["C"]), precompute=True)
        z = torch.from_numpy(inst["z"].view(np.int64)).to(dev)
        p = inst["p"]
        r_i, s_i = 0x1234567 * 3 ** 70 % p, 0x7654321 * 5 ** 60 % p
        lim = lambda v: np.array([(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(5)], dtype=np.uint64)
        proof = g.create_proof_dev(idx, z.data_ptr(), lim(r_i), lim(s_i))
        if not np.array_equal(proof.affine_limbs(), synthetic.expected_proof(ctx, inst, r_i, s_i)):
            raise SystemExit("bench: PCD-step proof (pairing %d) does not match its discrete logarithms" % pairing)
        sides.append((g, idx, z, lim(r_i), lim(s_i)))
    if log:
        log("PCD step: main 2^%d (MNT4-298) and helper 2^%d (MNT6-298) proofs verified" % (args.pcd_main_log_n,
                                                                                        args.pcd_help_log_n))

    def step():
        for g, idx, z, r, s in sides:  # helper strictly after main
            g.create_proof_dev(idx, z.data_ptr(), r, s)

    for _ in range(3):
        step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    reps = 5
    per = []
    for _ in range(reps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    for g, idx, z, r, s in sides:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        g.create_proof_dev(idx, z.data_ptr(), r, s)
        b.record(stream)
        torch.cuda.synchronize()
        per.append(a.elapsed_time(b))
        idx.close()
    return {"steps_per_s": 1e3 / ms, "ms_per_step": ms, "main": {"pairing": "MNT4-298", "domain": "2^%d" % args.pcd_main_log_n,
            "ms": per[0]}, "helper": {"pairing": "MNT6-298", "domain": "2^%d" % args.pcd_help_log_n, "ms": per[1]},
            "note": "prover kernels only (witness map + 4 G1 MSM + 1 G2 MSM + assembly per proof); main then helper, "
                    "sequential as in ECCyclePCD::prove"}


def gm17_figure(args, ctx, dev, stream, log):
    """GM17 proofs/s on MNT4-298 (the second SNARK the reference plugs into ECCyclePCD, tests/mnt4_gm17.rs:27-28):
    2^(k-1) - 2 constraints, SAP domain 2^k; the proof is checked against GM17's verification equations in the
    exponent (known trapdoor) before it is timed."""
    import torch

    import pcd_b200
    from pcd_b200 import synthetic
    k = args.gm17_log_n
    m = (1 << (k - 1)) - 2
    inst = synthetic.make_gm17_instance(ctx, pcd_b200.MNT4_298, m, seed=4242, verbose=log)
    assert inst["domain_size"] == 1 << k
    g = pcd_b200.GM17(ctx, pcd_b200.MNT4_298)
    idx = g.index(pcd_b200.GM17ProvingKey(pairing=0, **inst["pk"]),
                  pcd_b200.ConstraintMatrices(0, inst["num_inputs"], inst["num_witness"], inst["A"], inst["B"], inst["C"]),
                  precompute=True)
    z = torch.from_numpy(inst["z"].view(np.int64)).to(dev)
    p = inst["p"]
    d1, d2, r = 0x1234567 * 3 ** 70 % p, 0x7654321 * 5 ** 60 % p, 0xabcdef1 * 7 ** 50 % p
    lim = lambda v: np.array([(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(5)], dtype=np.uint64)
    proof = g.create_proof_dev(idx, z.data_ptr(), lim(d1), lim(d2), lim(r))
    if not np.array_equal(proof.affine_limbs(), synthetic.expected_gm17_proof(ctx, inst, d1, d2, r)):
        raise SystemExit("bench: GM17 proof does not satisfy the verification equations in the exponent")
    if log:
        log("GM17 proof (SAP domain 2^%d) verified against the trapdoor" % k)
    for _ in range(3):
        g.create_proof_dev(idx, z.data_ptr(), lim(d1), lim(d2), lim(r))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    reps = 5
    e0.record(stream)
    for _ in range(reps):
        g.create_proof_dev(idx, z.data_ptr(), lim(d1), lim(d2), lim(r))
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    idx.close()
    return {"proofs_per_s": 1e3 / ms, "ms_per_proof": ms, "pairing": "MNT4-298", "constraints": m,
            "sap_domain": "2^%d" % k, "msm_lengths": {"a": inst["num_inputs"] + inst["num_witness"] + m, "b_g2": "same",
                                                       "c1": "same - 2", "c2": "same", "g": (1 << k) + 1},
            "note": "SAP witness map (5 NTTs) + 4 G1 MSMs + 1 G2 MSM + assembly; one proof at a time"}


def kernel_figures(args, ctx, dev, stream, imad_peak, rank=0, world=1):
    """G1 MSM Mpts/s at 2^20 (uniform scalars; resident bases with and without the window tables) and
    the largest radix-2 NTT (coset FFT over r4), each timed with CUDA events on the launching stream."""
    import torch

    import pcd_b200
    from pcd_b200 import synthetic
    out = {}
    hbm_peak, _ = load_peaks()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, reps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    n = 1 << args.msm_log_n
    pts = synthetic.random_points_dev(ctx, pcd_b200.MNT4_G1, n, seed=3)
    sc = torch.from_numpy(synthetic.random_limbs(n, 0, 9).view(np.int64)).to(dev)
    res = torch.zeros(64, dtype=torch.int64, device=dev)
    ms_plain = timed(lambda: ctx.msm_dev(0, pts.data_ptr(), sc.data_ptr(), n, res.data_ptr()), 5)
    bases = pcd_b200.Bases(ctx, 0, pts.cpu().numpy().view(np.uint64), precompute=True)
    ms_pre = timed(lambda: bases.msm_dev(sc.data_ptr(), n, res.data_ptr()), 5)
    if world > 1:
        # the same MSM sharded by point range: every rank keeps n / world points resident, the xyzz
        # partials are all-gathered (NCCL) and summed on every rank; time = max over ranks
        import torch.distributed as dist
        from pcd_b200.sharding import gather_partials, shard_range
        lo, hi = shard_range(n, world, rank)
        pts_host = pts.cpu().numpy().view(np.uint64)
        shard = pcd_b200.Bases(ctx, 0, pts_host[lo:hi], precompute=True)
        sc_shard = sc[lo:hi].contiguous()

        def sharded():
            shard.msm_dev(sc_shard.data_ptr(), hi - lo, res.data_ptr())
            parts = gather_partials(ctx.xyzz_download(0, res.data_ptr()), device=dev)
            return ctx.xyzz_sum(0, parts)

        full = bases.msm(sc.cpu().numpy().view(np.uint64))
        ok = bool(np.array_equal(sharded(), full))
        for _ in range(2):
            sharded()
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            sharded()
        torch.cuda.synchronize()
        dt = torch.tensor([(time.perf_counter() - t0) / 5], dtype=torch.float64, device=dev)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        out["g1_msm_sharded"] = {"points": n, "gpus": world, "ms": float(dt.item()) * 1e3,
                                 "mpts_per_s": n / float(dt.item()) / 1e6, "matches_single_gpu": ok,
                                 "collective": "all_gather of one 160-byte xyzz partial per rank"}
        shard.close()
    bases.close()
    del pts
    out["g1_msm"] = {"points": n, "scalars": "uniform 298-bit", "mpts_per_s": n / ms_pre / 1e3, "ms": ms_pre,
                     "mode": "resident bases with precomputed window tables",
                     "variable_base_mpts_per_s": n / ms_plain / 1e3, "variable_base_ms": ms_plain}
    if world == 1 and not args.no_sweep:
        # the other sizes BASELINE.json names for the MSM (2^16 .. 2^22), resident tables, uniform scalars
        sweep = {}
        for lg in (16, 18, 22):
            if lg == args.msm_log_n:
                continue
            m = 1 << lg
            p_ = synthetic.random_points_dev(ctx, pcd_b200.MNT4_G1, m, seed=30 + lg)
            s_ = torch.from_numpy(synthetic.random_limbs(m, 0, 40 + lg).view(np.int64)).to(dev)
            b_ = pcd_b200.Bases(ctx, 0, p_.cpu().numpy().view(np.uint64), precompute=True)
            del p_
            t_ = timed(lambda: b_.msm_dev(s_.data_ptr(), m, res.data_ptr()), 5)
            sweep["2^%d" % lg] = {"ms": t_, "mpts_per_s": m / t_ / 1e3}
            b_.close()
            del s_
        out["g1_msm_sweep"] = sweep
    log_n = args.ntt_log_n
    x = torch.from_numpy(synthetic.random_limbs(1 << min(log_n, 20), 0, 5).view(np.int64)).to(dev)
    if log_n > 20:
        x = x.repeat(1 << (log_n - 20), 1)
    ms_ntt = timed(lambda: ctx.ntt_dev(0, x.data_ptr(), log_n, False, True), 5)
    gbs = 2 * 40 * (1 << log_n) / (ms_ntt * 1e-3) / 1e9
    butterflies = (1 << (log_n - 1)) * log_n
    timad = butterflies * MODMUL_IMADS / (ms_ntt * 1e-3) / 1e12
    out["ntt"] = {"log_n": log_n, "field": "r4 (MNT4-298 Fr)", "flavour": "coset_fft", "ms": ms_ntt, "GBps": gbs}
    if world == 1 and not args.no_sweep:
        nsw = {}
        for lg in (16, 20):
            if lg >= log_n:
                continue
            t_ = timed(lambda: ctx.ntt_dev(0, x.data_ptr(), lg, False, True), 10)
            nsw["2^%d" % lg] = {"ms": t_, "GBps": 2 * 40 * (1 << lg) / (t_ * 1e-3) / 1e9}
        y = x[:1 << 17].contiguous()
        t_ = timed(lambda: ctx.ntt_dev(1, y.data_ptr(), 17, False, True), 10)  # the helper field's largest radix-2 domain
        nsw["q4_2^17"] = {"ms": t_, "GBps": 2 * 40 * (1 << 17) / (t_ * 1e-3) / 1e9, "note": "L2-resident, not an HBM figure"}
        out["ntt_sweep"] = nsw
    out["roofline_ntt"] = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                           "traffic": None, "imad_frac": timad / (imad_peak / 1e12),
                           "note": "algorithmic bytes 2*N*40; a 298-bit NTT is bound by the integer pipe: imad_frac = "
                                   "(N/2 log2 N butterflies x 210 IMAD) / time / the IMAD.WIDE roof"}
    return out


_REAL_STDOUT = None


def emit(line: dict):
    """the ONE JSON line, on the process's real stdout (fd 1 is pointed at stderr while the job runs so that
    library chatter -- NCCL prints its version banner on stdout -- cannot end up next to it)"""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log-n", type=int, default=int(os.environ.get("PCD_BENCH_LOG_N", "20")))
    ap.add_argument("--msm-log-n", type=int, default=20)
    ap.add_argument("--ntt-log-n", type=int, default=24)
    ap.add_argument("--cpu-sample-log-n", type=int, default=17)
    ap.add_argument("--no-precompute", action="store_true")
    ap.add_argument("--no-gm17", action="store_true", help="skip the GM17 figure")
    ap.add_argument("--gm17-log-n", type=int, default=18, help="SAP domain of the GM17 figure (2^k)")
    ap.add_argument("--inflight", type=int, default=int(os.environ.get("PCD_BENCH_INFLIGHT", "0")),
                    help="independent proofs issued concurrently per GPU (each on its own context); 0 = the largest of "
                         "4, 3, 5 that divides --steps (so that every context proves the same number), else 2")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pcd-step", action="store_true")
    ap.add_argument("--no-sweep", action="store_true", help="skip the MSM / NTT size sweeps")
    ap.add_argument("--pcd-main-log-n", type=int, default=18)
    ap.add_argument("--pcd-help-log-n", type=int, default=16)
    ap.add_argument("--no-concurrency", action="store_true", help="run the five MSMs of a proof on one stream")
    args = ap.parse_args()
    if args.inflight <= 0:  # measured at 2^20 (8 steps): 2 in flight 26.1 ms per proof, 3: 25.4, 4: 24.9
        args.inflight = next((f for f in (4, 3, 5) if args.steps % f == 0), 2)
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_gpu(args, rank, local_rank, world)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
