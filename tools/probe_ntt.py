"""Development aid: time the (coset) NTT at a few sizes and print a checksum of the result, so that launch
parameters (PCDGPU_NTT_TILE_LOG, PCDGPU_NTT_THREADS) can be swept across processes and compared."""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pcd_b200  # noqa: E402
from pcd_b200 import synthetic  # noqa: E402

ctx = pcd_b200.Context(0)
dev = torch.device("cuda:0")
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
out = []
for log_n in (16, 20, 24):
    n = 1 << log_n
    base = torch.from_numpy(synthetic.random_limbs(min(n, 1 << 20), 0, 5).view(np.int64)).to(dev)
    x = base.repeat(n // base.shape[0], 1).contiguous()
    y = x.clone()
    ctx.ntt_dev(0, y.data_ptr(), log_n, False, True)
    torch.cuda.synchronize()
    digest = hashlib.sha256(y[:4096].cpu().numpy().tobytes() + y[-4096:].cpu().numpy().tobytes()).hexdigest()[:12]
    ctx.ntt_dev(0, y.data_ptr(), log_n, True, True)
    torch.cuda.synchronize()
    ok = bool(torch.equal(x, y))
    for _ in range(3):
        ctx.ntt_dev(0, y.data_ptr(), log_n, False, True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record(stream)
    for _ in range(reps):
        ctx.ntt_dev(0, y.data_ptr(), log_n, False, True)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    out.append("2^%d: %.3f ms (%.0f GB/s) roundtrip=%s %s" % (log_n, ms, 2 * n * 40 / ms / 1e6, ok, digest))
print(os.environ.get("PCDGPU_NTT_TILE_LOG", "11"), os.environ.get("PCDGPU_NTT_THREADS", "256"), " | ".join(out))
