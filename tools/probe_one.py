"""Run chosen modes of the integer-pipe microbenchmark (pcdgpu_bench_imad) -- an ncu target:
  ncu --set full -k regex:bench_ ... python tools/probe_one.py 5 14"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pcd_b200
ctx = pcd_b200.Context(0)
for mode in [int(a) for a in sys.argv[1:]]:
    ops, ms = ctx.bench_imad(mode, 200)
    print(mode, "%.3e /s  (%.2f ms)" % (ops, ms), flush=True)
