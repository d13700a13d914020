import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pcd_b200
ctx = pcd_b200.Context(0)
names = {0: "imad_wide independent", 3: "imad_wide carry chain", 1: "modmul32 r4 (64 warps/SM)", 2: "modmul32 q4 (64 warps/SM)",
         4: "modmul30 r4 (64 warps/SM)", 5: "modmul30 q4 (64 warps/SM)", 6: "modmul30 q4 2 chains/thread",
         7: "modmul30 q4 16 warps/SM", 8: "modmul30 q4 8 warps/SM", 9: "modmul32 q4 8 warps/SM",
         10: "mul_sc r4 (64 warps/SM)", 11: "mul_sc q4 (64 warps/SM)", 12: "mul_sc q4 8 warps/SM",
         13: "mul_cc r4 (64 warps/SM)", 14: "mul_cc q4 (64 warps/SM)", 15: "mul_cc q4 8 warps/SM",
         16: "imad_wide carry-out + IADD3.X count"}
for mode in (0, 3, 16, 13, 14, 15, 10, 11, 12, 4, 5):
    ops, ms = ctx.bench_imad(mode, 2000)
    print("%-32s %.3e /s  (%.2f ms)" % (names[mode], ops, ms), flush=True)
