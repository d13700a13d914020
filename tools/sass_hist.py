#!/usr/bin/env python3
"""Opcode histograms of selected kernels from the built objects (cuobjdump -sass): the committed evidence that the
field product is IMAD.WIDE.U32(.X) carry chains, that the NTT's butterflies and the lane-cooperative interpreter use
the same inlined product, and what else the hot loops issue.
usage: sass_hist.py > profiles/r02_sass_histograms.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "pcd_b200", "csrc", "build")
WANT = [  # (object, regex on the demangled-ish mangled name, label)
    ("msm_c0.o", r"msm_accumulate_kernelI11CurveMnt4G1Lb1", "msm_accumulate_kernel<CurveMnt4G1, tables>"),
    ("msm_c1.o", r"msm_accumulate_kernelI11CurveMnt4G2Lb1", "msm_accumulate_kernel<CurveMnt4G2, tables>"),
    ("msm_c3.o", r"msm_accumulate_sliced_kernelI12CurveMnt6G2SLb1", "msm_accumulate_sliced_kernel<CurveMnt6G2S, tables>"),
    ("msm_c0.o", r"wec_reduce_kernelI11CurveMnt4G1", "wec_reduce_kernel<CurveMnt4G1> (incl. wec_exec)"),
    ("msm_c3.o", r"wec_reduce_kernelI11CurveMnt6G2", "wec_reduce_kernel<CurveMnt6G2> (incl. wec_exec)"),
    ("ntt.o", r"ntt_pass_kernelI2FpI8ParamsR4", "ntt_pass_kernel<FpR4>"),
    ("groth16.o", r"bench_modmul_kernelI2FpI8ParamsQ4", "bench_modmul_kernel<FpQ4> (two dependent Montgomery products per iteration)"),
    ("groth16.o", r"spmv_kernelI2FpI8ParamsR4", "spmv_kernel<FpR4>"),
]


def functions(obj):
    out = subprocess.run(["cuobjdump", "-sass", os.path.join(BUILD, obj)], capture_output=True, text=True).stdout
    cur, table = None, {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            table[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            table[cur][m.group(1)] += 1
    return table


def main():
    cache = {}
    for obj, pat, label in WANT:
        if obj not in cache:
            cache[obj] = functions(obj)
        hits = [(n, c) for n, c in cache[obj].items() if re.search(pat, n)]
        print("== %s   [%s]" % (label, obj))
        if not hits:
            print("   (not found)")
            continue
        total = collections.Counter()
        for n, c in hits:  # the kernel and the functions it calls that match (e.g. out-of-line products) are listed apart
            total.update(c)
        s = sum(total.values())
        wide = sum(v for k, v in total.items() if k.startswith("IMAD.WIDE"))
        print("   %d SASS instructions; IMAD.WIDE* %d (%.1f %%), of which .X (carry-in) %d" % (
            s, wide, 100.0 * wide / s, sum(v for k, v in total.items() if k.startswith("IMAD.WIDE") and ".X" in k)))
        for k, v in total.most_common(14):
            print("   %-28s %6d  %5.1f %%" % (k, v, 100.0 * v / s))
    return 0


if __name__ == "__main__":
    sys.exit(main())
