#!/usr/bin/env python3
"""gpurun_out/r02_step_main_raw.csv + r02_list_main.csv (tools/ncu_step.sh) -> profiles/r02_ncu_step_main.csv,
profiles/r02_traffic.json, profiles/r02_launches_main_proof.csv"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
note = sys.argv[1] if len(sys.argv) > 1 else "round 2 final kernels"
with open(os.path.join(P, "r02_ncu_step_main.csv"), "w") as f:
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), os.path.join(G, "r02_step_main_raw.csv"),
                           "ncu --set full --clock-control none, tools/ncu_step.sh: SECOND Groth16 proof of the PCD step's MAIN "
                           "circuit (MNT4-298, 2^18, witness-like assignment, resident window tables), MSM lanes serialised; "
                           "accumulation, bucket-reduction and double-scalar kernels in launch order (b_g2, a, b_g1 + chain, l, h); "
                           + note], stdout=f)
rows = list(csv.reader(open(os.path.join(G, "r02_step_main_raw.csv"))))
hdr, units, data = rows[0], rows[1], rows[2:]
ki, ri, wi, ti = (hdr.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"))
tobytes = lambda v, u: float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
acc = {"msm_accumulate_g1": [], "msm_accumulate_g2_fq2": []}
for r in data:
    if "msm_accumulate_kernel" in r[ki]:
        acc["msm_accumulate_g1" if "Mnt4G1" in r[ki] else "msm_accumulate_g2_fq2"].append(
            {"kernel": r[ki].split("(")[0], "dram_bytes": tobytes(r[ri], units[ri]) + tobytes(r[wi], units[wi]), "ms": float(r[ti])})
out = {cls: {"bytes_per_launch": sum(x["dram_bytes"] for x in l) / len(l), "launches": l,
             "source": "profiles/r02_ncu_step_main.csv: dram__bytes_read.sum + dram__bytes_write.sum of msm_accumulate_kernel, "
                       "averaged over the launches of the class (a, b_g1, l, h for G1); " + note} for cls, l in acc.items() if l}
json.dump(out, open(os.path.join(P, "r02_traffic.json"), "w"), indent=1)
rows = [r for r in csv.reader(open(os.path.join(G, "r02_list_main.csv"))) if len(r) > 5]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
sel = [r for r in rows[1:] if not any(x in r[ki] for x in ("precompute", "fixed_", "ntt_tables"))]
half = sel[len(sel) // 2:]


def short(k):
    m = re.match(r"(?:void )?([\w:]+)(<[^(]*>)?", k)
    name, t = m.group(1), m.group(2) or ""
    for c in ("CurveMnt4G1", "CurveMnt4G2", "CurveMnt6G1", "CurveMnt6G2S", "CurveMnt6G2"):
        if c in t:
            name += "<" + c + ">"
            break
    return "cub::" + name.split("::")[-1] if name.startswith("cub::") else name


scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1, "msecond": 1}[half[0][ui]]
tot = collections.OrderedDict()
for r in half:
    e = tot.setdefault(short(r[ki]), [0, 0.0])
    e[0] += 1
    e[1] += float(r[vi].replace(",", "")) * scale
total = sum(v for _, v in tot.values())
with open(os.path.join(P, "r02_launches_main_proof.csv"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (tools/ncu_step.sh): every launch of the SECOND main proof "
            "of the PCD step (MNT4-298, 2^18), lanes serialised; cold-cache, serialised times -- compare SHARES with bench.py's "
            "kernel_time_shares, not absolutes; " + note + "\n")
    f.write("kernel,launches,total_ms,share\n")
    for k, (n, v) in sorted(tot.items(), key=lambda x: -x[1][1]):
        f.write("%s,%d,%.3f,%.4f\n" % (k, n, v, v / total))
    f.write("TOTAL,%d,%.3f,1.0\n" % (sum(n for n, _ in tot.values()), total))
print(json.dumps({k: v["bytes_per_launch"] for k, v in out.items()}))
