#!/usr/bin/env python3
"""registers / stack / spill bytes of every kernel of libpcdgpu.so from the ptxas -v logs of the build
(pcd_b200/csrc/build/*.ptxas.log).  usage: ptxas_table.py > profiles/r02_ptxas_registers_spills.txt"""
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = []
for log in sorted(glob.glob(os.path.join(ROOT, "pcd_b200", "csrc", "build", "*.ptxas.log"))):
    text = open(log).read()
    for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, "
                         r"(\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers", text):
        name = m.group(1)
        try:
            name = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
        except Exception:
            pass
        name = re.sub(r"\(.*$", "", name).replace("void ", "")
        rows.append((os.path.basename(log).replace(".ptxas.log", ""), name, int(m.group(5)), int(m.group(2)), int(m.group(3)),
                     int(m.group(4))))
print("# ptxas -v (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a -O3): every __global__ of libpcdgpu.so")
print("%-9s %-78s %5s %6s %7s %7s" % ("unit", "kernel", "regs", "stack", "spill_st", "spill_ld"))
for r in sorted(rows, key=lambda r: (-r[4], r[0], r[1])):
    print("%-9s %-78s %5d %6d %7d %7d" % (r[0], r[1][:78], r[2], r[3], r[4], r[5]))
spilling = [r for r in rows if r[4] or r[5]]
print("# kernels with spill bytes: %d of %d" % (len(spilling), len(rows)))
