#!/bin/bash
# Evidence for profiles/ at the PCD step's main proof (MNT4-298, 2^18; tools/ncu_step.py proves it twice with the lanes
# serialised): (1) the launch list of both proofs, (2) an `ncu --set full` capture of the second proof's accumulation,
# bucket-reduction and double-scalar kernels (11 launches per proof).  Run under gpurun from the repo root.
set -u
mkdir -p gpurun_out
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_list_main.csv \
  python tools/ncu_step.py > gpurun_out/ncu_list_main.log 2>&1
K='regex:msm_accumulate_kernel|wec_reduce_kernel|wec_multi_mul_kernel'
timeout 420 ncu --set full --clock-control none -k "$K" --launch-skip 11 -c 11 -f -o gpurun_out/r02_step_main \
  python tools/ncu_step.py > gpurun_out/ncu_step_main.log 2>&1
ncu -i gpurun_out/r02_step_main.ncu-rep --page raw --csv > gpurun_out/r02_step_main_raw.csv 2>> gpurun_out/ncu_step_main.log
rm -f gpurun_out/r02_step_main.ncu-rep
wc -l gpurun_out/r02_list_main.csv gpurun_out/r02_step_main_raw.csv
