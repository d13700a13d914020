#!/bin/bash
# ncu --set full capture of the PCD step's main (MNT4-298, 2^18) and helper (MNT6-298, 2^16) proofs, lanes serialised
# (tools/ncu_step.py proves each twice; the raw csv keeps both proofs, tools/ncu_summary.py is given the second one).
# Run under gpurun from the repo root; writes gpurun_out/r02_step_{main,help}_raw.csv
set -u
K='regex:msm_accumulate|msm_fold_parts|msm_heavy_finish|wec_reduce|wec_multi_mul|ntt_pass'
for which in main help; do
  NCU_STEP=$which ncu --set full --clock-control none --import-source on -k "$K" -c 100 -f \
    -o gpurun_out/r02_step_$which python tools/ncu_step.py > gpurun_out/ncu_step_$which.log 2>&1
  ncu -i gpurun_out/r02_step_$which.ncu-rep --page raw --csv > gpurun_out/r02_step_${which}_raw.csv 2>> gpurun_out/ncu_step_$which.log
  rm -f gpurun_out/r02_step_$which.ncu-rep
  wc -l gpurun_out/r02_step_${which}_raw.csv
done
