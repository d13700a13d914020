#!/bin/bash
# compute-sanitizer memcheck / racecheck over the shared-memory NTT, the MSM kernels (atomics, bucket walks, the
# lane-cooperative reduction) and the Marlin device-vector kernels, on small instances (the tools slow kernels ~50x).
# Run on the GPU box:  bash tools/sanitize.sh  -> gpurun_out/sanitize_{memcheck,racecheck}.log
mkdir -p gpurun_out
SEL='test_ntt_golden or test_ntt_vs_oracle or test_msm_golden or test_msm_heavy_bucket or (test_msm_vs_oracle and 1000) or test_groth16_golden or test_gm17_golden or (test_gpu_marlin_matches_oracle and 13)'
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 99 \
    python -m pytest tests/test_gpu_parity.py tests/test_marlin.py -m gpu -x -q -k "$SEL" \
    > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/sanitize_$tool.log
  tail -4 gpurun_out/sanitize_$tool.log
done
