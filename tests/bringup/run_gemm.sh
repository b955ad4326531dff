#!/bin/bash
# Run each bring-up group in its own process so a device trap in one cannot poison the others.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gemm_smi.log 2>&1
for g in "$@"; do
  timeout 400 python tests/bringup/gemm_bringup.py $g > gpurun_out/gemm_$g.log 2>&1
  echo "group $g exit $?" | tee -a gpurun_out/gemm_summary.log
  tail -5 gpurun_out/gemm_$g.log
done
