#!/bin/bash
# Runs the GPU kernel numerics tests group by group under a timeout so that a hung kernel
# cannot take the whole call (or the box) down.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for grp in gemm_bf16_plain gemm_bf16_epilogues gemm_bf16_inplace gemm_bf16_strided gemm_f32 layernorm attention patchify assemble; do
  echo "=== $grp" | tee -a gpurun_out/kernels.log
  timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "$grp" 2>&1 | tail -25 | tee -a gpurun_out/kernels.log
done
