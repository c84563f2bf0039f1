#!/bin/bash
# ncu --set full captures of the hot kernels at their benchmark shapes (one launch each).
mkdir -p gpurun_out
DEFAULT_CASES="attn:vit_attention_tc2 dwconv_ln:dwconv7_ln_cluster window_attn:window_attention_bf16 gemm_fc1:gemm_bf16_tcgen05 gemm_proj:gemm_bf16_tcgen05 dwconv_act:dwconv_act_pairs"
for spec in ${CASES:-$DEFAULT_CASES}; do
  case=${spec%%:*}; kern=${spec##*:}
  python tools/prof_kernels.py $case 2>&1 | tail -1 | tee -a gpurun_out/prof_timing.txt
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kern -s 2 -c 1 -f -o gpurun_out/prof_$case \
      python tools/prof_kernels.py $case > gpurun_out/prof_$case.log 2>&1
  tail -2 gpurun_out/prof_$case.log
done
ls -la gpurun_out/*.ncu-rep
