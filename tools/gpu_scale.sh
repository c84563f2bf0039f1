#!/bin/bash
# Multi-GPU bench (NCCL all-gather of logits), run under gpurun --gpus N.
mkdir -p gpurun_out
N=${N:-2}
MODEL=${MODEL:-vit_base_patch16_224}
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/scale_gpus.txt
for n in 1 $N; do
  if [ "$n" = "1" ]; then
    timeout 900 python bench.py --model $MODEL --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/scale_${MODEL}_1.json 2> gpurun_out/scale_${MODEL}_1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --model $MODEL --gpus $n --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/scale_${MODEL}_$n.json 2> gpurun_out/scale_${MODEL}_$n.err
  fi
  tail -2 gpurun_out/scale_${MODEL}_$n.err; cut -c1-300 gpurun_out/scale_${MODEL}_$n.json
done
echo "=== reference arm"
timeout 600 python bench.py --impl reference --model $MODEL --steps 3 --warmup 1 > gpurun_out/ref_${MODEL}.json 2> gpurun_out/ref_${MODEL}.err
cut -c1-400 gpurun_out/ref_${MODEL}.json; tail -2 gpurun_out/ref_${MODEL}.err
