#!/bin/bash
# 1 / 2 / 4 / 8-GPU weak-scaling run of the headline bench (one box, NCCL), under gpurun --gpus 8.
mkdir -p gpurun_out
MODEL=${MODEL:-vit_base_patch16_224}
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/scale_gpus.txt
for n in 1 2 4 8; do
  if [ "$n" = "1" ]; then
    timeout 600 python bench.py --model $MODEL --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/scale_${MODEL}_1.json 2> gpurun_out/scale_${MODEL}_1.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
      bench.py --model $MODEL --gpus $n --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/scale_${MODEL}_$n.json 2> gpurun_out/scale_${MODEL}_$n.err
  fi
  tail -2 gpurun_out/scale_${MODEL}_$n.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/scale_${MODEL}_$n.json").read().splitlines() if l.startswith("{")][-1])
    print("N=$n", round(d["value"]), "img/s", round(d["ms_per_step"],2), "ms", d["clocks"], "e2e", round(d["e2e"]["value"]))
except Exception as e:
    print("N=$n failed", e)
PY
done
