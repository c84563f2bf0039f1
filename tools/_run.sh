mkdir -p gpurun_out
for mode in 1cta auto; do
for m in vit_base_patch16_224 swin_base_patch4_window7_224 convnext_base; do
  TFIMM_B200_GEMM=$mode timeout 600 python bench.py --model $m --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${m}_$mode.json 2> gpurun_out/bench_${m}_$mode.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${m}_$mode.json").read().strip().splitlines()[-1])
print("$mode $m", round(d["value"]), round(d["ms_per_step"],2), d["roofline"]["families_ms"], d["clocks"])
PY
done; done
