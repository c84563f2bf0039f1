mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee -a gpurun_out/tests.log
for m in vit_base_patch16_224 convnext_base swin_base_patch4_window7_224 efficientnet_b4 resnet50; do
timeout 600 python bench.py --model $m --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$m.json 2> gpurun_out/bench_$m.err
tail -2 gpurun_out/bench_$m.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$m.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("$m", round(d["value"]), round(d["ms_per_step"],2), {k:round(v,2) for k,v in r["families_ms"].items() if v>0.05}, "frac",round(r["frac"],3), r["bound"], d["clocks"]["sm_mhz"], "e2e", round(d["e2e"]["value"]))
PY
done
