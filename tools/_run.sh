mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "se_gate" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_models_gpu.py -m gpu -x -q -k "efficientnet or resnet" 2>&1 | tail -3
for m in efficientnet_b4; do
timeout 600 python bench.py --model $m --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$m.json 2> gpurun_out/bench_$m.err
tail -2 gpurun_out/bench_$m.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$m.json").read().strip().splitlines()[-1])
print("$m", round(d["value"]), round(d["ms_per_step"],2), d["roofline"]["families_ms"])
PY
done
