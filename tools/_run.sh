mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "attention" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_models_gpu.py -m gpu -x -q -k "vit or deit or golden or graph or uint8 or pipeline" 2>&1 | tail -3
for pr in 1 0; do
TFIMM_B200_VIT_PRUNE=$pr timeout 600 python bench.py --model vit_base_patch16_224 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_vit_prune$pr.json 2> gpurun_out/bench_vit_prune$pr.err
tail -2 gpurun_out/bench_vit_prune$pr.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_vit_prune$pr.json").read().strip().splitlines()[-1])
print("prune=$pr", round(d["value"]), round(d["ms_per_step"],2), d["roofline"]["families_ms"], d["clocks"]["sm_mhz"], d["gpu_launches"]//25)
PY
done
