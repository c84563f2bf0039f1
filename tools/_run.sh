mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "implicit_gemm_conv or gemm_bf16" 2>&1 | tail -12
timeout 600 python -m pytest tests/test_models_gpu.py -m gpu -x -q -k "resnet" 2>&1 | tail -4
for c in implicit im2col; do
TFIMM_B200_CONV=$c timeout 600 python bench.py --model resnet50 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_resnet50_$c.json 2> gpurun_out/bench_resnet50_$c.err
tail -2 gpurun_out/bench_resnet50_$c.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_resnet50_$c.json").read().strip().splitlines()[-1])
print("$c", round(d["value"]), round(d["ms_per_step"],2), d["roofline"]["families_ms"])
PY
done
