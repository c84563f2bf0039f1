mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gemm" 2>&1 | tail -4
python tools/bench_gemm.py 50432 3072 768 gelu bf16 0 2 0 0 | tail -1
python tools/bench_gemm.py 50432 3072 768 gelu bf16 0 2 1 0 | tail -1
python tools/bench_gemm.py 50432 3072 768 gelu bf16 0 256 0 0 | tail -1
python tools/bench_gemm.py 50432 2304 768 none bf16 0 2 1 0 | tail -1
for f in 0 1; do
  TFIMM_B200_LN_FOLD=$f timeout 600 python bench.py --model vit_base_patch16_224 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_vit_fold$f.json 2> gpurun_out/bench_vit_fold$f.err
  tail -2 gpurun_out/bench_vit_fold$f.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_vit_fold$f.json").read().strip().splitlines()[-1])
print("fold=$f", round(d["value"]), round(d["ms_per_step"],2), d["roofline"]["families_ms"], d["clocks"])
PY
done
