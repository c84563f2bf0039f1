timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gemm" 2>&1 | tail -2
for rep in 1 2; do
for lib in /root/repo/gpurun_libA.so /root/repo/tensorflow-image-models_b200/tfimm/backend/libtfimm_b200.so; do
echo "== $lib"
TFIMM_B200_LIB=$lib python tools/bench_gemm.py 50432 3072 768 gelu bf16 0 2 | tail -1
TFIMM_B200_LIB=$lib python tools/bench_gemm.py 50432 3072 768 none bf16 0 2 | tail -1
TFIMM_B200_LIB=$lib python tools/bench_gemm.py 50432 2304 768 none bf16 0 2 | tail -1
TFIMM_B200_LIB=$lib python tools/bench_gemm.py 50176 2048 512 gelu bf16 0 2 | tail -1
TFIMM_B200_LIB=$lib python tools/bench_gemm.py 16384 8192 8192 none bf16 0 2 | tail -1
done; done
