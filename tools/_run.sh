timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "attention" 2>&1 | tail -3
for rep in 1 2; do
for lib in /root/repo/gpurun_libA.so /root/repo/tensorflow-image-models_b200/tfimm/backend/libtfimm_b200.so; do
echo "== $lib"; TFIMM_B200_LIB=$lib python tools/prof_kernels.py attn | tail -1
done; done
