mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "dwconv7x7" 2>&1 | tail -8
python tools/prof_kernels.py dwconv_ln 2>&1 | tail -1
timeout 600 python bench.py --model convnext_base --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_convnext_base.json 2> gpurun_out/bench_convnext_base.err
tail -3 gpurun_out/bench_convnext_base.err; cut -c1-300 gpurun_out/bench_convnext_base.json
