#!/bin/bash
# One GPU visit: kernel + model parity tests, bench runs, and the ncu launch list of the bench command.
mkdir -p gpurun_out
MODELS=${MODELS:-"vit_base_patch16_224 convnext_base"}
echo "=== pytest gpu" | tee gpurun_out/tests.log
timeout 1200 python -m pytest tests -m gpu -x -q -s ${PYTEST_ARGS} 2>&1 | grep -v "^$" | tail -60 | tee -a gpurun_out/tests.log
echo "=== smoke" | tee -a gpurun_out/tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee -a gpurun_out/tests.log
for MODEL in $MODELS; do
  echo "=== bench $MODEL"
  timeout 900 python bench.py --model $MODEL --steps 20 --warmup 5 ${BENCH_ARGS} > gpurun_out/bench_$MODEL.json 2> gpurun_out/bench_$MODEL.err
  tail -3 gpurun_out/bench_$MODEL.err; cut -c1-400 gpurun_out/bench_$MODEL.json
  if [ -n "$NCU_LIST" ]; then
    timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 150 --csv \
       --log-file gpurun_out/launches_$MODEL.csv python bench.py --model $MODEL --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  fi
done
