#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_reid.py -q -m gpu --tb=short 2>&1 | tail -30 > gpurun_out/diag_tests.log
tail -4 gpurun_out/diag_tests.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 64 --warmup 8 --no-cpu-baseline > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/bench_$name.json')); print('$name', d['value'], d['ms_per_step'], d['e2e']['value'], d['stage_ms'])" || tail -3 gpurun_out/bench_$name.err
}
run graph_pdl_200 YDST_GRAPH=1
run graph_pdl_100 YDST_GRAPH=1 YDST_SMEM_BUDGET_KB=100
run graph_nopdl_200 YDST_GRAPH=1 YDST_PDL=0
run nograph_pdl_200 YDST_GRAPH=0
run nograph_pdl_100 YDST_GRAPH=0 YDST_SMEM_BUDGET_KB=100
