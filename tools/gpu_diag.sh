#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_detector.py tests/test_gpu_reid.py -x -q 2>&1 | tail -4
run() { name=$1; extra=$2; shift; shift
  env "$@" timeout 600 python bench.py --steps 256 --warmup 16 --no-cpu-baseline $extra --dump-ops gpurun_out/ops_$name.csv > gpurun_out/bench_$name.json 2> gpurun_out/plan_$name.txt
  python -c "
import json,sys; d=json.load(open('gpurun_out/bench_$name.json')); print('$name', d['value'], d['ms_per_step'], d['e2e']['value'], d['stage_ms'], d['roofline']['achieved'])" || tail -3 gpurun_out/plan_$name.txt
}
run decode "" YDST_X=0
