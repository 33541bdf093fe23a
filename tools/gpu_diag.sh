#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -q -m gpu --tb=line -k first 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_reid.py tests/test_gpu_detector.py tests/test_gpu_pipeline.py tests/test_gpu_ingest.py -q -m gpu --tb=short 2>&1 | tail -8
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --steps 256 --warmup 16 --no-cpu-baseline --dump-ops gpurun_out/ops.csv > gpurun_out/bench_$name.json 2> gpurun_out/plan.txt
  python -c "
import json,sys; d=json.load(open('gpurun_out/bench_$name.json')); print('$name', d['value'], d['ms_per_step'], d['e2e']['value'], d['stage_ms'], d['roofline']['achieved'])" || tail -3 gpurun_out/plan.txt
}
run default YDST_DEBUG_PLAN=1
