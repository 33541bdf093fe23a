#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_detector.py tests/test_gpu_ingest.py tests/test_gpu_pipeline.py -q -m gpu --tb=short 2>&1 | tail -6
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --steps 256 --warmup 16 --no-cpu-baseline > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/bench_$name.json')); print('$name', d['value'], d['ms_per_step'], d['e2e']['value'], d['stage_ms'], d['roofline']['achieved'])" || tail -3 gpurun_out/bench_$name.err
}
run default YDST_GRAPH=1
