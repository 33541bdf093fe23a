#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_reid.py tests/test_gpu_pipeline.py -x -q 2>&1 | tail -3
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_bucket.json 2> gpurun_out/bench_bucket.err; tail -2 gpurun_out/bench_bucket.err
python -c "
import json; d=json.load(open('gpurun_out/bench_bucket.json')); print(d['value'], d['e2e']['value'], d['stage_ms'])"
timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 3 | cut -c1-160
