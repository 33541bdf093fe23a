#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipeline.py -q -m gpu --tb=short 2>&1 | tail -8
for mb in 4 8; do
timeout 600 python bench.py --steps 256 --warmup 16 --no-cpu-baseline --micro-batch $mb > gpurun_out/bench_mb$mb.json 2> gpurun_out/bench_mb$mb.err
python -c "
import json; d=json.load(open('gpurun_out/bench_mb$mb.json')); print('mb$mb', d['value'], d['ms_per_step'], d['e2e']['value'], d['stage_ms'], d['roofline']['achieved'], d['config']['dets_per_frame'], d['config']['track_rows_per_frame'])" || tail -5 gpurun_out/bench_mb$mb.err
done
