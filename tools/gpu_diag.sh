#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -q -m gpu --tb=line -x -k persistent 2>&1 | tail -3
YDST_DEBUG_PLAN=1 timeout 600 python bench.py --steps 128 --warmup 8 --no-cpu-baseline --dump-ops gpurun_out/ops.csv > gpurun_out/bench_diag.json 2> gpurun_out/plan.txt
python -c "
import json; d=json.load(open('gpurun_out/bench_diag.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['stage_ms'], d['roofline']['achieved'])"
YDST_CONV_TRACE=2 YDST_GRAPH=0 timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tl.json 2> gpurun_out/timeline.txt
grep -A1 conv_timeline gpurun_out/timeline.txt | tail -400 > gpurun_out/timeline_tail.txt
