#!/bin/bash
mkdir -p gpurun_out
YDST_DEBUG_PLAN=1 timeout 600 python bench.py --steps 64 --warmup 8 --no-cpu-baseline --micro-batch 4 --dump-ops gpurun_out/ops.csv > gpurun_out/bench_diag.json 2> gpurun_out/plan.txt
python -c "
import json; d=json.load(open('gpurun_out/bench_diag.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['stage_ms'], d['roofline'])"
