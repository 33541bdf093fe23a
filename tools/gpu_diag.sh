#!/bin/bash
mkdir -p gpurun_out
YDST_DEBUG_PLAN=1 timeout 900 python -m pytest tests/test_gpu_conv.py -q -m gpu --tb=line -s 2>&1 | grep -v "^  tiling" | grep "res 1\|passed\|failed\|Error" | tail -12
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --steps 256 --warmup 16 --no-cpu-baseline > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/bench_$name.json')); print('$name', d['value'], d['ms_per_step'], d['e2e']['value'], d['stage_ms'], d['roofline']['achieved'])" || tail -3 gpurun_out/bench_$name.err
}
run default YDST_GRAPH=1
run pers_res YDST_PERSISTENT=1
run pers_nores YDST_PERSISTENT=1 YDST_B_RESIDENT=0
