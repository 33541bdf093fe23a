#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --steps 256 --warmup 16 --no-cpu-baseline --dump-ops gpurun_out/ops_$name.csv > gpurun_out/bench_$name.json 2> gpurun_out/plan_$name.txt
  python -c "
import json,sys; d=json.load(open('gpurun_out/bench_$name.json')); print('$name', d['value'], d['ms_per_step'], d['e2e']['value'], d['stage_ms'], d['roofline']['achieved'])" || tail -3 gpurun_out/plan_$name.txt
}
run bn256_96 YDST_DEBUG_PLAN=1
run bn256_96_tpb1 YDST_TPB1_BN128=1 YDST_DEBUG_PLAN=1
run bn256_48 YDST_BN256_MIN_CTAS=48 YDST_DEBUG_PLAN=1
