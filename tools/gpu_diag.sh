#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --steps 256 --warmup 16 --no-cpu-baseline > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/bench_$name.json')); print('$name', d['value'], d['ms_per_step'], d['e2e']['value'], d['stage_ms'], d['roofline']['achieved'])" || tail -3 gpurun_out/bench_$name.err
}
run smem_model YDST_MODEL_SMEM=1
run old_model YDST_MODEL_SMEM=0
