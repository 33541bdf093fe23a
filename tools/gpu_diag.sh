#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -x -q 2>&1 | tail -4 > gpurun_out/conv_tests.txt
cat gpurun_out/conv_tests.txt
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --steps 256 --warmup 16 --no-cpu-baseline > gpurun_out/bench_$name.json 2> gpurun_out/plan_$name.txt
  python -c "
import json,sys; d=json.load(open('gpurun_out/bench_$name.json')); print('$name', d['value'], d['ms_per_step'], d['e2e']['value'], d['stage_ms'], d['roofline']['achieved'])" || tail -3 gpurun_out/plan_$name.txt
}
run v3 YDST_X=0
run v3_noearly YDST_BO_MODE=2
timeout 300 python tools/conv_probe.py 2> gpurun_out/probe_v3.txt
