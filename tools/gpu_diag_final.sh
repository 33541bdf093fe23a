#!/bin/bash
# final artefacts of a round: per-op dump, planner choices, in-kernel timeline (the bench line itself comes from gpu_round.sh)
mkdir -p gpurun_out
timeout 600 python bench.py --steps 64 --warmup 16 --no-cpu-baseline --dump-ops gpurun_out/ops.csv > /dev/null 2> gpurun_out/ops.err
YDST_DEBUG_PLAN=1 timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "^conv_plan" | awk '!seen[$0]++' > gpurun_out/plan.txt
YDST_CONV_TRACE=2 timeout 600 python bench.py --steps 16 --warmup 3 --no-cpu-baseline 2> gpurun_out/timeline.txt > /dev/null
grep "^conv_timeline" gpurun_out/timeline.txt | tail -130 > gpurun_out/timeline_tail.txt
wc -l gpurun_out/plan.txt gpurun_out/timeline_tail.txt gpurun_out/ops.csv
