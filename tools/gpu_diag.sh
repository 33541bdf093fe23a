#!/bin/bash
# final artefacts of a round: bench line (with per-op dump), planner choices, in-kernel timeline
mkdir -p gpurun_out
timeout 900 python bench.py --dump-ops gpurun_out/ops.csv > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -2 gpurun_out/bench.err; cut -c1-200 gpurun_out/bench.json
YDST_DEBUG_PLAN=1 timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "^conv_plan" | awk '!seen[$0]++' > gpurun_out/plan.txt
YDST_CONV_TRACE=2 timeout 600 python bench.py --steps 16 --warmup 3 --no-cpu-baseline 2> gpurun_out/timeline.txt > /dev/null
grep "^conv_timeline" gpurun_out/timeline.txt | tail -130 > gpurun_out/timeline_tail.txt
wc -l gpurun_out/plan.txt gpurun_out/timeline_tail.txt gpurun_out/ops.csv
