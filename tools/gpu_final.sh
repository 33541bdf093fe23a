#!/bin/bash
# Final artefacts of a round: GPU tests (-s: measured errors in the log), both bench arms as the driver runs them, a longer run,
# per-op table, planner choices, ncu launch list, side configs.   usage: bash tools/gpu_final.sh <tag>
tag=${1:-r02_final}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -q -m gpu -s --timeout 120 2>&1 | grep -av "^conv_plan\|^  tiling" | tail -120 > gpurun_out/gpu_tests_$tag.log; tail -3 gpurun_out/gpu_tests_$tag.log | cut -c1-200
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 >> gpurun_out/gpu_tests_$tag.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_reference_$tag.json 2> gpurun_out/bench_reference_$tag.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
timeout 600 python bench.py --steps 256 --warmup 16 --no-cpu-baseline --no-api --dump-ops gpurun_out/ops_$tag.csv > gpurun_out/bench_256_$tag.json 2>> gpurun_out/bench_$tag.err
YDST_DEBUG_PLAN=1 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-api --no-b1 2>&1 >/dev/null | grep "^conv_plan" | awk '!seen[$0]++' > gpurun_out/conv_plan_$tag.txt
YDST_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 600 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 16 --warmup 8 --no-cpu-baseline --no-api --no-b1 > gpurun_out/launches_run_$tag.log 2>&1
timeout 400 python bench.py --config yolov4 --steps 64 --warmup 16 --no-api > gpurun_out/bench_yolov4_$tag.json 2> gpurun_out/bench_yolov4_$tag.err
timeout 400 python bench.py --config reid > gpurun_out/bench_reid_$tag.json 2> gpurun_out/bench_reid_$tag.err
timeout 400 python bench.py --config assoc > gpurun_out/bench_assoc_$tag.json 2> gpurun_out/bench_assoc_$tag.err
for f in bench_$tag bench_256_$tag bench_reference_$tag bench_yolov4_$tag bench_reid_$tag bench_assoc_$tag; do echo "== $f"; grep -a "^{" gpurun_out/$f.json | cut -c1-260; done
