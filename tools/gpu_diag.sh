#!/bin/bash
mkdir -p gpurun_out
YDST_CONV_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_trace.json 2> gpurun_out/trace.txt
grep conv_trace gpurun_out/trace.txt | tail -186 | head -93 > gpurun_out/trace_frame.txt
wc -l gpurun_out/trace_frame.txt
