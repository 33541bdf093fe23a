#!/bin/bash
# ncu --set full with source correlation on four representative conv launches of one frame (batch-1 yolov3-608):
#   A: 1x1 256->128 @76 and 3x3 128->256 @76     B: 1x1 1024->512 @19 and 3x3 512->1024 @19
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 288 -c 2 -o gpurun_out/prof_full_76 -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_a.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 322 -c 2 -o gpurun_out/prof_full_19 -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_b.log 2>&1
ls -la gpurun_out/*.ncu-rep
