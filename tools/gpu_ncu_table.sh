#!/bin/bash
# ncu --set full of every kernel of one steady-state micro-batch (tools/ncu_forward.py), summarised by tools/ncu_table.py.
# usage: bash tools/gpu_ncu_table.sh <tag>          (outputs gpurun_out/ncu_full_<tag>.csv / .md; the .ncu-rep stays on the box)
tag=${1:-r02}
mkdir -p gpurun_out
YDST_GRAPH=0 timeout 1500 ncu --set full --clock-control none --profile-from-start off -o /tmp/ncu_full_$tag -f \
    python tools/ncu_forward.py yolov3 8 > gpurun_out/ncu_full_$tag.log 2>&1
ncu -i /tmp/ncu_full_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_full_$tag.csv 2>/dev/null
python tools/ncu_table.py gpurun_out/ncu_full_$tag.csv gpurun_out/ncu_full_$tag.md \
    "ncu --set full of EVERY kernel of one steady-state micro-batch (8 frames of yolov3-608, ~408 ReID crops, 8 DeepSort updates): YDST_GRAPH=0 ncu --set full --clock-control none --profile-from-start off python tools/ncu_forward.py yolov3 8"
ls -la gpurun_out/ncu_full_$tag.* /tmp/ncu_full_$tag.ncu-rep; tail -2 gpurun_out/ncu_full_$tag.log
