#!/bin/bash
# ncu pass over the convolutions of one frame (limited sections -> small report), exported to CSV on the box.
mkdir -p gpurun_out
timeout 1500 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section LaunchStats --section Occupancy \
    --clock-control none -k regex:conv_tc -s 279 -c 93 -o gpurun_out/prof_conv_sections -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_run.log 2>&1
ncu -i gpurun_out/prof_conv_sections.ncu-rep --page raw --csv > gpurun_out/prof_conv_sections.csv 2> /dev/null
ls -la gpurun_out/
sz=$(stat -c %s gpurun_out/prof_conv_sections.ncu-rep); if [ "$sz" -gt 40000000 ]; then rm gpurun_out/prof_conv_sections.ncu-rep; fi
tail -3 gpurun_out/ncu_run.log | cut -c1-200
