#!/bin/bash
# One GPU visit: parity tests, bench line (both arms), ncu launch list of a bench run, ncu --set full of one frame's convs.
# usage: bash tools/gpu_round.sh [tests] [bench] [launches] [full]
mkdir -p gpurun_out
what="${@:-tests bench launches full}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for w in $what; do
case $w in
tests)
  timeout 1200 python -m pytest tests -q -m gpu --tb=short -s 2>&1 | tail -70 > gpurun_out/gpu_tests.log
  tail -25 gpurun_out/gpu_tests.log ;;
bench)
  timeout 900 python bench.py --impl reference --steps 12 --warmup 2 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
  timeout 900 python bench.py --steps 256 --warmup 16 > gpurun_out/bench.json 2> gpurun_out/bench.err
  tail -3 gpurun_out/bench.err; cat gpurun_out/bench_reference.json gpurun_out/bench.json ;;
launches)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 480 -c 400 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/launches_run.log 2>&1
  tail -2 gpurun_out/launches_run.log | cut -c1-300 ;;
full)
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 279 -c 93 -o gpurun_out/prof_conv -f \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/full_run.log 2>&1
  tail -2 gpurun_out/full_run.log | cut -c1-300; ls -la gpurun_out/*.ncu-rep ;;
esac
done
