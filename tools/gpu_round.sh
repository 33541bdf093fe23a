#!/bin/bash
# One GPU visit: parity tests, bench line (both arms), ncu launch list of a bench run, ncu --set full of representative convs.
# usage: bash tools/gpu_round.sh [tests] [bench] [launches] [full] [artefacts]        (outputs under gpurun_out/, kept < 64 MiB)
mkdir -p gpurun_out
what="${@:-tests bench launches full}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for w in $what; do
case $w in
tests)
  timeout 1500 python -m pytest tests -q -m gpu --tb=short -s 2>&1 | tail -80 > gpurun_out/gpu_tests.log
  tail -12 gpurun_out/gpu_tests.log
  timeout 600 python __graft_entry__.py smoke 2>&1 | tail -3 ;;
bench)
  timeout 900 python bench.py --impl reference --steps 12 --warmup 2 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
  timeout 900 python bench.py --steps 256 --warmup 16 > gpurun_out/bench.json 2> gpurun_out/bench.err
  tail -3 gpurun_out/bench.err; cut -c1-300 gpurun_out/bench_reference.json; cat gpurun_out/bench.json ;;
launches)
  # graphs off so that every kernel is a separate stream launch in the list (same kernels, same order)
  YDST_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 500 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/launches_run.log 2>&1
  tail -1 gpurun_out/launches_run.log | cut -c1-200 ;;
full)
  # (micro-batch 1 so that launch indices address single-frame kernels) the four representative convolutions of a yolov3-608 frame: 1x1 256->128 @76, 3x3 128->256 @76, 1x1 1024->512 @19, 3x3 512->1024 @19
  YDST_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 288 -c 2 -o gpurun_out/prof_full_76 -f \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --micro-batch 1 > gpurun_out/ncu_full_a.log 2>&1
  YDST_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 322 -c 2 -o gpurun_out/prof_full_19 -f \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --micro-batch 1 > gpurun_out/ncu_full_b.log 2>&1
  # the same 3x3 128->256 @76x76 layer at the bench's default micro-batch (what roofline.traffic quotes)
  YDST_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 289 -c 1 -o gpurun_out/prof_full_mb -f \
      python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_c.log 2>&1
  for f in 76 19 mb; do ncu -i gpurun_out/prof_full_$f.ncu-rep --page raw --csv > gpurun_out/prof_full_$f.csv 2>/dev/null; done
  ls -la gpurun_out/*.ncu-rep ;;
artefacts)
  # per-op CUDA-event table, the planner's choices and the in-kernel timelines that tools/summarise_profiles.py files under profiles/
  timeout 600 python bench.py --steps 64 --warmup 16 --no-cpu-baseline --dump-ops gpurun_out/ops.csv > /dev/null 2> gpurun_out/ops.err
  YDST_DEBUG_PLAN=1 timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "^conv_plan" | awk '!seen[$0]++' > gpurun_out/plan.txt
  YDST_CONV_TRACE=2 timeout 600 python bench.py --steps 16 --warmup 3 --no-cpu-baseline 2> gpurun_out/timeline.txt > /dev/null
  grep "^conv_timeline" gpurun_out/timeline.txt | tail -130 > gpurun_out/timeline_tail.txt
  wc -l gpurun_out/plan.txt gpurun_out/timeline_tail.txt gpurun_out/ops.csv ;;
esac
done
