#!/bin/bash
# First-contact GPU run: every test file in its own process (a faulting kernel poisons the CUDA context), logs kept.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for f in test_gpu_conv test_gpu_assoc test_gpu_reid test_gpu_detector test_gpu_pipeline; do
  echo "=== $f" | tee -a gpurun_out/first.log
  timeout 900 python -m pytest tests/$f.py -q -m gpu -x -s --tb=short > gpurun_out/$f.log 2>&1
  echo "exit $?" | tee -a gpurun_out/first.log
  tail -n 40 gpurun_out/$f.log >> gpurun_out/first.log
done
if grep -q "failed\|error" gpurun_out/test_gpu_conv.log; then
  for k in 1x1_s1_19 3x3_s1_38 3x3_s1_bk32 3x3_s1_bk16 3x3_s2_76 1x1_s2_batch3 head_255_f32 res_after_act deep_k_wide_n first_s1; do
    echo "--- conv case $k" >> gpurun_out/first.log
    timeout 300 python -m pytest tests/test_gpu_conv.py -q -m gpu -k "$k" --tb=line 2>&1 | tail -n 6 >> gpurun_out/first.log
  done
fi
tail -n 150 gpurun_out/first.log
