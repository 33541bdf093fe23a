mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv.py -x -q 2>&1 | tail -5 > gpurun_out/r2k_conv.log; cat gpurun_out/r2k_conv.log
timeout 800 python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/r2k_tests.log; tail -3 gpurun_out/r2k_tests.log
timeout 300 python bench.py --steps 64 --warmup 16 --no-cpu-baseline --no-api --no-b1 --dump-ops gpurun_out/r2k_ops.csv > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
YDST_DEBUG_PLAN=1 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-api --no-b1 2>&1 >/dev/null | grep "^conv_plan" | awk '!seen[$0]++' > gpurun_out/r2k_plan.txt
for P in 0 1; do
YDST_PERSISTENT=$P YDST_CONV_TRACE=1 timeout 120 python tools/conv_probe_one.py 408 64 32 64 64 3 0 2>&1 | grep -A1 conv_trace | tail -4 > gpurun_out/r2k_trace_l1_p$P.txt
YDST_PERSISTENT=$P YDST_CONV_TRACE=1 timeout 120 python tools/conv_probe_one.py 8 76 76 256 128 1 0 2>&1 | grep -A1 conv_trace | tail -4 > gpurun_out/r2k_trace_1x1_p$P.txt
done
cat gpurun_out/r2k_trace_*.txt
cut -c1-400 gpurun_out/r2k_bench.json
