mkdir -p gpurun_out
for T in 1 0; do
YDST_TAP_PERSISTENT=$T timeout 200 python bench.py --steps 64 --warmup 16 --no-cpu-baseline --no-api --no-b1 --dump-ops gpurun_out/r3g_ops_$T.csv > gpurun_out/r3g_bench_$T.json 2> gpurun_out/r3g_bench_$T.err
done
YDST_DEBUG_PLAN=1 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-api --no-b1 2>&1 >/dev/null | grep "^conv_plan k" | grep persistent | awk '!seen[$0]++' | cut -c1-220
