mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -q -m gpu --timeout 120 > gpurun_out/r2z_tests.log 2>&1; tail -5 gpurun_out/r2z_tests.log | cut -c1-300
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 200 python bench.py --steps 64 --warmup 16 --no-cpu-baseline --no-api --no-b1 --dump-ops gpurun_out/r2z_ops.csv > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err
cut -c1-200 gpurun_out/r2z_bench.json; tail -3 gpurun_out/r2z_bench.err
YDST_DEBUG_PLAN=1 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-api --no-b1 2>&1 >/dev/null | grep "^conv_plan" | awk '!seen[$0]++' > gpurun_out/r2z_plan.txt
