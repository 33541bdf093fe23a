mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for C in 0 1; do
YDST_CTA2=$C timeout 200 python bench.py --steps 64 --warmup 16 --no-cpu-baseline --no-api --no-b1 --dump-ops gpurun_out/r2u_ops_c$C.csv > gpurun_out/r2u_bench_c$C.json 2> gpurun_out/r2u_bench_c$C.err
cut -c1-200 gpurun_out/r2u_bench_c$C.json; tail -3 gpurun_out/r2u_bench_c$C.err
done
YDST_CTA2=1 YDST_DEBUG_PLAN=1 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-api --no-b1 2>&1 >/dev/null | grep "^conv_plan" | awk '!seen[$0]++' > gpurun_out/r2u_plan_c1.txt
