mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 200 python -m pytest tests/test_gpu_conv.py -x -q --timeout 30 > gpurun_out/r2r_conv_all.log 2>&1; tail -3 gpurun_out/r2r_conv_all.log | cut -c1-400
for sh in "8 304 304 64 32 1 0" "408 64 32 64 64 3 2"; do
echo "== shape $sh"
YDST_CONV_TRACE=1 timeout 60 python tools/conv_probe_one.py $sh 2>&1 | grep -A12 "conv_trace" | tail -13 | cut -c1-300
done > gpurun_out/r2r_traces.txt 2>&1
cat gpurun_out/r2r_traces.txt
timeout 200 python bench.py --steps 64 --warmup 16 --no-cpu-baseline --no-api --no-b1 --dump-ops gpurun_out/r2r_ops.csv > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err
cut -c1-200 gpurun_out/r2r_bench.json; tail -3 gpurun_out/r2r_bench.err
