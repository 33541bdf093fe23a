mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_conv.py tests/test_gpu_reid.py -q -x --timeout 120 2>&1 | tail -2 | cut -c1-300
for K in 20 64 256; do
timeout 200 python bench.py --steps $K --warmup 5 --no-cpu-baseline --no-api --no-b1 > gpurun_out/r3d_bench_$K.json 2> gpurun_out/r3d_bench_$K.err
python - gpurun_out/r3d_bench_$K.json <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[1], d['value'], d['e2e']['value'], d['roofline']['achieved'], d['stage_ms'], d['windows'])
PY
tail -2 gpurun_out/r3d_bench_$K.err
done
