mkdir -p gpurun_out
timeout 200 python bench.py --steps 64 --warmup 16 --no-cpu-baseline --no-api --no-b1 > gpurun_out/r3h_bench.json 2> gpurun_out/r3h_bench.err
python - gpurun_out/r3h_bench.json <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[1], d['value'], d['e2e']['value'], d['roofline']['achieved'], d['stage_ms'], d['config']['latency_frames'])
PY
tail -2 gpurun_out/r3h_bench.err
timeout 300 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_e2e.py -q -x --timeout 200 2>&1 | tail -2
