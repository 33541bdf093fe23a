#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dropin.py tests/test_gpu_pipeline.py -q -m gpu --tb=short 2>&1 | tail -12
