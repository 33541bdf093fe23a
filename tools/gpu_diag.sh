#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_reid.py -q -m gpu --tb=short 2>&1 | tail -6
