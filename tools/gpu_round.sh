#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python tools/gpu_check.py > gpurun_out/check.log 2>&1; tail -30 gpurun_out/check.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
