#!/bin/bash
out=gpurun_out; mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/r02n_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $out/r02n_pytest_gpu.log
RT_DEBUG_TIMING=1 RTX_DEBUG_BATCH=1 python tools/gpu_perf.py cornell-lucy 8 2>&1 | grep -E "LoadOBJ|scene upload|ParseOBJ" | head
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $out/r02n_bench.json 2> $out/r02n_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('$out/r02n_bench.json')); print(d['value'], d['e2e']['value'], d['e2e_cold']); print(d['e2e']['breakdown_ms'])"
