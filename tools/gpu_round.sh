#!/bin/bash
timeout 300 python tools/gpu_check.py cornell-lucy cornell 2>&1 | grep -E "trace|render|secondary|Error|error" 
for s in 64; do timeout 300 python tools/gpu_perf.py cornell-lucy $s 2>&1 | tail -1; done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
