#!/bin/bash
for s in 16 64; do python tools/gpu_perf.py cornell-lucy $s 2>&1 | tail -1; done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
