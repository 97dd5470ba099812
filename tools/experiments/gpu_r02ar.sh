#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 1700 python -m pytest tests -m gpu -q > $out/r02ar_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $out/r02ar_pytest.log
for s in "cornell-lucy 16" "random 64"; do timeout 60 python tools/gpu_perf.py $s 2>&1 | tail -1 | cut -c1-175; done
