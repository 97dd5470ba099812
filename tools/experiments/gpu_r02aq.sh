#!/bin/bash
out=gpurun_out; mkdir -p $out
for s in "hdri-test 64" "cornell-lucy 64" "cornell-glossy 256" "random 64"; do timeout 60 python tools/gpu_perf.py $s 2>&1 | tail -1 | cut -c1-175; done
timeout 1500 python -m pytest tests -m gpu -q -x > $out/r02aq_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $out/r02aq_pytest.log
