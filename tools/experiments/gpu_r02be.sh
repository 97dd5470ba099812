#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 300 python -m pytest tests -m gpu -x -q -k "full_size or lucy_at_depth or sample_slices or test_order or checked or golden or level1_configured" > $out/r02be_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/r02be_pytest.log
bash tools/gpu_bench_only.sh r02y 2>&1 | cut -c1-160
