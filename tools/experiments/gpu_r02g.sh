#!/bin/bash
out=gpurun_out; mkdir -p $out
python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "small_batch" > $out/r02g_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $out/r02g_pytest.log
for sb in 0 65536 262144 1048576 4194304 1073741824; do RTX_OPTS=simple_below=$sb python tools/gpu_perf.py cornell-lucy 16 2>&1 | tail -1 | cut -c1-230; done
for sb in 0 262144 1048576; do RTX_OPTS=simple_below=$sb python tools/gpu_perf.py cornell-lucy 64 2>&1 | tail -1 | cut -c1-230; RTX_OPTS=simple_below=$sb python tools/gpu_perf.py random 64 2>&1 | tail -1 | cut -c1-200; done
