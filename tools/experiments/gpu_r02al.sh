#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "small_batch or level2 or full_size or sample_slices or image_textures" > $out/r02al_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/r02al_pytest.log
for fd in 0 1048576; do for spp in 16 64; do timeout 60 env RTX_OPTS=fuse_drain=$fd python tools/gpu_perf.py cornell-lucy $spp 2>&1 | tail -1 | cut -c1-170; done; timeout 40 env RTX_OPTS=fuse_drain=$fd python tools/gpu_perf.py random 64 2>&1 | tail -1 | cut -c1-150;  timeout 40 env RTX_OPTS=fuse_drain=$fd python tools/gpu_perf.py cornell-smoke 64 2>&1 | tail -1 | cut -c1-150; done
