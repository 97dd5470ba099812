#!/bin/bash
# Final validation of a round: full GPU suite, smoke, bench lines, scene table, recapture of the kernels that changed last.
tag=${1:-r02z}; out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest_gpu.log; tail -3 $out/${tag}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $out/${tag}_smoke.log | cut -c1-200
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_bounce_flat -s 6 -c 1 -f -o $out/${tag}_prof_bounce python tools/gpu_perf.py hdri-test 16 > $out/${tag}_ncu_bounce.log 2>&1
python tools/ncu_to_json.py $out/${tag}_prof_bounce.ncu-rep k_bounce_flat 16777216 "ncu --set full --clock-control none --import-source on -k regex:k_bounce_flat -s 6 -c 1 python tools/gpu_perf.py hdri-test 16" > $out/${tag}_k_bounce_flat.json
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 10 -c 1 -f -o $out/${tag}_prof_shade env RTX_OPTS=pool_paths=8388608 python tools/gpu_perf.py cornell-lucy 64 > $out/${tag}_ncu_shade.log 2>&1
python tools/ncu_to_json.py $out/${tag}_prof_shade.ncu-rep k_shade 8388608 "ncu --set full --clock-control none --import-source on -k regex:k_shade -s 10 -c 1 env RTX_OPTS=pool_paths=8388608 python tools/gpu_perf.py cornell-lucy 64" > $out/${tag}_k_shade.json
python tools/ncu_summary.py $out/${tag}_prof_shade.ncu-rep 120 > $out/${tag}_k_shade_sass.txt 2>&1
rm -f $out/${tag}_prof_bounce.ncu-rep $out/${tag}_prof_shade.ncu-rep
cp $out/${tag}_k_bounce_flat.json profiles/r02_k_bounce_flat.json   # bench.py reads its ncu keys from profiles/
bash tools/gpu_bench_only.sh $tag
for s in cornell random cornell-glossy cornell-lucy hdri-test quads earth cornell-smoke checkered simple glossy-metal perlin primitives; do timeout 100 python tools/gpu_perf.py $s 64 2>&1 | tail -1 | cut -c1-230; done > $out/${tag}_scenes.log; cut -c1-110 $out/${tag}_scenes.log
