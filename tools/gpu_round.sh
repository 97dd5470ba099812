#!/bin/bash
# One GPU-box round: parity tests, smoke, both bench arms, then the ncu launch list and one full capture of k_extend.
# Usage (under gpurun): bash tools/gpu_round.sh <tag>     outputs land in gpurun_out/<tag>_*
tag=${1:-rNN}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest_gpu.log
tail -3 $out/${tag}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $out/${tag}_smoke.log
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"; cat $out/${tag}_bench.json
python bench.py --impl reference --steps 1 --warmup 1 > $out/${tag}_bench_ref.json 2>> $out/${tag}_bench.err; cat $out/${tag}_bench_ref.json
if [ "$2" != "noncu" ]; then
python bench.py --spp 4 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --spp 4 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $out/${tag}_ncu_l.log 2>&1
# launch 10 of a 64-spp pass: the stream is still full (8,388,608 rays with pool_paths = 2^23), bounces are mixed
RTX_OPTS=pool_paths=8388608 python tools/gpu_perf.py cornell-lucy 64 > $out/${tag}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_extend -s 10 -c 1 -f -o $out/${tag}_prof_extend \
    env RTX_OPTS=pool_paths=8388608 python tools/gpu_perf.py cornell-lucy 64 > $out/${tag}_ncu.log 2>&1
tail -2 $out/${tag}_ncu.log
# the other two kernels that carry a BASELINE config: shadow rays of cornell-lucy, the one-kernel bounce of hdri-test
ncu --set full --clock-control none --import-source on -k regex:k_connect -s 10 -c 1 -f -o $out/${tag}_prof_connect \
    env RTX_OPTS=pool_paths=8388608 python tools/gpu_perf.py cornell-lucy 64 > $out/${tag}_ncu_connect.log 2>&1
python tools/gpu_perf.py hdri-test 16 > $out/${tag}_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_bounce_flat -s 6 -c 1 -f -o $out/${tag}_prof_bounce \
    python tools/gpu_perf.py hdri-test 16 > $out/${tag}_ncu_bounce.log 2>&1
fi
# per-scene throughput table (short passes) and pool-size A/B
for s in cornell random cornell-glossy cornell-lucy hdri-test; do python tools/gpu_perf.py $s 64 2>&1 | tail -1; done > $out/${tag}_scenes.log
for pp in 2097152 8388608; do RTX_OPTS=pool_paths=$pp python tools/gpu_perf.py cornell-lucy 64 2>&1 | tail -1; RTX_OPTS=pool_paths=$pp python tools/gpu_perf.py hdri-test 32 2>&1 | tail -1; done >> $out/${tag}_scenes.log
cat $out/${tag}_scenes.log
