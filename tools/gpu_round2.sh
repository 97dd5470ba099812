#!/bin/bash
# One GPU-box round of evidence for round 2: parity tests, smoke, both bench arms, the ncu launch list, the three ncu --set full captures
# (each after the same command exited 0 without ncu) and their JSON summaries. Usage (under gpurun): bash tools/gpu_round2.sh <tag> [nopytest]
tag=${1:-r02}; out=gpurun_out; mkdir -p $out
if [ "$2" != "nopytest" ]; then
python -m pytest tests -m gpu -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest_gpu.log; tail -3 $out/${tag}_pytest_gpu.log
fi
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $out/${tag}_smoke.log
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"; cut -c1-400 $out/${tag}_bench.json
python bench.py --impl reference --steps 1 --warmup 1 > $out/${tag}_bench_ref.json 2>> $out/${tag}_bench.err; cut -c1-300 $out/${tag}_bench_ref.json
for wl in cornell random cornell-glossy hdri-test; do python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_$wl.json 2>> $out/${tag}_bench.err; cut -c1-120 $out/${tag}_bench_$wl.json; done
python bench.py --spp 4 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --spp 4 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $out/${tag}_ncu_l.log 2>&1
bash tools/gpu_ncu.sh $tag > $out/${tag}_ncu_run.log 2>&1; tail -3 $out/${tag}_ncu_run.log
rl=$(grep -o "extension rays [0-9]*" $out/${tag}_rays_lucy.txt | grep -o "[0-9]*$"); rs=$(grep -o "shadow rays [0-9]*" $out/${tag}_rays_lucy.txt | grep -o "[0-9]*$"); rh=$(grep -o "extension rays [0-9]*" $out/${tag}_rays_hdri.txt | grep -o "[0-9]*$")
python tools/ncu_to_json.py $out/${tag}_prof_extend.ncu-rep k_extend $rl "ncu --set full --clock-control none --import-source on -k regex:k_extend -s 10 -c 1 env RTX_OPTS=pool_paths=8388608 python tools/gpu_perf.py cornell-lucy 64" > $out/${tag}_k_extend.json
python tools/ncu_to_json.py $out/${tag}_prof_connect.ncu-rep k_connect $rs "ncu --set full --clock-control none --import-source on -k regex:k_connect -s 10 -c 1 env RTX_OPTS=pool_paths=8388608 python tools/gpu_perf.py cornell-lucy 64" > $out/${tag}_k_connect.json
python tools/ncu_to_json.py $out/${tag}_prof_bounce.ncu-rep k_bounce_flat $rh "ncu --set full --clock-control none --import-source on -k regex:k_bounce_flat -s 6 -c 1 python tools/gpu_perf.py hdri-test 16" > $out/${tag}_k_bounce_flat.json
python tools/ncu_summary.py $out/${tag}_prof_extend.ncu-rep 120 > $out/${tag}_k_extend_sass.txt 2>&1
python tools/ncu_summary.py $out/${tag}_prof_connect.ncu-rep 120 > $out/${tag}_k_connect_sass.txt 2>&1
for s in cornell random cornell-glossy cornell-lucy hdri-test quads earth cornell-smoke checkered simple glossy-metal perlin primitives; do python tools/gpu_perf.py $s 64 2>&1 | tail -1 | cut -c1-230; done > $out/${tag}_scenes.log; cat $out/${tag}_scenes.log | cut -c1-120
# the merge back is limited to 64 MiB: keep the k_extend report, drop the other two (their JSON / SASS summaries stay)
rm -f $out/${tag}_prof_connect.ncu-rep $out/${tag}_prof_bounce.ncu-rep
du -sh $out | tail -1
