out=gpurun_out; mkdir -p $out
P() { tail -1 | sed 's/ nodes\/ray.*//' | cut -c1-190; }
python tools/gpu_perf.py cornell-lucy 64 2>&1 | P
RTX_OPTS=pretest_bare=8 python tools/gpu_perf.py cornell-lucy 64 2>&1 | P
for v in n2 n4 ts2 ts6; do RTX_B200_LIB=$PWD/build/ab/librtx_$v.so python tools/gpu_perf.py cornell-lucy 64 2>&1 | P; done
RTX_OPTS=pool_paths=8388608 python tools/gpu_perf.py cornell-lucy 64 > $out/r01p_plain_lucy.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_extend -s 10 -c 1 -f -o $out/r01p_prof_extend \
    env RTX_OPTS=pool_paths=8388608 python tools/gpu_perf.py cornell-lucy 64 > $out/r01p_ncu_extend.log 2>&1
tail -2 $out/r01p_ncu_extend.log; tail -1 $out/r01p_plain_lucy.log | P
