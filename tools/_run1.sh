out=gpurun_out; mkdir -p $out
RTX_OPTS=pool_paths=8388608 python tools/gpu_perf.py cornell-lucy 64 > $out/r01m_plain_lucy.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_connect -s 10 -c 1 -f -o $out/r01m_prof_connect \
    env RTX_OPTS=pool_paths=8388608 python tools/gpu_perf.py cornell-lucy 64 > $out/r01m_ncu_connect.log 2>&1
tail -2 $out/r01m_ncu_connect.log
python tools/gpu_perf.py hdri-test 16 > $out/r01m_plain_hdri.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_bounce_flat -s 6 -c 1 -f -o $out/r01m_prof_bounce \
    python tools/gpu_perf.py hdri-test 16 > $out/r01m_ncu_bounce.log 2>&1
tail -2 $out/r01m_ncu_bounce.log
ls -la $out/*.ncu-rep
