# Developer tool (GPU box): the end-of-round bench line, launch list and ncu --set full capture behind profiles/r01g_*.
set -x
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r01g.json 2> gpurun_out/bench_r01g.err
python tools/profile_frame.py 3 > gpurun_out/r01g_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01g_launches.csv python tools/profile_frame.py 3 > gpurun_out/r01g_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_extend|k_shade|k_shadow_point' --launch-skip 12 --launch-count 12 -f -o gpurun_out/r01g_prof python tools/profile_frame.py 3 > gpurun_out/r01g_ncu_full.log 2>&1
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_r01g_ref.json 2> gpurun_out/bench_r01g_ref.err
cat gpurun_out/bench_r01g.json; tail -2 gpurun_out/r01g_plain.log; tail -2 gpurun_out/r01g_ncu_full.log; cat gpurun_out/bench_r01g_ref.json
