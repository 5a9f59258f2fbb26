# Developer tool (GPU box): the round-2 launch list, ncu --set full capture of all four bounce levels, and the local-memory
# counters of the traversal stack (default build vs the shared-memory stack variant) behind profiles/r02_*.
set -x
python tools/profile_frame.py 3 > gpurun_out/r02_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python tools/profile_frame.py 3 > gpurun_out/r02_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_extend|k_shade|k_shadow_point' --launch-skip 12 --launch-count 12 -f -o gpurun_out/r02_prof python tools/profile_frame.py 3 > gpurun_out/r02_ncu_full.log 2>&1
M=smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sector_hit_rate.pct,smsp__inst_executed.sum,gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
ncu --metrics $M --clock-control none -k regex:'k_extend|k_shadow_point' --launch-skip 8 --launch-count 8 --csv --log-file gpurun_out/r02_stack_local.csv python tools/profile_frame.py 3 > gpurun_out/r02_stack_local.log 2>&1
if [ -f variants/lib_smem16.so ]; then
RTB200_LIB=$PWD/variants/lib_smem16.so ncu --metrics $M --clock-control none -k regex:'k_extend|k_shadow_point' --launch-skip 8 --launch-count 8 --csv --log-file gpurun_out/r02_stack_smem16.csv python tools/profile_frame.py 3 > gpurun_out/r02_stack_smem16.log 2>&1
fi
tail -1 gpurun_out/r02_plain.log; tail -2 gpurun_out/r02_ncu_full.log
