# Round 2: A/B of leaf-brick staging in shared memory: default (LDG.E.U8.CONSTANT through L1) vs
# stage1 (-DWX_STAGE_LEAF=1: cooperative LDG.128 + STS.128) vs stage2 (-DWX_STAGE_LEAF=2: cp.async.bulk / UBLKCP + mbarrier).
mkdir -p gpurun_out
bash tools/gpu/ab.sh default stage1 stage2 > /dev/null 2>&1
cp gpurun_out/ab.txt gpurun_out/r2_stage_ab.txt
M=gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__t_sector_hit_rate.pct,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,lts__t_sectors.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
for n in default stage1 stage2; do
  if [ "$n" = default ]; then unset WOXEL_B200_LIB; else export WOXEL_B200_LIB=$PWD/woxel_b200/libwoxel_b200_$n.so; fi
  echo "== $n" >> gpurun_out/r2_stage_ab.txt
  WX_LONG_FIRST=0 timeout 120 ncu --metrics $M --clock-control none -k regex:raycast_kernel -s 2 -c 1 --csv python tools/prof_run.py 2>/dev/null | grep -E "raycast_kernel" | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' >> gpurun_out/r2_stage_ab.txt
done
cat gpurun_out/r2_stage_ab.txt
