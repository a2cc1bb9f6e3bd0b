cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,sm__warps_active.avg.per_cycle_active,smsp__thread_inst_executed_per_inst_executed.ratio,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:c3_poa_grp --csv --log-file gpurun_out/r2_kern6.csv python tools/grp_ncu_run.py 37888 > gpurun_out/r2_run6.txt 2>&1
tail -2 gpurun_out/r2_run6.txt
