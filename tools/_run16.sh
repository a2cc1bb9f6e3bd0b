cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
C3POA_GRP_GROW_PCT=15 C3POA_GRP_BUDGET_PCT=78 python tools/grp_ncu_run.py 100000 > gpurun_out/r2_run16a.txt 2>&1; cat gpurun_out/r2_run16a.txt
C3POA_GRP_GROW_PCT=15 C3POA_GRP_BUDGET_PCT=78 timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__thread_inst_executed_per_inst_executed.ratio,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:c3_poa_gr --csv --log-file gpurun_out/r2_kern16.csv python tools/grp_ncu_run.py 100000 > gpurun_out/r2_run16.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t16.log 2>&1; tail -5 gpurun_out/r2_t16.log
