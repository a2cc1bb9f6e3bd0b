cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/grp_ncu_run.py 37888 > gpurun_out/r2_run11a.txt 2>&1; cat gpurun_out/r2_run11a.txt
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -k regex:c3_poa_grp --csv --log-file gpurun_out/r2_kern11.csv python tools/grp_ncu_run.py 37888 > gpurun_out/r2_run11.txt 2>&1
timeout 900 ncu --section SourceCounters --import-source on --clock-control none -k regex:c3_poa_grp_graph -s 2 -c 1 -o gpurun_out/r2_graph_v5 -f python tools/grp_ncu_run.py 37888 > gpurun_out/r2_run11b.txt 2>&1
