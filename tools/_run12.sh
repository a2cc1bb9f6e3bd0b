cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/grp_ncu_run.py 37888 > gpurun_out/r2_run12a.txt 2>&1; cat gpurun_out/r2_run12a.txt
python tools/grp_ncu_run.py 100000 > gpurun_out/r2_run12b.txt 2>&1; cat gpurun_out/r2_run12b.txt
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__thread_inst_executed_per_inst_executed.ratio,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:c3_poa_gr --csv --log-file gpurun_out/r2_kern12.csv python tools/grp_ncu_run.py 37888 > gpurun_out/r2_run12.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "auto and (poa or lane or fused)" > gpurun_out/r2_t12.log 2>&1
tail -4 gpurun_out/r2_t12.log
