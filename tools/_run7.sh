cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --import-source on --clock-control none -k regex:c3_poa_grp_graph -s 2 -c 1 -o gpurun_out/r2_graph_v3 -f python tools/grp_ncu_run.py 37888 > gpurun_out/r2_run7.txt 2>&1
tail -2 gpurun_out/r2_run7.txt
timeout 900 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --import-source on --clock-control none -k regex:c3_poa_grp_dp -s 2 -c 1 -o gpurun_out/r2_dp_v3 -f python tools/grp_ncu_run.py 37888 > gpurun_out/r2_run7b.txt 2>&1
tail -2 gpurun_out/r2_run7b.txt
