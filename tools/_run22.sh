cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --section SourceCounters --import-source on --clock-control none -k regex:c3_poa_graph_kernel -s 1 -c 1 -o gpurun_out/r2_graph_v8 -f python tools/grp_ncu_run.py 37888 > gpurun_out/r2_run22a.txt 2>&1
timeout 600 ncu --section SourceCounters --import-source on --clock-control none -k regex:c3_poa_grp_dp -s 1 -c 1 -o gpurun_out/r2_dp_v8 -f python tools/grp_ncu_run.py 37888 > gpurun_out/r2_run22b.txt 2>&1
tail -2 gpurun_out/r2_run22a.txt gpurun_out/r2_run22b.txt
