cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --section SourceCounters --import-source on --clock-control none -k regex:c3_poa_grp_dp -s 1 -c 1 -o gpurun_out/r2_dp_v4 -f python tools/grp_ncu_run.py 37888 > gpurun_out/r2_run10.txt 2>&1
timeout 900 ncu --section SourceCounters --import-source on --clock-control none -k regex:c3_poa_grp_graph -s 2 -c 1 -o gpurun_out/r2_graph_v4 -f python tools/grp_ncu_run.py 37888 > gpurun_out/r2_run10b.txt 2>&1
