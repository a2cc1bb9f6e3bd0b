cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --section SourceCounters --import-source on --clock-control none -k regex:c3_poa_graph -s 2 -c 1 -o gpurun_out/r2_graph_v6 -f python tools/grp_ncu_run.py 37888 > gpurun_out/r2_run13.txt 2>&1
