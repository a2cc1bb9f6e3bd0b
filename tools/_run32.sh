cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:c3_peaks_kernel -c 1 -o gpurun_out/r02_peaks_v1 -f python tools/grp_ncu_run.py 50000 auto > gpurun_out/r2_run32.txt 2>&1
tail -2 gpurun_out/r2_run32.txt | cut -c1-200
