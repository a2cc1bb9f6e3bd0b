cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/grp_prof_run.py 37888 > gpurun_out/r2_gprof3.txt 2>&1
cat gpurun_out/r2_gprof3.txt
timeout 900 ncu --section SourceCounters --section SchedulerStats --section WarpStateStats --section LaunchStats --section Occupancy --section MemoryWorkloadAnalysis --import-source on --clock-control none -k regex:c3_poa_grp_kernel -c 1 -o gpurun_out/r2_grp_v2 -f python tools/grp_ncu_run.py 18944 > gpurun_out/r2_ncu2.log 2>&1
tail -3 gpurun_out/r2_ncu2.log
