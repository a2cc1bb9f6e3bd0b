cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/grp_prof_run.py 37888 > gpurun_out/r2_gprof2.txt 2>&1
cat gpurun_out/r2_gprof2.txt
