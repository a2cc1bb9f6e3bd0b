cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export C3POA_GRP_TIMING=1
(python tools/grp_ncu_run.py 100000; python tools/grp_ncu_run.py 37888) 2>&1 > gpurun_out/r2_run23.txt
cat gpurun_out/r2_run23.txt
