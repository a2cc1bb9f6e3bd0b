cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
C3POA_GRP_TIMING=1 python tools/grp_ncu_run.py 100000 > gpurun_out/r2_run18a.txt 2>&1; cat gpurun_out/r2_run18a.txt
C3POA_GRP_TIMING=1 python tools/grp_ncu_run.py 37888 > gpurun_out/r2_run18b.txt 2>&1; cat gpurun_out/r2_run18b.txt
