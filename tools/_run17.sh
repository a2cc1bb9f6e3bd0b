cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
C3POA_GRP_GROW_PCT=15 C3POA_GRP_BUDGET_PCT=78 python tools/grp_ncu_run.py 100000 > gpurun_out/r2_run17a.txt 2>&1; cat gpurun_out/r2_run17a.txt
C3POA_GRP_GROW_PCT=20 C3POA_GRP_BUDGET_PCT=70 python tools/grp_ncu_run.py 100000 > gpurun_out/r2_run17b.txt 2>&1; cat gpurun_out/r2_run17b.txt
