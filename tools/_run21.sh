cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export C3POA_GRP_TIMING=1
(for v in pf0 pf1 pf0bk8 pf0bk2; do echo $v; C3POA_GPU_LIB=build/variants/lib_$v.so python tools/grp_ncu_run.py 100000; done
echo gprof; python tools/grp_prof_run.py 100000 ) 2>&1 | grep -v "^timings" > gpurun_out/r2_run21.txt
cat gpurun_out/r2_run21.txt
