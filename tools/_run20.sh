cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export C3POA_GRP_TIMING=1
(echo base; python tools/grp_ncu_run.py 100000
echo nopf2; C3POA_GPU_LIB=build/variants/lib_nopf2.so python tools/grp_ncu_run.py 100000
echo fetch32; C3POA_L2_FETCH=32 python tools/grp_ncu_run.py 100000
echo nopf2+fetch32; C3POA_L2_FETCH=32 C3POA_GPU_LIB=build/variants/lib_nopf2.so python tools/grp_ncu_run.py 100000
echo dp3; C3POA_GRP_DP_CTAS=3 python tools/grp_ncu_run.py 100000
echo dp2; C3POA_GRP_DP_CTAS=2 python tools/grp_ncu_run.py 100000 ) 2>&1 | grep -v "^timings" > gpurun_out/r2_run20.txt
cat gpurun_out/r2_run20.txt
