cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export C3POA_GRP_TIMING=1
(echo base; python tools/grp_ncu_run.py 100000; python tools/grp_ncu_run.py 16000
for v in bk4 mk8; do echo $v; C3POA_GPU_LIB=build/variants/lib_$v.so python tools/grp_ncu_run.py 100000; C3POA_GPU_LIB=build/variants/lib_$v.so python tools/grp_ncu_run.py 16000; done
echo SURVEY; python tools/config_survey.py auto 4 ) 2>&1 | grep -v "^timings" > gpurun_out/r2_run25.txt
cat gpurun_out/r2_run25.txt
