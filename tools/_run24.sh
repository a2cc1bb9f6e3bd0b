cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export C3POA_GRP_TIMING=1
(echo AUTO; python tools/config_survey.py auto 4; echo WARP; python tools/config_survey.py warp 4; echo LANE; python tools/config_survey.py lane 4) > gpurun_out/r2_run24.txt 2>&1
cat gpurun_out/r2_run24.txt
