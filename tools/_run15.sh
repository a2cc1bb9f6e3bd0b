cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/grp_ncu_run.py 37888 > gpurun_out/r2_run15a.txt 2>&1; cat gpurun_out/r2_run15a.txt
python tools/grp_ncu_run.py 100000 > gpurun_out/r2_run15b.txt 2>&1; cat gpurun_out/r2_run15b.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t15.log 2>&1; tail -5 gpurun_out/r2_t15.log
