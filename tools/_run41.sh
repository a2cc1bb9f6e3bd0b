cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "conk" > gpurun_out/r2_t41.log 2>&1; tail -5 gpurun_out/r2_t41.log | cut -c1-400
python tools/grp_ncu_run.py 100000 auto 2>&1 | tail -1 | cut -c1-200
C3POA_CONK_INT32=1 python tools/grp_ncu_run.py 100000 auto 2>&1 | tail -1 | cut -c1-200
