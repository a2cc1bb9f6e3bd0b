cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "peaks or fused or golden or split or auto_group" > gpurun_out/r2_t31.log 2>&1; tail -4 gpurun_out/r2_t31.log
python tools/grp_ncu_run.py 100000 auto 2>&1 | tail -1 | cut -c1-330
python tools/grp_ncu_run.py 12500 auto 2>&1 | tail -1 | cut -c1-330
