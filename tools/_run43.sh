cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q -x -k "conk or fused_cfg2 or auto_group" > gpurun_out/r2_t43.log 2>&1; tail -2 gpurun_out/r2_t43.log | cut -c1-300
python tools/_run36.py 2>&1 | tail -1 | cut -c1-120
