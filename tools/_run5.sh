cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "auto and (poa or lane or fused)" > gpurun_out/r2_t2.log 2>&1
tail -4 gpurun_out/r2_t2.log
python tools/grp_ncu_run.py 37888 > gpurun_out/r2_run5a.txt 2>&1; cat gpurun_out/r2_run5a.txt
python tools/grp_ncu_run.py 100000 > gpurun_out/r2_run5b.txt 2>&1; cat gpurun_out/r2_run5b.txt
nsys --version 2>/dev/null | head -1
