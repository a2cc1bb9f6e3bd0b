set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "poa or lane or fused" > gpurun_out/r2_t1.log 2>&1
tail -5 gpurun_out/r2_t1.log
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_b1_auto.json 2> gpurun_out/r2_b1_auto.err
cat gpurun_out/r2_b1_auto.json | head -c 1500
tail -3 gpurun_out/r2_b1_auto.err
