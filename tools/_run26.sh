cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke26.txt 2>&1; tail -4 gpurun_out/r2_smoke26.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t26.log 2>&1; tail -5 gpurun_out/r2_t26.log
python bench.py > gpurun_out/r2_bench26.json 2> gpurun_out/r2_bench26.err; tail -c 3000 gpurun_out/r2_bench26.json; tail -3 gpurun_out/r2_bench26.err
