cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke27.txt 2>&1; tail -4 gpurun_out/r2_smoke27.txt | cut -c1-300
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_t27.log 2>&1; tail -8 gpurun_out/r2_t27.log
