cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2_t44.log 2>&1; tail -2 gpurun_out/r2_t44.log | cut -c1-300
python tools/parity_100k.py 100000 > gpurun_out/r2_parity44.txt 2>&1; tail -1 gpurun_out/r2_parity44.txt | cut -c1-400
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench44.json 2> gpurun_out/r2_bench44.err; python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench44.json"))
print(round(d["value"]), round(d["e2e"]["value"]), d["step_ms_rank0"], d["e2e"]["step_ms_rank0"], d["stage_ms_per_step"], d["parity"]["mismatches"], d["conk_gcups"])
PY
timeout 600 ncu --set full --import-source on --clock-control none -k regex:c3_conk2_kernel -c 1 -o gpurun_out/r02_conk2_full -f python tools/grp_ncu_run.py 30000 auto > gpurun_out/r2_run44.txt 2>&1; tail -1 gpurun_out/r2_run44.txt | cut -c1-100
