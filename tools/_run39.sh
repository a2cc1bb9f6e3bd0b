cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r2_t39.log 2>&1; tail -3 gpurun_out/r2_t39.log | cut -c1-300
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench39.json 2> gpurun_out/r2_bench39.err; python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench39.json"))
print(round(d["value"]), round(d["e2e"]["value"]), d["step_ms_rank0"], d["e2e"]["step_ms_rank0"], d["stage_ms_per_step"]["poa_ms"], d["parity"]["mismatches"])
PY
