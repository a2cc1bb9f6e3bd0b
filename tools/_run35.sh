cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2_t35.log 2>&1; tail -12 gpurun_out/r2_t35.log | cut -c1-400
python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench35.json 2> gpurun_out/r2_bench35.err; python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench35.json"))
print(round(d["value"]), round(d["e2e"]["value"]), d["stage_ms_per_step"], d["parity"], d["roofline_int"]["frac"], d["roofline"]["traffic"], d["roofline"]["frac"])
PY
