cd $GRAFT_REPO_ROOT
C3POA_GRP_TIMING=1 python tools/_run36.py 2>&1 | grep "host:" | awk '{print $9}' | tr '\n' ' '; echo
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench38.json 2> gpurun_out/r2_bench38.err; python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench38.json"))
print(round(d["value"]), round(d["e2e"]["value"]), d["step_ms_rank0"], d["e2e"]["step_ms_rank0"], d["stage_ms_per_step"]["poa_ms"], d["parity"]["mismatches"])
PY
