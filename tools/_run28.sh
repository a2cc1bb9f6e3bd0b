cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/parity_100k.py 100000 > gpurun_out/r2_parity28.txt 2>&1; tail -1 gpurun_out/r2_parity28.txt | cut -c1-900
for c in cfg3 cfg4 cfg5 cfg1; do
  python bench.py --config $c > gpurun_out/r2_bench28_$c.json 2> gpurun_out/r2_bench28_$c.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_bench28_$c.json"))
    print("$c", round(d["value"]), "reads/s resident", round(d["e2e"]["value"]), "e2e", d["stage_ms_per_step"], d["poa_kernel"], "gcups", d["poa_gcups"], d["parity"], d["cpu_baseline"]["value"], d["roofline"]["kernel"], d["roofline_int"]["frac"])
except Exception as e:
    print("$c failed", e); print(open("gpurun_out/r2_bench28_$c.err").read()[-1500:])
PY
done
