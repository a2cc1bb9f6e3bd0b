cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_b46.log 2>&1
tail -c 200 gpurun_out/r2_b46.log
for c in cfg3 cfg4 cfg5 cfg1; do
  python bench.py --config $c > gpurun_out/r2_bench46_$c.json 2> gpurun_out/r2_bench46_$c.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_bench46_$c.json"))
    print("$c", round(d["value"]), "resident", round(d["e2e"]["value"]), "e2e", round(d["ms_per_step"],1), "ms", {k: round(v,1) for k,v in d["stage_ms_per_step"].items()}, d["poa_kernel"]["group_kernel_done_rank0"], d["poa_kernel"]["warp_kernel_reads_rank0"], "gcups", round(d["poa_gcups"],1), d["parity"]["mismatches"], round(d["cpu_baseline"]["value"]), d["roofline"]["kernel"])
except Exception as e:
    print("$c failed", e); print(open("gpurun_out/r2_bench46_$c.err").read()[-1500:])
PY
done
