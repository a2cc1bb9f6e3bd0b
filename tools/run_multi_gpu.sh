cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
run() {  # name, args...
  name=$1; shift
  timeout 900 $TR bench.py --gpus $N "$@" > gpurun_out/r02_n${N}_$name.json 2> gpurun_out/r02_n${N}_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_n${N}_$name.json").read().strip().splitlines()[-1])
    print("$name N=$N", round(d["value"]), "reads/s", round(d["e2e"]["value"]), "e2e", "ms/step", round(d["ms_per_step"],1), d["scaling"], "gcups", d.get("poa_gcups"), d.get("poa_kernel"))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/r02_n${N}_$name.err").read()[-800:])
PY
}
if [ "$N" = "8" ]; then
run cfg2_weak --steps 3 --warmup 3
run cfg2_strong --steps 3 --warmup 3 --scaling strong --reads 100000
run cfg2_ref --impl reference --steps 1 --warmup 1
run cfg3 --config cfg3 --steps 2 --warmup 3
run cfg4 --config cfg4 --steps 2 --warmup 3
fi
run cfg5 --config cfg5 --steps 2 --warmup 3
