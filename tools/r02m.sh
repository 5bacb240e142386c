#!/bin/bash
# scaling points at N GPUs (every command under timeout): C4 mesh, C5 Lamellar, WTE
N=${1:-2}
mkdir -p gpurun_out/r02m
run() { name=$1; shift; timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" > gpurun_out/r02m/$name.json 2> gpurun_out/r02m/$name.err; echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02m/$name.json").read().strip().splitlines()[-1])
    print("$name", "ms", round(d["ms_per_step"],4), "frac", round(d["roofline"]["frac"],3), "e2e_ms", round(1e3/d["e2e"]["value"],3), "parity", (d.get("parity") or {}).get("ok"), (d.get("parity") or {}).get("cv_rel"))
except Exception as e: print("$name failed", e)
PY
}
run c5_${N}gpu --workload C5 --steps 200 --warmup 10
run wte_${N}gpu --workload WTE --steps 200 --warmup 10
if [ "$2" != "nomesh" ]; then run c4_${N}gpu --workload C4 --steps 50 --warmup 5; fi
