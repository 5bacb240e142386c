#!/bin/bash
# 8 GPUs: C4 mesh (peer memory), C5 / WTE sharded with the peer-memory all-reduce and with NCCL
mkdir -p gpurun_out/r02i
N=${1:-8}
run() { name=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" > gpurun_out/r02i/$name.json 2> gpurun_out/r02i/$name.err; echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02i/$name.json").read().strip().splitlines()[-1])
    print("$name", "ms", round(d["ms_per_step"],4), "frac", round(d["roofline"]["frac"],3), "e2e", round(1e3/d["e2e"]["value"],3), "parity", (d.get("parity") or {}).get("ok"), (d.get("parity") or {}).get("cv_rel"))
    if "segment_ms_max_over_ranks" in d["roofline"]: print("   ", d["roofline"]["segment_ms_max_over_ranks"])
except Exception as e: print("$name failed", e)
PY
}
run c4_${N}gpu --workload C4 --steps 50 --warmup 5
run c5_${N}gpu --workload C5 --steps 200 --warmup 10
run c5_${N}gpu_nccl --workload C5 --steps 200 --warmup 10 --comm nccl --no-parity
run c5_${N}gpu_nograph --workload C5 --steps 200 --warmup 10 --no-graph --no-parity
run wte_${N}gpu --workload WTE --steps 200 --warmup 10
run wte_${N}gpu_nccl --workload WTE --steps 200 --warmup 10 --comm nccl --no-parity
tail -n 3 gpurun_out/r02i/*.err | cut -c1-200 | tail -20
