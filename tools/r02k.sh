#!/bin/bash
mkdir -p gpurun_out/r02k
B="timeout 120 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-parity --no-extra-legs"
run() { name=$1; shift; "$@" > gpurun_out/r02k/$name.json 2>gpurun_out/r02k/$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02k/$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["ms_per_step"],4), d["roofline"]["stage_ms"])
except Exception as e: print("$name failed", e)
PY
}
run default $B
METAD_GATHER_VARIANT=5 run g192x4 $B
METAD_GATHER_VARIANT=8 run g160x4 $B
METAD_GATHER_VARIANT=3 run g128x5 $B
METAD_GATHER_VARIANT=1 run g256x2 $B
