#!/bin/bash
mkdir -p gpurun_out/r02o
B="timeout 100 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extra-legs --no-parity"
run() { name=$1; shift; "$@" > gpurun_out/r02o/$name.json 2>gpurun_out/r02o/$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02o/$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["ms_per_step"],4), d["roofline"]["stage_ms"])
except Exception as e: print("$name failed", e)
PY
}
for v in 1 2 3; do METAD_XY_VARIANT=$v run c4_v$v $B; done
for v in 1 2; do METAD_XY_VARIANT=$v run c3_v$v $B --workload C3; done
METAD_XY_VARIANT=2 timeout 100 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_xy" 2>&1 | tail -2
METAD_XY_VARIANT=1 timeout 100 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_xy" 2>&1 | tail -2
