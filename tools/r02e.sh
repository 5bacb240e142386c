#!/bin/bash
mkdir -p gpurun_out/r02e
python -m pytest tests -m gpu -x -q > gpurun_out/r02e/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02e/pytest.log
B="python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-parity"
run() { name=$1; shift; $B "$@" > gpurun_out/r02e/$name.json 2>gpurun_out/r02e/$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02e/$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["ms_per_step"],4), d["roofline"]["stage_ms"])
except Exception as e: print("$name failed", e)
PY
}
run default
run notmaflush --mesh-knob 10=0
run c3 --workload C3
