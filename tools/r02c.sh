#!/bin/bash
mkdir -p gpurun_out/r02c
./tools/micro/tma_reduce_test > gpurun_out/r02c/tma_reduce_test.log 2>&1; cat gpurun_out/r02c/tma_reduce_test.log
B="python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-parity"
run() { name=$1; shift; $B "$@" > gpurun_out/r02c/$name.json 2>gpurun_out/r02c/$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02c/$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["ms_per_step"],4), d["roofline"]["stage_ms"])
except Exception as e: print("$name failed", e)
PY
}
run default
run noflush --mesh-knob 12=1
run noatomics --mesh-knob 12=2
run neither --mesh-knob 12=3
run cache --mesh-knob 9=1
run cache_noflush --mesh-knob 9=1 --mesh-knob 12=1
run notmagather --mesh-knob 11=0
python -m pytest tests -m gpu -x -q > gpurun_out/r02c/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02c/pytest.log
