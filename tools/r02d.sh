#!/bin/bash
mkdir -p gpurun_out/r02d
B="python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-parity"
run() { name=$1; shift; $B "$@" > gpurun_out/r02d/$name.json 2>gpurun_out/r02d/$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02d/$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["ms_per_step"],4), d["roofline"]["stage_ms"])
except Exception as e: print("$name failed", e)
PY
}
run default
run noflush --late-knob 12=1
run noatomics --late-knob 12=2
run neither --late-knob 12=3
run cache_noflush --mesh-knob 9=1 --late-knob 12=1
run cache_neither --mesh-knob 9=1 --late-knob 12=3
