#!/bin/bash
mkdir -p gpurun_out/r02n
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_xy or every_fft_length or full_size_c3 or c1_golden" > gpurun_out/r02n/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02n/pytest.log
B="timeout 120 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extra-legs"
run() { name=$1; shift; "$@" > gpurun_out/r02n/$name.json 2>gpurun_out/r02n/$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02n/$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["ms_per_step"],4), d["roofline"]["stage_ms"], (d.get("parity") or {}).get("ok"), (d.get("parity") or {}).get("cv_rel"), (d.get("parity") or {}).get("force_rel_max"))
except Exception as e: print("$name failed", e)
PY
}
run fused $B
run unfused $B --mesh-knob 15=0 --no-parity
run c3_fused $B --workload C3 --no-parity
run c3_unfused $B --workload C3 --mesh-knob 15=0 --no-parity
tail -n 3 gpurun_out/r02n/fused.err | cut -c1-300
