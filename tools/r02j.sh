#!/bin/bash
mkdir -p gpurun_out/r02j
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r02j/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02j/pytest.log
timeout 400 python bench.py --steps 50 --warmup 5 > gpurun_out/r02j/bench_c4.json 2> gpurun_out/r02j/bench_c4.err; echo "bench rc=$?"
tail -n 4 gpurun_out/r02j/bench_c4.err | cut -c1-300
python - <<PY
import json
d=json.loads(open("gpurun_out/r02j/bench_c4.json").read().strip().splitlines()[-1])
for k in ("value","ms_per_step","e2e","parity","tile_order","orders_ms_per_step","moving","host_class"): print(k, d.get(k))
print(d["roofline"]["stage_ms"], d["roofline"]["frac"], d["roofline"]["step_frac"])
PY
timeout 120 python bench.py --workload WTE --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r02j/bench_wte.json 2> gpurun_out/r02j/bench_wte.err; echo "wte rc=$?"; tail -c 600 gpurun_out/r02j/bench_wte.json | head -c 400
