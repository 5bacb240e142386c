#!/bin/bash
# final validation of the round: GPU tests, smoke, the default bench (C4) and the small single-GPU workloads
mkdir -p gpurun_out/r02p
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r02p/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02p/pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02p/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02p/smoke.log
timeout 400 python bench.py --steps 50 --warmup 5 > gpurun_out/r02p/bench_c4.json 2> gpurun_out/r02p/bench_c4.err; echo "bench c4 rc=$?"
timeout 100 python bench.py --workload C3 --steps 100 --warmup 5 --cpu-budget 5 > gpurun_out/r02p/bench_c3.json 2> gpurun_out/r02p/bench_c3.err; echo "bench c3 rc=$?"
timeout 60 python bench.py --workload C1 --steps 200 --warmup 5 --cpu-budget 3 > gpurun_out/r02p/bench_c1.json 2> gpurun_out/r02p/bench_c1.err; echo "bench c1 rc=$?"
python - <<PY
import json
for n in ("c4","c3","c1"):
    try:
        d=json.loads(open("gpurun_out/r02p/bench_%s.json" % n).read().strip().splitlines()[-1])
        print(n, round(d["ms_per_step"],4), d["roofline"].get("stage_ms"), "parity", d["parity"]["ok"], d["parity"]["cv_rel"], d["parity"]["force_rel_max"], "e2e", d["e2e"]["value"])
    except Exception as e: print(n, "failed", e)
PY
