#!/bin/bash
# 1 GPU: tests, then the small workloads with graph replay
mkdir -p gpurun_out/r02h
python -m pytest tests -m gpu -x -q > gpurun_out/r02h/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02h/pytest.log
for wl in C2 C5 WTE; do
  python bench.py --workload $wl --steps 200 --warmup 10 --cpu-budget 5 > gpurun_out/r02h/bench_${wl,,}.json 2> gpurun_out/r02h/bench_${wl,,}.err; echo "$wl rc=$?"
  python bench.py --workload $wl --steps 200 --warmup 10 --no-cpu-baseline --no-parity --no-graph > gpurun_out/r02h/bench_${wl,,}_nograph.json 2>&1
  python - <<PY
import json
for tag in ("", "_nograph"):
    try:
        d=json.loads(open("gpurun_out/r02h/bench_${wl,,}%s.json" % tag).read().strip().splitlines()[-1])
        print("$wl"+tag, round(d["ms_per_step"],4), "frac", round(d["roofline"]["frac"],3), d.get("parity",{}).get("ok"), d.get("parity",{}).get("cv_rel"), d.get("parity",{}).get("force_rel_max"))
    except Exception as e: print("$wl"+tag, "failed", e)
PY
done
tail -3 gpurun_out/r02h/*.err
