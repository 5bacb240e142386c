#!/bin/bash
# GPU experiment batch r02b: tests + A/B of the spread / gather variants at C4
mkdir -p gpurun_out/r02b
python -m pytest tests -m gpu -x -q > gpurun_out/r02b/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02b/pytest.log
B="python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-parity"
$B > gpurun_out/r02b/c4_default.json 2>gpurun_out/r02b/c4_default.err
$B --mesh-knob 9=1 > gpurun_out/r02b/c4_cache.json 2>gpurun_out/r02b/c4_cache.err
$B --mesh-knob 10=0 > gpurun_out/r02b/c4_notmaflush.json 2>&1
$B --mesh-knob 11=0 > gpurun_out/r02b/c4_notmagather.json 2>&1
$B --mesh-knob 9=1 --mesh-knob 10=0 --mesh-knob 11=0 > gpurun_out/r02b/c4_cache_notma.json 2>&1
for f in default cache notmaflush notmagather cache_notma; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02b/c4_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["ms_per_step"],4), d["roofline"]["stage_ms"])
except Exception as e: print("$f failed", e)
PY
done
python bench.py --steps 40 --warmup 5 --cpu-budget 8 > gpurun_out/r02b/c4_full.json 2>gpurun_out/r02b/c4_full.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r02b/c4_full.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['parity'])"
