#!/bin/bash
# Row a3 (triclinic boxes, any mesh size): the three B200 calls that validated it, each one `gpurun` call under `timeout`.
# 1. the 26 new GPU tests + a short bench (no regression of the orthorhombic kernels)
timeout 85 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py tests/test_gpu_api.py -m gpu -q --tb=short \
  -k "triclinic or general or rejects_unsupported or rejects_bad or power_of_two" > gpurun_out/a3_tests.txt 2>&1
timeout 28 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra-legs --no-parity > gpurun_out/a3_bench.json 2> gpurun_out/a3_bench.err
# 2. the one test whose tolerance was corrected + the existing mesh tests
timeout 52 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py tests/test_gpu_sharded.py -m gpu -q --tb=short -x \
  -k "stale_order_and_epilogues or test_mesh_cv_forces_cells or qmax_and_virial or every_fft_length or test_mesh_against_reference or c1_golden or log_quantities or reference_test_mesh or peer_memory_path or dense_cells" > gpurun_out/a3b_tests.txt 2>&1
# 3. per-stage timings of the new paths
timeout 25 python tools/a3_timing.py > gpurun_out/a3_timing.json 2> gpurun_out/a3_timing.err
