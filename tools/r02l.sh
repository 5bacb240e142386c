#!/bin/bash
# ncu evidence for profiles/: launch list of the default bench command, then one full capture of every kernel of the step
mkdir -p gpurun_out/r02l
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/r02l/launches.csv \
    python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-parity --no-extra-legs > gpurun_out/r02l/ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"mesh_spread|mesh_gather|fft_" -s 40 -c 7 -o gpurun_out/r02l/full_c4 \
    python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-parity --no-extra-legs > gpurun_out/r02l/ncu_full.log 2>&1; echo "full rc=$?"
ls -la gpurun_out/r02l
