#!/bin/bash
mkdir -p gpurun_out/r02g
python -m pytest tests -m gpu -x -q > gpurun_out/r02g/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02g/pytest.log
