#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s)
nproc; free -g | head -2
timeout 1500 python -m pytest tests -m gpu -q -x --durations=15 2>&1 | tail -40 > gpurun_out/r2_pytest_gpu_2.txt
cat gpurun_out/r2_pytest_gpu_2.txt; echo "pytest done at $(( $(date +%s) - S )) s"
