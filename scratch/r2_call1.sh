#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s)
timeout 400 python scratch/r2_dos_ab.py 2>&1 | tail -12 | tee gpurun_out/r2_dos_ab.txt
echo "ab done at $(( $(date +%s) - S )) s"
timeout 400 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2_pytest_gpu_1.txt
cat gpurun_out/r2_pytest_gpu_1.txt; echo "pytest done at $(( $(date +%s) - S )) s"
timeout 200 python tests/gpu_random_sweep.py --seed 1 --scenes 12 > gpurun_out/r2_random_sweep_exact.jsonl 2>&1; tail -3 gpurun_out/r2_random_sweep_exact.jsonl
echo "sweep done at $(( $(date +%s) - S )) s"
