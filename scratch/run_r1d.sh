#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gradient.py -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest_gradient.txt
cat gpurun_out/pytest_gradient.txt
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.txt
cat gpurun_out/pytest_gpu.txt
