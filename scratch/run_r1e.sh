#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multiscaling.py -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest_multiscaling.txt
cat gpurun_out/pytest_multiscaling.txt
