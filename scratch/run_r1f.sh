#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_eval_harness.py tests/test_host_cpu.py -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest_eval.txt
cat gpurun_out/pytest_eval.txt
