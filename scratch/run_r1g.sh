#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_iso.py -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_iso.txt
cat gpurun_out/pytest_iso.txt
