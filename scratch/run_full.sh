export PYTHONPATH=$PWD
timeout 1200 python -m pytest tests/test_full_size_gpu.py -m gpu -q -x -s > gpurun_out/full.txt 2>&1
grep -E "^SAT texels|^non-finite|^finite pixels|^worst|passed|failed|Error|^E  " gpurun_out/full.txt | head -20
