#!/bin/bash
# GPU call: new VCT-brick tests first, then the whole GPU suite, then the default bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dist.py -m gpu -x -q -s 2>&1 | tail -25 > gpurun_out/pytest_vct_bricks.txt
cat gpurun_out/pytest_vct_bricks.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.txt
cat gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_cfg2_r1b.json 2> gpurun_out/bench_cfg2_r1b.err
tail -c 3000 gpurun_out/bench_cfg2_r1b.json; tail -5 gpurun_out/bench_cfg2_r1b.err
