#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/vct_brick_one.py --res 256 --size 640 360 --world 8 --steps 3 > gpurun_out/vct_brick_one_256.json 2> gpurun_out/vct_brick_one_256.err
cat gpurun_out/vct_brick_one_256.json; tail -3 gpurun_out/vct_brick_one_256.err
timeout 420 python tools/vct_brick_one.py --res 2048 --size 3840 2160 --world 8 --steps 5 > gpurun_out/vct_brick_one_2048_exact.json 2> gpurun_out/vct_brick_one_2048_exact.err
cat gpurun_out/vct_brick_one_2048_exact.json; tail -3 gpurun_out/vct_brick_one_2048_exact.err
timeout 420 python tools/vct_brick_one.py --res 2048 --size 3840 2160 --world 8 --steps 5 --filter hardware > gpurun_out/vct_brick_one_2048_hw.json 2> gpurun_out/vct_brick_one_2048_hw.err
cat gpurun_out/vct_brick_one_2048_hw.json; tail -3 gpurun_out/vct_brick_one_2048_hw.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg2_r1c.json 2> gpurun_out/bench_cfg2_r1c.err
python -c "
import json; d=json.load(open('gpurun_out/bench_cfg2_r1c.json')); print(d['ms_per_step'], d['roofline'])"
