#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s)
timeout 300 python tools/vct_brick_one.py --res 2048 --size 3840 2160 --world 8 | tail -1 | tee gpurun_out/r2_cfg5_one_rank_of_8_2048_u16_4k_exact.json
echo "plain at $(( $(date +%s) - S )) s"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_vct_brick\|k_rc1pass_brick --launch-skip 6 -c 18 -o gpurun_out/r2_k_brick_window_2048 -f python tools/vct_brick_one.py --res 2048 --size 3840 2160 --world 8 > gpurun_out/ncu_a.log 2>&1
tail -2 gpurun_out/ncu_a.log
echo "done at $(( $(date +%s) - S )) s"
