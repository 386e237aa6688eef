#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s)
timeout 900 ncu --set full --clock-control none -k regex:k_vct_brick --launch-skip 6 -c 2 -o /tmp/brick_vct -f python tools/vct_brick_one.py --res 2048 --size 3840 2160 --world 8 > gpurun_out/ncu_a.log 2>&1
python profiles/summarize.py /tmp/brick_vct.ncu-rep > gpurun_out/r2_cfg5_window_k_vct_brick.txt
timeout 900 ncu --set full --clock-control none -k regex:k_rc1pass_brick --launch-skip 6 -c 2 -o /tmp/brick_rc -f python tools/vct_brick_one.py --res 2048 --size 3840 2160 --world 8 > gpurun_out/ncu_b.log 2>&1
python profiles/summarize.py /tmp/brick_rc.ncu-rep > gpurun_out/r2_cfg5_window_k_rc1pass_brick.txt
ls -la /tmp/*.ncu-rep
cat gpurun_out/r2_cfg5_window_k_vct_brick.txt gpurun_out/r2_cfg5_window_k_rc1pass_brick.txt | grep -v "^stall\|launch__\|cycles_elapsed"
echo "done at $(( $(date +%s) - S )) s"
