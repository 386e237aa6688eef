#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_list_march -c 1 -o gpurun_out/r2_k_list_march_cfg3 -f python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_a.log 2>&1
tail -2 gpurun_out/ncu_a.log
echo "ncu a at $(( $(date +%s) - S )) s"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gt_shade -c 1 -o gpurun_out/r2_k_gt_shade_cfg4 -f python bench.py --workload cfg4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
tail -2 gpurun_out/ncu_b.log
echo "ncu b at $(( $(date +%s) - S )) s"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2_launches_cfg3_b.csv python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
echo "done at $(( $(date +%s) - S )) s"
