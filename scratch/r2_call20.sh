#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s)
timeout 300 python -m pytest tests/test_gt_vct.py -m gpu -q -x 2>&1 | tail -2
for k in deferred ray; do
VRB_VCT_KERNEL=$k timeout 300 python bench.py --workload cfg5-1gpu --steps 10 --warmup 3 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg5-1gpu $k', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['dominant_kernel'], d['ms_dominant_kernel_rank0'], d['e2e']['checksum'])"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_list\|k_ebs\|k_vct -c 12 --csv --log-file gpurun_out/r2_launches_cfg2.csv python bench.py --workload cfg2 --extras none --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
grep -o 'k_list[a-z_]*.*\|k_ebs_shade.*' gpurun_out/r2_launches_cfg2.csv | awk -F'","' '{print substr($1,1,30), $NF}' | tail -6
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_list\|k_vct -c 12 --csv --log-file gpurun_out/r2_launches_cfg5.csv python bench.py --workload cfg5-1gpu --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
grep -o 'k_list[a-z_]*.*\|k_vct_shade.*' gpurun_out/r2_launches_cfg5.csv | awk -F'","' '{print substr($1,1,30), $NF}' | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_vct_shade -c 1 -o gpurun_out/r2_k_vct_shade_cfg5 -f python bench.py --workload cfg5-1gpu --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_a.log 2>&1
echo "done at $(( $(date +%s) - S )) s"
