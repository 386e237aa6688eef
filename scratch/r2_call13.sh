#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
echo "pytest done at $(( $(date +%s) - S )) s"
for sk in 1 0; do
VRB_LIST_SKIP=$sk timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_list\|k_dos_shade -c 9 --csv --log-file gpurun_out/r2_launches_cfg3_d$sk.csv python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
echo "skip=$sk"; grep -o 'k_list[a-z_]*.*\|k_dos_shade.*' gpurun_out/r2_launches_cfg3_d$sk.csv | awk -F'","' '{print substr($1,1,30), $NF}' | tail -3
done
timeout 300 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg3', d['value'], d['ms_per_step'], d['ms_per_frame_render_call_rank0'], d['e2e']['ms_per_step'], d['dominant_kernel'], d['ms_dominant_kernel_rank0'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_list_march -c 1 -o gpurun_out/r2_k_list_march_cfg3_v4 -f python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_a.log 2>&1
echo "done at $(( $(date +%s) - S )) s"
