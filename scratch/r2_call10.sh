#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_list\|k_dos_shade -c 9 --csv --log-file gpurun_out/r2_launches_cfg3_c.csv python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
grep -o 'k_list[a-z_]*.*\|k_dos_shade.*' gpurun_out/r2_launches_cfg3_c.csv | awk -F'","' '{print substr($1,1,30), $NF}' | tail -6
timeout 300 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_b_cfg3.json 2> gpurun_out/r2_b_cfg3.err || tail -5 gpurun_out/r2_b_cfg3.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_b_cfg3.json').read().strip().splitlines()[-1])
print('cfg3', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['dominant_kernel'], d['ms_dominant_kernel_rank0'])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_list_march -c 1 -o gpurun_out/r2_k_list_march_cfg3_v2 -f python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_a.log 2>&1
timeout 600 python -m pytest tests/test_dos.py tests/test_gt_vct.py -m gpu -q -x 2>&1 | tail -3
echo "done at $(( $(date +%s) - S )) s"
