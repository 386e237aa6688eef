#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
echo "pytest done at $(( $(date +%s) - S )) s"
for sk in 1 0; do
VRB_LIST_SKIP=$sk timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_list\|k_dos_shade -c 9 --csv --log-file gpurun_out/r2_launches_cfg3_d$sk.csv python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
echo "skip=$sk"; grep -o 'k_list[a-z_]*.*\|k_dos_shade.*' gpurun_out/r2_launches_cfg3_d$sk.csv | awk -F'","' '{print substr($1,1,30), $NF}' | tail -3
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_list_march -c 1 -o gpurun_out/r2_k_list_march_cfg3_v5 -f python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_a.log 2>&1
timeout 100 python tests/gpu_random_sweep.py --seed 2 --scenes 12 > gpurun_out/r2_random_sweep_exact_b.jsonl 2>&1; tail -1 gpurun_out/r2_random_sweep_exact_b.jsonl
echo "done at $(( $(date +%s) - S )) s"
