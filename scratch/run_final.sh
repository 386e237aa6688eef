#!/bin/bash
# final validation of the round: GPU suite, smoke, default bench line, ncu launch list of the same bench command
mkdir -p gpurun_out
S=$(date +%s)
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu_final.txt
cat gpurun_out/pytest_gpu_final.txt; echo "pytest done at $(( $(date +%s) - S )) s"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_cfg2_final.json 2> gpurun_out/bench_cfg2_final.err
python -c "
import json; d=json.load(open('gpurun_out/bench_cfg2_final.json'))
print('cfg2', d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['checksum'], 'launches', d['gpu_launches'], 'frac', d['roofline']['frac'], 'sat', d['roofline_sat']['frac'], d['clocks'])"
echo "bench done at $(( $(date +%s) - S )) s"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^(k_ebs|k_frame|k_sat_fill|k_sat_scan|k_sat_atlas|k_iota|k_gather|k_l1|Device)' -c 300 --csv --log-file gpurun_out/launches_cfg2_final.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
wc -l gpurun_out/launches_cfg2_final.csv; echo "ncu done at $(( $(date +%s) - S )) s"
