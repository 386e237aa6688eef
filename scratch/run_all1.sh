export PYTHONPATH=$PWD
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
python bench.py 2>/dev/null | tail -1 > gpurun_out/bench_cfg2_default.json
python -c "
import json; d=json.load(open('gpurun_out/bench_cfg2_default.json'))
print('cfg2', d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['checksum'], d['e2e']['nonfinite_values'], 'launches', d['gpu_launches'])
print(d['roofline']['frac'], d['roofline_sat'])"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^(k_ebs|k_frame|k_sat_fill|k_sat_scan|k_sat_atlas|k_iota|Device)' -c 300 --csv --log-file gpurun_out/launches_cfg2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
wc -l gpurun_out/launches_cfg2.csv
