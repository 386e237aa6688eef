export PYTHONPATH=$PWD
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
python bench.py 2>/dev/null | tail -1 > gpurun_out/bench_cfg2_default.json
python -c "
import json; d=json.load(open('gpurun_out/bench_cfg2_default.json'))
print('cfg2', d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['checksum'], d['e2e']['nonfinite_values'])
print(d['roofline']); print(d['roofline_sat']); print(d['cpu_baseline']); print(d['clocks']); print('launches', d['gpu_launches'])"
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-600
