#!/bin/bash
# the round's closing single-GPU record: python bench.py with the driver's flags -> gpurun_out/r2_bench_N1_final.json
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_N1_final.json 2> gpurun_out/r2_bench_N1_final.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r2_bench_N1_final.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_N1_final.json').read().strip().splitlines()[-1])
print('cfg2', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['dominant_kernel'], d['roofline']['frac'], d.get('roofline_issue',{}).get('frac'), d['gpu_launches'], d['e2e']['checksum'])
for k,v in d.get('workloads',{}).items():
    if 'error' in v: print(k, v); continue
    print(k, round(v['value'],3), round(v['ms_per_step'],3), round(v['e2e']['ms_per_step'],3), v['dominant_kernel'], round(v['ms_dominant_kernel_rank0'],3), round(v.get('roofline_issue',{}).get('frac',0),3), v['e2e']['checksum'])
PY
