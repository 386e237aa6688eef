#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s)
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1_a.json 2> gpurun_out/r2_bench_n1_a.err
echo "bench rc=$? at $(( $(date +%s) - S )) s"; tail -c 600 gpurun_out/r2_bench_n1_a.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n1_a.json').read().strip().splitlines()[-1])
print('cfg2', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['dominant_kernel'], d['ms_dominant_kernel_rank0'], d['roofline']['frac'])
for k,v in d.get('workloads',{}).items():
    if 'error' in v: print(k, v); continue
    print(k, v['value'], v['ms_per_step'], v['e2e']['ms_per_step'], v['dominant_kernel'], v['ms_dominant_kernel_rank0'], v['roofline']['frac'], v.get('roofline_ldg16',{}).get('frac'), v['wall_s'])
PY
echo "done at $(( $(date +%s) - S )) s"
