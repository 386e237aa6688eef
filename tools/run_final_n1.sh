#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/r2_pytest_gpu_final.txt
echo "pytest done at $(( $(date +%s) - S )) s"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_N1_final.json 2> gpurun_out/r2_bench_N1_final.err
echo "bench rc=$? at $(( $(date +%s) - S )) s"; tail -c 300 gpurun_out/r2_bench_N1_final.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_N1_final.json').read().strip().splitlines()[-1])
print('cfg2', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['dominant_kernel'], d['roofline']['frac'], d['gpu_launches'], d['e2e']['checksum'])
for k,v in d.get('workloads',{}).items():
    if 'error' in v: print(k, v); continue
    print(k, round(v['value'],3), round(v['ms_per_step'],3), round(v['e2e']['ms_per_step'],3), v['dominant_kernel'], round(v['ms_dominant_kernel_rank0'],3), v['e2e']['checksum'])
PY
