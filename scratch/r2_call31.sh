#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/r2_pytest_gpu_final.txt
echo "pytest done at $(( $(date +%s) - S )) s"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/r2_smoke_final.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_N1_final.json 2> gpurun_out/r2_bench_N1_final.err
echo "bench rc=$? at $(( $(date +%s) - S )) s"; tail -c 400 gpurun_out/r2_bench_N1_final.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_N1_final.json').read().strip().splitlines()[-1])
print('cfg2', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['dominant_kernel'], d['ms_dominant_kernel_rank0'], d['roofline']['frac'], d['roofline'].get('traffic'), d['gpu_launches'], d['e2e']['checksum'])
print({k:d['roofline_sat'][k] for k in ('ms','frac','reference_order_ms','reference_order_frac')})
print('cpu_baseline', d.get('cpu_baseline',{}).get('value'), d.get('cpu_baseline',{}).get('cores'))
for k,v in d.get('workloads',{}).items():
    if 'error' in v: print(k, v); continue
    print(k, round(v['value'],3), round(v['ms_per_step'],3), round(v['e2e']['ms_per_step'],3), v['dominant_kernel'], round(v['ms_dominant_kernel_rank0'],3), round(v['roofline']['frac'],3), v.get('roofline_ldg16',{}).get('frac'), v['roofline'].get('traffic'), round(v['wall_s'],1), v['e2e']['checksum'])
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('reference arm', d['value'], d['ms_per_step'], d['cpu_baseline']['cores'])"
echo "done at $(( $(date +%s) - S )) s"
