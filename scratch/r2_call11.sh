#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
echo "pytest done at $(( $(date +%s) - S )) s"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1_b.json 2> gpurun_out/r2_bench_n1_b.err
echo "bench rc=$? at $(( $(date +%s) - S )) s"; tail -c 600 gpurun_out/r2_bench_n1_b.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n1_b.json').read().strip().splitlines()[-1])
print('cfg2', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['dominant_kernel'], d['ms_dominant_kernel_rank0'], d['roofline']['frac'], d['e2e']['checksum'])
for k,v in d.get('workloads',{}).items():
    if 'error' in v: print(k, v); continue
    print(k, v['value'], v['ms_per_step'], v['e2e']['ms_per_step'], v['dominant_kernel'], v['ms_dominant_kernel_rank0'], v['roofline']['frac'], v.get('roofline_ldg16',{}).get('frac'), v['wall_s'], v['e2e']['checksum'])
PY
VRB_EBS_KERNEL=coop timeout 300 python bench.py --steps 10 --warmup 3 --extras none --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg2 coop', d['value'], d['ms_per_step'], d['dominant_kernel'], d['ms_dominant_kernel_rank0'], d['e2e']['checksum'])"
echo "done at $(( $(date +%s) - S )) s"
