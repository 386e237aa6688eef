#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_n8_a.json 2> gpurun_out/r2_bench_n8_a.err
echo "bench rc=$? at $(( $(date +%s) - S )) s"; tail -c 1500 gpurun_out/r2_bench_n8_a.err | grep -v OMP_NUM | tail -5
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n8_a.json').read().strip().splitlines()[-1])
print('cfg2', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['dominant_kernel'], d['ms_dominant_kernel_rank0'], d['ms_per_frame_render_call_rank0'], d['e2e']['checksum'])
for k,v in d.get('workloads',{}).items():
    print(k, {a:b for a,b in v.items() if a not in ('roofline','roofline_hbm','roofline_ldg16','config','init','roofline_l1_ldg')})
PY
echo "done at $(( $(date +%s) - S )) s"
