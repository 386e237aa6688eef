#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_N8_final.json 2> gpurun_out/r2_bench_N8_final.err
echo "bench rc=$? at $(( $(date +%s) - S )) s"; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r2_bench_N8_final.err | tail -5
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_N8_final.json').read().strip().splitlines()[-1])
print('cfg2', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['dominant_kernel'], d['ms_dominant_kernel_rank0'], [round(x,3) for x in d['ms_per_frame_render_call_by_rank']], d['e2e']['checksum'])
v=d['workloads']['cfg3']
print('cfg3', v['value'], v['ms_per_step'], v['e2e']['ms_per_step'], v['ms_dominant_kernel_rank0'], [round(x,3) for x in v['ms_per_frame_render_call_by_rank']], v['e2e']['checksum'])
v=d['workloads']['cfg5']
print('cfg5', {k:v[k] for k in ('ms_per_step','value','checksum','prepass_ms')}, v['e2e']['ms_per_step'])
print({k:(round(p['max_ms'],3),p['by_rank_ms']) for k,p in v['phases'].items()})
PY
echo "done at $(( $(date +%s) - S )) s"
