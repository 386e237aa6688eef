#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s)
for tile in 16 32; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2952$((tile/16)) bench.py --gpus 8 --steps 10 --warmup 3 --tile $tile --cfg5-res 512 > gpurun_out/r2_bench_n8_t$tile.json 2> gpurun_out/r2_bench_n8_t$tile.err
echo "bench tile=$tile rc=$? at $(( $(date +%s) - S )) s"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_n8_t$tile.json').read().strip().splitlines()[-1])
print('cfg2', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['dominant_kernel'], d['ms_dominant_kernel_rank0'], [round(x,3) for x in d['ms_per_frame_render_call_by_rank']], d['e2e']['checksum'])
v=d['workloads']['cfg3']
print('cfg3', v['value'], v['ms_per_step'], v['e2e']['ms_per_step'], v['ms_dominant_kernel_rank0'], [round(x,3) for x in v['ms_per_frame_render_call_by_rank']], v['e2e']['checksum'])
PY
done
echo "done at $(( $(date +%s) - S )) s"
