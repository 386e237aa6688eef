#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s)
timeout 600 python -m pytest tests/test_dist.py -m gpu -q -x 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_n8_b.json 2> gpurun_out/r2_bench_n8_b.err
echo "bench rc=$? at $(( $(date +%s) - S )) s"; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r2_bench_n8_b.err | tail -5
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_n8_b.json').read().strip().splitlines()[-1])
print('cfg2', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['dominant_kernel'], d['ms_dominant_kernel_rank0'], [round(x,3) for x in d['ms_per_frame_render_call_by_rank']], d['e2e']['checksum'])
v=d['workloads']['cfg3']
print('cfg3', v['value'], v['ms_per_step'], v['e2e']['ms_per_step'], v['ms_dominant_kernel_rank0'], [round(x,3) for x in v['ms_per_frame_render_call_by_rank']], v['e2e']['checksum'])
print('cfg5', d['workloads']['cfg5'])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 tools/sort_last_run.py --res 1024 --size 3840 2160 --dtype u16 --renderer vct --gen device --volume noise --check --steps 5 2> gpurun_out/r2_sl_check.err | tail -1 | tee gpurun_out/r2_sort_last_vct_1024_check_N8.json
grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r2_sl_check.err | tail -5
echo "done at $(( $(date +%s) - S )) s"
