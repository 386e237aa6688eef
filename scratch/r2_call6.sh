#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --cfg5-res 1024 > gpurun_out/r2_bench_n2_a.json 2> gpurun_out/r2_bench_n2_a.err
echo "bench rc=$? at $(( $(date +%s) - S )) s"; tail -c 1500 gpurun_out/r2_bench_n2_a.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n2_a.json').read().strip().splitlines()[-1])
print('cfg2', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['dominant_kernel'], d['ms_dominant_kernel_rank0'], d['roofline']['frac'])
for k,v in d.get('workloads',{}).items():
    print(k, {a:b for a,b in v.items() if a not in ('roofline','roofline_hbm','roofline_ldg16','config','init')})
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 | tail -c 900
echo "done at $(( $(date +%s) - S )) s"
