#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s)
n=4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2959$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2_bench_N${n}_final.json 2> gpurun_out/r2_bench_N${n}_final.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_N${n}_final.json').read().strip().splitlines()[-1])
print('N=$n cfg2', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'])
v=d['workloads']['cfg3']; print('cfg3', v['value'], v['ms_per_step'])
v=d['workloads']['cfg5']; print('cfg5', v.get('ms_per_step'), v.get('error'))
PY
echo "done at $(( $(date +%s) - S )) s"
