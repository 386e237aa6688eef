#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/r2_pytest_gpu_final.txt
echo "pytest done at $(( $(date +%s) - S )) s"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/r2_smoke_final.txt
for n in 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2959$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2_bench_N${n}_final.json 2> gpurun_out/r2_bench_N${n}_final.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_N${n}_final.json').read().strip().splitlines()[-1])
print('N=$n cfg2', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'])
v=d['workloads']['cfg3']; print('cfg3', v['value'], v['ms_per_step'])
v=d['workloads']['cfg5']; print('cfg5', v.get('ms_per_step'), v.get('error'))
PY
done
echo "done at $(( $(date +%s) - S )) s"
