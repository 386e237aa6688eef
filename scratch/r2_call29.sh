#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s)
timeout 600 python -m pytest tests/test_ebs_gpu.py -m gpu -q -x 2>&1 | tail -3
echo "pytest at $(( $(date +%s) - S )) s"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tools/sat_sharded_run.py --res 512 2> gpurun_out/sat_sh.err | tail -1 | tee gpurun_out/r2_sat_sharded_512_N2.json
grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/sat_sh.err | tail -5
timeout 300 python bench.py --workload cfg2 --extras none --steps 5 --warmup 3 --no-cpu-baseline 2> /dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg2', d['value'], d['ms_per_step'], d['dominant_kernel'], d['e2e']['checksum']); print({k:d['roofline_sat'][k] for k in ('ms','frac','reference_order_ms','reference_order_frac','reference_order_call_ms')})"
echo "done at $(( $(date +%s) - S )) s"
