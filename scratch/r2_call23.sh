#!/bin/bash
S=$(date +%s)
for th in 64 128; do
VRB_MARCH_THREADS=$th timeout 300 python bench.py --workload cfg2 --extras none --steps 5 --warmup 3 --no-cpu-baseline 2> /dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg2 threads $th', d['value'], d['ms_per_step'], d['dominant_kernel'], d['ms_dominant_kernel_rank0'], d['e2e']['checksum'], d['clocks'])"
done
VRB_EBS_KERNEL=coop timeout 300 python bench.py --workload cfg2 --extras none --steps 5 --warmup 3 --no-cpu-baseline 2> /dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg2 coop', d['value'], d['ms_per_step'], d['dominant_kernel'], d['ms_dominant_kernel_rank0'], d['e2e']['checksum'], d['clocks'])"
echo "done at $(( $(date +%s) - S )) s"
