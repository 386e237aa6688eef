#!/bin/bash
S=$(date +%s)
for rf in 32 8; do
VRB_MARCH_REFILL=$rf timeout 300 python bench.py --workload cfg2 --extras none --steps 5 --warmup 3 --no-cpu-baseline 2> /dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg2 refill $rf', d['value'], d['ms_per_step'], d['dominant_kernel'], d['ms_dominant_kernel_rank0'], d['e2e']['checksum'])"
done
VRB_MARCH_REFILL=32 timeout 300 python bench.py --workload cfg3 --extras none --steps 5 --warmup 3 --no-cpu-baseline 2> /dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg3 refill 32', d['value'], d['ms_per_step'], d['dominant_kernel'], d['ms_dominant_kernel_rank0'], d['e2e']['checksum'])"
echo "done at $(( $(date +%s) - S )) s"
