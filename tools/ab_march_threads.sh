#!/bin/bash
# A/B of the march CTA size (VRB_MARCH_THREADS) on the deferred workloads: prints ms per frame and the march kernel's share.
for w in cfg3 cfg4 cfg5-1gpu; do
for th in 128 256; do
VRB_TRACE=1 VRB_MARCH_THREADS=$th timeout 300 python bench.py --workload $w --extras none --steps 5 --warmup 3 --no-cpu-baseline 2> /tmp/tr.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w threads $th', round(d['ms_per_step'],3), d['dominant_kernel'], round(d['ms_dominant_kernel_rank0'],3), d['e2e']['checksum'])"
grep "vrb trace" /tmp/tr.err | tail -1
done
done
