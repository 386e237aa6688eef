#!/bin/bash
S=$(date +%s)
for w in cfg2 cfg3; do
VRB_TRACE=1 timeout 300 python bench.py --workload $w --extras none --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/trace_$w.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w', d['value'], d['ms_per_step'], d['ms_per_frame_render_call_rank0'], d['dominant_kernel'], d['ms_dominant_kernel_rank0'])"
grep "vrb trace" gpurun_out/trace_$w.err | sed -n '3,4p;12,14p'
done
echo "done at $(( $(date +%s) - S )) s"
