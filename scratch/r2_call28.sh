#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s)
timeout 600 python -m pytest tests/test_ebs_gpu.py tests/test_full_size_gpu.py::test_config2_full_size_ebs_matches_oracle -m gpu -q -x 2>&1 | tail -3
echo "pytest at $(( $(date +%s) - S )) s"
timeout 300 python bench.py --workload cfg2 --extras none --steps 5 --warmup 3 --no-cpu-baseline 2> /dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg2', d['value'], d['ms_per_step'], d['dominant_kernel'], d['e2e']['checksum']); print(d['roofline_sat'])"
VRB_SAT_REFERENCE=planes timeout 300 python bench.py --workload cfg2 --extras none --steps 5 --warmup 3 --no-cpu-baseline 2> /dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg2 planes', d['e2e']['checksum'], d['roofline_sat']['reference_order_ms'])"
echo "done at $(( $(date +%s) - S )) s"
