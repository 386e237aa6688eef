#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s)
timeout 900 python -m pytest tests/test_dos.py tests/test_gt_vct.py tests/test_full_size_gpu.py tests/test_zz_gpu_vs_reference_shader.py tests/test_golden.py -m gpu -q -x 2>&1 | tail -5
echo "pytest done at $(( $(date +%s) - S )) s"
for k in deferred ray; do
VRB_VCT_KERNEL=$k timeout 300 python bench.py --workload cfg5-1gpu --steps 10 --warmup 3 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg5-1gpu $k', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['dominant_kernel'], d['ms_dominant_kernel_rank0'], d['e2e']['checksum'])"
done
timeout 120 python tools/sanitize_small.py | tail -3
echo "plain done at $(( $(date +%s) - S )) s"
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_small.py > gpurun_out/r2_sanitizer_memcheck.txt 2>&1; tail -6 gpurun_out/r2_sanitizer_memcheck.txt
echo "memcheck done at $(( $(date +%s) - S )) s"
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_small.py > gpurun_out/r2_sanitizer_racecheck.txt 2>&1; tail -6 gpurun_out/r2_sanitizer_racecheck.txt
echo "done at $(( $(date +%s) - S )) s"
