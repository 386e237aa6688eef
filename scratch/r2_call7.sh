#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s)
timeout 900 python -m pytest tests/test_dos.py tests/test_gt_vct.py tests/test_full_size_gpu.py tests/test_zz_gpu_vs_reference_shader.py tests/test_golden.py -m gpu -q -x 2>&1 | tail -15
echo "pytest done at $(( $(date +%s) - S )) s"
for w in cfg3 cfg4; do
timeout 300 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_b_$w.json 2> gpurun_out/r2_b_$w.err || tail -5 gpurun_out/r2_b_$w.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_b_$w.json').read().strip().splitlines()[-1])
print('$w', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['dominant_kernel'], d['ms_dominant_kernel_rank0'], d['roofline']['frac'], d.get('roofline_ldg16',{}).get('frac'), d['samples_per_frame'], d['secondary_units_per_frame'])
PY
done
echo "done at $(( $(date +%s) - S )) s"
