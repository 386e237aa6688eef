#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s)
timeout 900 python -m pytest tests/test_rc1pass_gpu.py tests/test_golden.py tests/test_dist.py tests/test_zz_gpu_vs_reference_shader.py tests/test_host_cpu.py tests/test_multiscaling.py tests/test_eval_harness.py -m gpu -q -x 2>&1 | tail -3
timeout 100 python tests/gpu_random_sweep.py --seed 3 --scenes 8 2>&1 | tail -1
for k in list ray; do
VRB_RC1_KERNEL=$k timeout 300 python bench.py --workload cfg1 --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg1 $k', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['dominant_kernel'], d['ms_dominant_kernel_rank0'], d['samples_per_frame'], d['e2e']['checksum'])"
done
timeout 600 ncu --set full --clock-control none -k regex:k_list_march -c 1 -o /tmp/cap_rc1 -f python bench.py --workload cfg1 --extras none --steps 1 --warmup 3 --no-cpu-baseline > /tmp/cap_rc1.log 2>&1
python profiles/summarize.py /tmp/cap_rc1.ncu-rep > gpurun_out/r2_final_cfg1_k_list_march_rc1pass.txt; grep -v "launch__\|cycles_elapsed\|requests_pipe\|sectors_pipe" gpurun_out/r2_final_cfg1_k_list_march_rc1pass.txt | head -16
echo "done at $(( $(date +%s) - S )) s"
