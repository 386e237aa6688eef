#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s)
timeout 400 python scratch/r2_dos_ab.py 2>&1 | tail -12 | tee gpurun_out/r2_dos_ab.txt
echo "ab done at $(( $(date +%s) - S )) s"
timeout 300 python -m pytest tests/test_dos.py tests/test_dist.py tests/test_golden.py tests/test_zz_gpu_vs_reference_shader.py -m gpu -q -x 2>&1 | tail -5
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dos_shade -c 1 -o gpurun_out/r2_k_dos_shade_cfg3 -f python scratch/r2_dos_ab.py quick > gpurun_out/ncu_dos.log 2>&1
tail -3 gpurun_out/ncu_dos.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_dos -c 12 --csv --log-file gpurun_out/r2_launches_dos_cfg3.csv python scratch/r2_dos_ab.py quick > /dev/null 2>&1
grep -o 'k_dos[a-z_]*.*' gpurun_out/r2_launches_dos_cfg3.csv | cut -c1-30,120-200 | tail -9
echo "ncu done at $(( $(date +%s) - S )) s"
