export PYTHONPATH=$PWD
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for f in hardware exact; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gt -c 1 -f -o gpurun_out/prof_cfg4_k_gt_$f python bench.py --workload cfg4 --steps 1 --warmup 1 --no-cpu-baseline --filter $f > gpurun_out/prof_cfg4_$f.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dos -c 1 -f -o gpurun_out/prof_cfg3_k_dos_hardware python bench.py --workload cfg3 --steps 1 --warmup 1 --no-cpu-baseline --filter hardware > gpurun_out/prof_cfg3_hw.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rc1pass -c 1 -f -o gpurun_out/prof_cfg1_k_rc1pass_hardware python bench.py --workload cfg1 --steps 1 --warmup 1 --no-cpu-baseline --filter hardware > gpurun_out/prof_cfg1_hw.log 2>&1
grep -o '"samples_per_frame": [0-9]*' gpurun_out/prof_cfg1_hw.log gpurun_out/prof_cfg3_hw.log gpurun_out/prof_cfg4_exact.log
