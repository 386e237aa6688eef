#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s)
cap() {  # workload kernel-regex out-name
  timeout 600 ncu --set full --clock-control none -k regex:$2 -c 1 -o /tmp/cap_$3 -f python bench.py --workload $1 --extras none --steps 1 --warmup 3 --no-cpu-baseline > /tmp/cap_$3.log 2>&1
  python profiles/summarize.py /tmp/cap_$3.ncu-rep > gpurun_out/$3.txt
  echo "$3 at $(( $(date +%s) - S )) s: $(grep gpu__time_duration gpurun_out/$3.txt | head -1)"
}
cap cfg1 k_rc1pass r2_final_cfg1_k_rc1pass
cap cfg2 k_ebs_coop r2_final_cfg2_k_ebs_coop
cap cfg3 k_dos_shade r2_final_cfg3_k_dos_shade
cap cfg3 k_list_march r2_final_cfg3_k_list_march
cap cfg4 k_gt_shade r2_final_cfg4_k_gt_shade
cap cfg5-1gpu k_vct_shade r2_final_cfg5-1gpu_k_vct_shade
cap cfg2 k_sat_tiles r2_final_cfg2_k_sat_tiles
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_final_launches_cfg3.csv python bench.py --workload cfg3 --extras none --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_final_launches_cfg2.csv python bench.py --workload cfg2 --extras none --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
echo "done at $(( $(date +%s) - S )) s"
