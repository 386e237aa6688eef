export PYTHONPATH=$PWD
export VRB_EBS_LANES=4
timeout 600 ncu --set full --clock-control none -k regex:k_ebs_coop -s 2 -c 1 -f -o gpurun_out/prof_cfg2_part8_m4 python scratch/exp_part_ncu.py 8 > gpurun_out/part8.log 2>&1
python scratch/exp_tiles.py 2>&1 | tail -6
