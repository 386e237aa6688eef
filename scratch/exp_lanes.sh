export PYTHONPATH=$PWD
timeout 900 python -m pytest tests/test_ebs_gpu.py tests/test_dist.py tests/test_full_size_gpu.py -m gpu -q -x 2>&1 | tail -3
for CAP in 0 8 16 32; do echo "cap $CAP: $(VRB_EBS_CAP=$CAP python scratch/exp_partition2.py 2>&1 | tail -1)"; done
