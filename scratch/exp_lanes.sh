for L in 1 2 4 8; do echo "== LANES=$L"; VRB_EBS_LANES=$L python -m pytest tests/test_ebs_gpu.py -m gpu -x -q -k "ebs" 2>&1 | tail -1; VRB_EBS_LANES=$L python scratch/exp_partition2.py; done
