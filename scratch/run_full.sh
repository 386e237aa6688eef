export PYTHONPATH=$PWD
timeout 600 python -m pytest tests/test_ebs_gpu.py -m gpu -q -x 2>&1 | tail -5
timeout 1200 python -m pytest tests/test_full_size_gpu.py -m gpu -q -x -s > gpurun_out/full.txt 2>&1
grep -E "^SAT texels|^non-finite|^finite pixels|passed|failed|Error" gpurun_out/full.txt
for o in reference scan; do VRB_SAT_ORDER=$o python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_cfg2_sat_$o.json; python -c "
import json; d=json.load(open('gpurun_out/bench_cfg2_sat_$o.json')); print('$o', d['ms_per_step'], d['e2e']['checksum'], d['e2e']['nonfinite_values'], d['roofline_sat'])"; done
