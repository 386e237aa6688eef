set -x
export PYTHONPATH=$PWD
timeout 600 python -m pytest tests/test_rc1pass_gpu.py tests/test_gt_vct.py -m gpu -q -x 2>&1 | tail -5
python bench.py --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_cfg2_pipe.json
python -c "import json; d=json.load(open('gpurun_out/bench_cfg2_pipe.json')); print('cfg2', d['ms_per_step'], d['e2e'])"
cp cpp_volume_rendering_b200/libvrb200.so /tmp/keep.so
for v in 1 2 4; do
  cp scratch/so/libvrb200_ilp$v.so cpp_volume_rendering_b200/libvrb200.so
  for f in exact hardware; do
    python bench.py --workload cfg4 --steps 2 --warmup 3 --no-cpu-baseline --filter $f 2>/dev/null | tail -1 > gpurun_out/bench_cfg4_ilp${v}_$f.json
    python -c "import json; d=json.load(open('gpurun_out/bench_cfg4_ilp${v}_$f.json')); print('cfg4 ilp$v $f', d['ms_per_step'], d['e2e'].get('checksum'), d['secondary_units_per_frame'])"
  done
done
cp /tmp/keep.so cpp_volume_rendering_b200/libvrb200.so
