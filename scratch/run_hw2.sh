set -x
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
python scratch/exp_rc1.py 2>&1 | tail -14
for wl in cfg3 cfg4 cfg5-1gpu cfg1; do for f in exact hardware; do python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline --filter $f 2>/dev/null | tail -1 > gpurun_out/bench_${wl}_$f.json; python -c "import sys,json; d=json.load(open('gpurun_out/bench_${wl}_$f.json')); print('$wl', '$f', d['ms_per_step'], d['e2e'].get('checksum'), d['samples_per_frame'], d['secondary_units_per_frame'])"; done; done
