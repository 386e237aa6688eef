set -x
timeout 600 python -m pytest tests/test_rc1pass_gpu.py tests/test_dos.py -m gpu -q 2>&1 | tail -15
python scratch/exp_rc1.py 2>&1 | tail -14
for f in exact hardware; do python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline --filter $f 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg3', '$f', d['ms_per_step'], d['checksum'] if 'checksum' in d else d['e2e'].get('checksum'), d['samples_per_frame'])"; done
