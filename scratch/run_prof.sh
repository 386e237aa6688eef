set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.txt
cat gpurun_out/pytest_gpu.txt
for W in cfg2:k_ebs_coop cfg3:k_dos cfg5-1gpu:k_vct cfg1:k_rc1pass; do
  wl=${W%%:*}; k=${W##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/prof_${wl}_$k python bench.py --workload $wl --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/prof_${wl}.log 2>&1
done
for wl in cfg1 cfg3 cfg5-1gpu; do python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_$wl.json; done
python bench.py 2>/dev/null | tail -1 > gpurun_out/bench_cfg2.json
cat gpurun_out/bench_cfg2.json
