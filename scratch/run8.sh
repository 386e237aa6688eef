set -x
for N in 8 4; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 10 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_cfg2_N$N.json
  python -c "import json; d=json.load(open('gpurun_out/bench_cfg2_N$N.json')); print('N=$N ms/frame', d['ms_per_step'], 'value', d['value'], 'e2e ms', d['e2e']['ms_per_step'], 'kernel rank0', d['ms_per_frame_kernel_only_rank0'])"
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tools/sort_last_run.py --res 1024 --size 3840 2160 --dtype u16 --volume noise --gen device --check 2>&1 | tail -1 | tee gpurun_out/sort_last_1024_N8.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 tools/sort_last_run.py --res 2048 --size 3840 2160 --dtype u16 --volume noise --gen device 2>&1 | tail -1 | tee gpurun_out/sort_last_2048_N8.json
