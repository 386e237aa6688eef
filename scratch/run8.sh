export PYTHONPATH=$PWD
for N in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$N$((RANDOM % 10)) bench.py --gpus $N --steps 20 --warmup 5 2>gpurun_out/err_$N.txt | tail -1 > gpurun_out/bench_cfg2_N${N}.json
  python -c "
import json; d=json.load(open('gpurun_out/bench_cfg2_N${N}.json')); print('N=$N ms/frame', d['ms_per_step'], 'value', d['value'], 'e2e ms', d['e2e']['ms_per_step'], 'kernel rank0', d['ms_per_frame_kernel_only_rank0'], 'checksum', d['e2e']['checksum'], d['e2e']['nonfinite_values'])" || tail -5 gpurun_out/err_$N.txt
done
