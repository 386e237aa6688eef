export PYTHONPATH=$PWD
for N in 8 4; do for a in p2p reduce; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$N$((RANDOM % 10)) bench.py --gpus $N --steps 20 --warmup 5 --assemble $a 2>gpurun_out/err_$a.txt | tail -1 > gpurun_out/bench_cfg2_N${N}_$a.json
  python -c "
import json; d=json.load(open('gpurun_out/bench_cfg2_N${N}_$a.json')); print('N=$N $a ms/frame', d['ms_per_step'], 'value', d['value'], 'e2e ms', d['e2e']['ms_per_step'], 'kernel rank0', d['ms_per_frame_kernel_only_rank0'], 'checksum', d['e2e']['checksum'], d['e2e']['nonfinite_values'])" || tail -5 gpurun_out/err_$a.txt
done; done
