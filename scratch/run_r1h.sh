#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29511 tools/sort_last_run.py --renderer vct --res 256 --size 640 360 --dtype u16 --volume noise --gen device --steps 3 --check > gpurun_out/sort_last_vct_w1.json 2> gpurun_out/sort_last_vct_w1.err
echo "rc=$? elapsed $(( $(date +%s) - S )) s"
cat gpurun_out/sort_last_vct_w1.json; tail -5 gpurun_out/sort_last_vct_w1.err
