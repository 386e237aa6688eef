#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s)
timeout 75 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tools/sort_last_run.py --renderer vct --res 2048 --size 3840 2160 --dtype u16 --volume noise --gen device --steps 5 > gpurun_out/sort_last_vct_2048_N8_exact.json 2> gpurun_out/sort_last_vct_2048_N8_exact.err
echo "exact rc=$? at $(( $(date +%s) - S )) s"; cat gpurun_out/sort_last_vct_2048_N8_exact.json; tail -3 gpurun_out/sort_last_vct_2048_N8_exact.err
timeout 60 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 tools/sort_last_run.py --renderer vct --res 2048 --size 3840 2160 --dtype u16 --volume noise --gen device --steps 5 --filter hardware > gpurun_out/sort_last_vct_2048_N8_hw.json 2> gpurun_out/sort_last_vct_2048_N8_hw.err
echo "hw rc=$? at $(( $(date +%s) - S )) s"; cat gpurun_out/sort_last_vct_2048_N8_hw.json; tail -3 gpurun_out/sort_last_vct_2048_N8_hw.err
