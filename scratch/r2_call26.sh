#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s)
timeout 600 python -m pytest tests/test_dist.py tests/test_gt_vct.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/sort_last_run.py --res 1024 --size 3840 2160 --dtype u16 --renderer vct --gen device --volume noise --steps 5 --check 2> gpurun_out/r2_sl2.err | tail -1 | tee gpurun_out/r2_sort_last_vct_1024_N2.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('ms_per_step','max_abs_err','parity_ok','samples_per_frame','samples_per_frame_single_gpu','checksum')}); print({k:(v['max_ms'],v['by_rank_ms']) for k,v in d['phases'].items()})"
grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r2_sl2.err | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 tools/sort_last_run.py --res 512 --size 1920 1080 --dtype u8 --renderer rc1pass --volume gauss_noise --steps 5 --check 2> gpurun_out/r2_sl3.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('ms_per_step','max_abs_err','parity_ok','samples_per_frame','samples_per_frame_single_gpu')})"
echo "done at $(( $(date +%s) - S )) s"
