#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/sort_last_run.py --res 1024 --size 3840 2160 --dtype u16 --renderer vct --gen device --volume noise --steps 5 2> gpurun_out/r2_sl2.err | tail -1 | tee gpurun_out/r2_sort_last_vct_1024_N2.json
grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r2_sl2.err | tail -5
echo "sl done at $(( $(date +%s) - S )) s"
timeout 600 python -m pytest tests/test_rc1pass_gpu.py tests/test_golden.py tests/test_dist.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python bench.py --workload cfg1 --steps 20 --warmup 5 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg1', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['dominant_kernel'], d['ms_dominant_kernel_rank0'], d['e2e']['checksum'])"
VRB_VOL_QUADS=0 timeout 300 python bench.py --workload cfg1 --steps 20 --warmup 5 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg1 linear', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['dominant_kernel'], d['ms_dominant_kernel_rank0'], d['e2e']['checksum'])"
echo "done at $(( $(date +%s) - S )) s"
