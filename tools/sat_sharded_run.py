#!/usr/bin/env python
"""Sharded SAT build over N GPUs (dist.sat_build_sharded), checked against the single-GPU scan build on rank 0 and timed.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/sat_sharded_run.py --res 512"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                       # noqa: E402
import cpp_volume_rendering_b200 as vrb                            # noqa: E402
from cpp_volume_rendering_b200 import dist as vdist               # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", dest="n", type=int, default=256)
    ap.add_argument("--dtype", default="u8")
    args = ap.parse_args()
    env = bench.Env()
    torch, dist, rank, world = env.torch, env.dist, env.rank, env.world
    assert dist is not None, "run under torch.distributed.run with at least 2 ranks"
    n = args.n
    vox = bench.make_volume(dict(volume="noise", dtype=args.dtype, n=n))
    _, _, lut = bench.host_tf_arrays("bonsai", vox.dtype.itemsize)
    ctx = vrb.Context(env.local)
    ctx.set_stream(env.stream.cuda_stream)
    ctx.volume_upload(vox)
    dev = torch.device("cuda", env.local)
    vdist.sat_build_sharded(ctx, lut, n, rank, world, dev)        # warm-up (allocations)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    vdist.sat_build_sharded(ctx, lut, n, rank, world, dev)
    torch.cuda.synchronize(); dist.barrier()
    ms = env.max_over_ranks((time.perf_counter() - t0) * 1e3)
    got = ctx.sat_read(vox.shape)
    res = None
    if rank == 0:
        ctx.sat_set_order("scan")
        ctx.sat_build(lut)
        t0 = time.perf_counter(); ctx.sat_build(lut); ctx.synchronize(); single_ms = (time.perf_counter() - t0) * 1e3
        want = ctx.sat_read(vox.shape)
        ulp = np.spacing(np.abs(want).astype(np.float32))
        diff = np.abs(got - want)
        res = {"sat_sharded": True, "n_gpus": world, "volume": f"{n}^3 {args.dtype}", "ms_sharded_call": ms, "ms_single_gpu_scan_call": single_ms,
               "texels_differing": int((diff > 0).sum()), "max_diff_in_ulps": float((diff / np.maximum(ulp, 1e-30)).max()),
               "within_one_ulp": bool(np.all(diff <= ulp))}
        print(json.dumps(res))
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
