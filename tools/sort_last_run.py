#!/usr/bin/env python
"""Command-line front end of cpp_volume_rendering_b200/sort_last.py: sort-last rc1pass (and, with --renderer vct, rc1pass +
voxel-cone-traced shadows: BASELINE config 5) over N GPUs, one process per GPU.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/sort_last_run.py --res 512 --size 1920 1080

With --check rank 0 also renders the whole volume on its own GPU and compares (2/255, 50 dB)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                # noqa: E402
from cpp_volume_rendering_b200 import sort_last             # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", dest="n", type=int, default=256)
    ap.add_argument("--size", type=int, nargs=2, default=[1280, 720])
    ap.add_argument("--dtype", default="u8")
    ap.add_argument("--tf", default="bonsai")
    ap.add_argument("--volume", default="gauss_noise")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--gen", default="host", choices=["host", "device"], help="device: V-noise generated per brick on the GPU (2048^3 does not fit the host)")
    ap.add_argument("--renderer", default="rc1pass", choices=["rc1pass", "vct"],
                    help="vct: every brick carries the cone-reach halo and its window of the super-voxel pyramid (dist.vct_brick_plan)")
    ap.add_argument("--filter", default="exact", choices=["exact", "hardware"])
    ap.add_argument("--ordered", action="store_true", help="independent segments + ordered over (error <= 0.01) instead of the exact two-pass mode")
    args = ap.parse_args()
    env = bench.Env()
    if env.dist is None:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=env.torch.device("cuda", env.local))
        env.dist = dist
    r = sort_last.run(env, args.n, args.size[0], args.size[1], dtype=args.dtype, renderer=args.renderer, steps=args.steps,
                      filter_mode=args.filter, gen=args.gen, ordered=args.ordered, check=args.check, volume=args.volume, tf=args.tf)
    if env.rank == 0:
        print(json.dumps(r))
    env.dist.destroy_process_group()


if __name__ == "__main__":
    main()
