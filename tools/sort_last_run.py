#!/usr/bin/env python
"""Sort-last rc1pass (and, with --renderer vct, rc1pass + voxel-cone-traced shadows: BASELINE config 5) over N GPUs, one
process per GPU: bricks + compositing through CUDA-IPC peer loads.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/sort_last_run.py --n 512 --size 1920 1080

Every rank owns one brick of a seeded synthetic volume, renders its partial frame (vrb_rc1pass_render_brick), publishes
the buffer with vrb_ipc_export; every rank then composites its strip of the image from ALL partial frames in visibility
order with ONE kernel that loads the peers' pixels over NVLink (vrb_composite_ordered), and the strips are gathered on
rank 0.  With --check rank 0 also renders the whole volume on its own GPU and compares (2/255, 50 dB)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cpp_volume_rendering_b200 as vrb                      # noqa: E402
from cpp_volume_rendering_b200 import capi, synth, dist as vdist   # noqa: E402
import bench                                                # noqa: E402


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
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = args.n; W, H = args.size
    H -= H % world                                            # strips of equal height
    wl = dict(volume=args.volume, dtype=args.dtype, n=n)
    vox = bench.make_volume(wl) if args.gen == "host" else None   # host: every rank generates the same seeded volume, keeps its brick
    bpv = 1 if args.dtype == "u8" else 2
    rgbt, rgba, _ = bench.host_tf_arrays(args.tf, bpv)
    eye, center, up = synth.camera_state(0, n)
    cam = capi.make_camera(eye, center, up, W, H)
    vct = args.renderer == "vct"
    n_levels = halo = 0
    if vct:
        opc = capi.host_opacity_by_density(synth.TFS[args.tf], bpv)
        prm = capi.default_vct_params(255.0 if bpv == 1 else 65535.0, 1.0, 0.5)      # max_stddev filled in after the pre-pass
        light = capi.default_lighting(light_pos=synth.light_position(n))
        plans, n_levels, halo = vdist.vct_brick_plan((n, n, n), world, prm)
    else:
        plans = vdist.brick_plan((n, n, n), world)
    order = vdist.visibility_order(plans, eye, (n, n, n))
    p = plans[rank]
    brick = capi.Brick()
    brick.global_dims[:] = [n, n, n]; brick.origin[:] = list(p["origin"]); brick.owned[:] = list(p["owned"])
    brick.ghost_lo[:] = list(p["ghost_lo"]); brick.ghost_hi[:] = list(p["ghost_hi"])
    ctx = vrb.Context(local)
    stream = torch.cuda.Stream(device=local); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
    if args.gen == "host":
        ctx.volume_upload(np.ascontiguousarray(vox[p["slices_zyx"]]))
    else:
        blk = synth.volume_noise_torch(n, p["slices_zyx"], args.dtype, device=torch.device("cuda", local))
        torch.cuda.synchronize()
        ctx.volume_upload_device(blk.data_ptr(), blk.shape[2], blk.shape[1], blk.shape[0], bpv)
        ctx.synchronize()
        del blk
        torch.cuda.empty_cache()
    ctx.tf_upload(rgbt, rgba); ctx.frame_resize(W, H)
    prepass_ms = None
    if vct:
        # pre-pass: window pyramid per brick, the levels above from the gathered last window level, one LUT for all
        torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
        lmax = ctx.sv_build_brick(brick, n_levels)
        gathered = [None] * world
        dist.all_gather_object(gathered, (lmax, ctx.sv_top_means(brick)))
        gmax = vdist.vct_global_max_stddev(ctx, [g[0] for g in gathered], [g[1] for g in gathered], (n, n, n), n_levels)
        ctx.preint_build(opc, gmax)
        ctx.synchronize(); dist.barrier(); prepass_ms = (time.perf_counter() - t0) * 1e3
        prm.volume_max_stddev = np.float32(gmax)
        ctx.set_filter(args.filter)
        ctx.vct_render_brick(cam, light, prm, brick, capi.BRICK_ALPHA)   # allocates the partial frame
    else:
        ctx.rc1pass_render_brick(cam, brick, 0.5)             # allocates the partial frame
    ctx.synchronize()
    my_ptr = ctx.partial_device_ptr()
    my_alpha = ctx.brick_alpha_device_ptr()
    handles = [None] * world
    dist.all_gather_object(handles, (ctx.ipc_export(my_ptr), ctx.ipc_export(my_alpha)))
    ptrs = [my_ptr if r == rank else ctx.ipc_import(handles[r][0]) for r in range(world)]
    aptrs = [my_alpha if r == rank else ctx.ipc_import(handles[r][1]) for r in range(world)]
    front = [aptrs[r] for r in order[:order.index(rank)]]
    r0, r1 = vdist.strip_rows(H, world)[rank]
    fptr, _, _ = ctx.frame_device_ptr()

    class _Wrap:
        __cuda_array_interface__ = {"shape": (H, W, 4), "typestr": "<f2", "data": (fptr, False), "version": 2}
    frame_t = torch.as_tensor(_Wrap(), device=torch.device("cuda", local))
    strips = [torch.empty((r1 - r0, W, 4), dtype=torch.float16, device="cuda") for _ in range(world)] if rank == 0 else None
    token = torch.zeros(1, device="cuda")

    def frame():
        if vct:
            if args.ordered:
                ctx.vct_render_brick(cam, light, prm, brick, capi.BRICK_SEGMENT)
                dist.all_reduce(token)
                ctx.composite_ordered([ptrs[r] for r in order], r0, r1 - r0)
            else:
                ctx.vct_render_brick(cam, light, prm, brick, capi.BRICK_ALPHA)
                dist.all_reduce(token)
                ctx.vct_render_brick(cam, light, prm, brick, capi.BRICK_EXACT, front)
                dist.all_reduce(token)
                ctx.composite_sum(ptrs, r0, r1 - r0)
        elif args.ordered:
            ctx.rc1pass_render_brick(cam, brick, 0.5)
            dist.all_reduce(token)                            # every partial frame is complete before anyone reads it
            ctx.composite_ordered([ptrs[r] for r in order], r0, r1 - r0)
        else:
            ctx.rc1pass_brick_alpha(cam, brick, 0.5)          # pass 1: opacity of my segment
            dist.all_reduce(token)
            ctx.rc1pass_render_brick_exact(cam, brick, front, 0.5)   # pass 2 reads the front bricks' opacity over NVLink
            dist.all_reduce(token)
            ctx.composite_sum(ptrs, r0, r1 - r0)
        dist.gather(frame_t[r0:r1], strips, dst=0)
        dist.all_reduce(token)                                # nobody overwrites a partial frame that is still being read

    for _ in range(3):
        frame()
    torch.cuda.synchronize(); dist.barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        frame()
    e1.record(stream)
    torch.cuda.synchronize(); dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device="cuda", dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    result = None
    if rank == 0:
        img = torch.cat(strips, 0).float().cpu().numpy()
        result = {"sort_last": True, "n_gpus": world, "volume": f"{n}^3 {args.dtype} ({args.gen}-generated)", "frame": [W, H], "ms_per_frame": float(ms[0]),
                  "brick_grid": vdist.split_counts(world), "visibility_order": order,
                  "mode": "ordered-over" if args.ordered else "exact two-pass", "renderer": args.renderer, "filter": args.filter}
        if vct:
            result.update(pyramid_levels_per_brick=n_levels, halo_voxels=halo, window=[int(s.stop - s.start) for s in p["slices_zyx"]][::-1],
                          prepass_ms=prepass_ms, max_stddev=float(prm.volume_max_stddev))
        if args.check:
            full = vrb.Context(local)
            if args.gen == "host":
                full.volume_upload(vox)
            else:
                fv = synth.volume_noise_torch(n, None, args.dtype, device=torch.device("cuda", local))
                torch.cuda.synchronize()
                full.volume_upload_device(fv.data_ptr(), n, n, n, bpv); full.synchronize(); del fv
            full.tf_upload(rgbt, rgba); full.frame_resize(W, H)
            if vct:
                full.vct_build(opc)
                assert np.float32(full.vct_info()[2]) == np.float32(prm.volume_max_stddev), (full.vct_info()[2], prm.volume_max_stddev)
                prm.count_samples = 1
                full.set_filter(args.filter)
                full.vct_render(cam, light, prm)
            else:
                full.rc1pass_render(cam, 0.5, count_samples=True)
            want = full.frame_read()
            err = float(np.abs(img - want).max())
            mse = float(np.mean((img.astype(np.float64) - want) ** 2))
            result.update(max_abs_err=err, psnr_db=(float("inf") if mse == 0 else float(10 * np.log10(1.0 / mse))),
                          parity_ok=bool(err <= 2.0 / 255.0), samples_per_frame=full.last_sample_count)
            full.close()
        print(json.dumps(result))
    for r in range(world):
        if r != rank:
            ctx.ipc_close(ptrs[r]); ctx.ipc_close(aptrs[r])
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
