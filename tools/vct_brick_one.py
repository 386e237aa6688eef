#!/usr/bin/env python
"""One rank's share of BASELINE config 5 (2048^3 u16, rc1pass + voxel-cone-traced shadows, 3840x2160, 2x2x2 bricks) on
ONE GPU: the brick window (owned cells + cone-reach halo) is generated on the device, its pyramid and the LUT are built,
and the two passes of the exact sort-last mode (opacity pre-pass, shaded pass) are timed with CUDA events.

  python tools/vct_brick_one.py --res 2048 --size 3840 2160 --world 8 --rank 7

This is a single-GPU measurement of the per-rank kernels, NOT a multi-GPU frame time: the front bricks' opacity is taken
as zero (nothing in front), the deviation range is the window's own, and no compositing exchange runs.  The full run is
tools/sort_last_run.py --renderer vct under torchrun."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cpp_volume_rendering_b200 as vrb                      # noqa: E402
from cpp_volume_rendering_b200 import capi, synth, dist as vdist   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", dest="n", type=int, default=2048)
    ap.add_argument("--size", type=int, nargs=2, default=[3840, 2160])
    ap.add_argument("--dtype", default="u16")
    ap.add_argument("--tf", default="bonsai")
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--rank", type=int, default=-1, help="-1: the brick nearest the eye")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--filter", default="exact", choices=["exact", "hardware"])
    args = ap.parse_args()
    n = args.n; W, H = args.size
    bpv = 1 if args.dtype == "u8" else 2
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    h = capi.load_host()
    rgb, a = synth.TFS[args.tf]
    import bench
    rgbt, rgba, _ = bench.host_tf_arrays(args.tf, bpv)
    opc = capi.host_opacity_by_density(synth.TFS[args.tf], bpv)
    eye, center, up = synth.camera_state(0, n)
    cam = capi.make_camera(eye, center, up, W, H)
    light = capi.default_lighting(light_pos=synth.light_position(n))
    prm = capi.default_vct_params(255.0 if bpv == 1 else 65535.0, 1.0, 0.5)
    plans, n_levels, halo = vdist.vct_brick_plan((n, n, n), args.world, prm)
    order = vdist.visibility_order(plans, eye, (n, n, n))
    rank = order[0] if args.rank < 0 else args.rank
    p = plans[rank]
    brick = capi.Brick()
    brick.global_dims[:] = [n, n, n]; brick.origin[:] = list(p["origin"]); brick.owned[:] = list(p["owned"])
    brick.ghost_lo[:] = list(p["ghost_lo"]); brick.ghost_hi[:] = list(p["ghost_hi"])
    ctx = vrb.Context(0)
    stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
    t0 = time.perf_counter()
    blk = synth.volume_noise_torch(n, p["slices_zyx"], args.dtype, device=dev)
    torch.cuda.synchronize()
    gen_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    ctx.volume_upload_device(blk.data_ptr(), blk.shape[2], blk.shape[1], blk.shape[0], bpv)
    ctx.synchronize()
    upload_s = time.perf_counter() - t0
    window = [int(blk.shape[2]), int(blk.shape[1]), int(blk.shape[0])]
    del blk
    torch.cuda.empty_cache()
    ctx.tf_upload(rgbt, rgba); ctx.frame_resize(W, H)
    t0 = time.perf_counter()
    lmax = ctx.sv_build_brick(brick, n_levels)
    ctx.synchronize()
    pyr_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    ctx.preint_build(opc, lmax)
    ctx.synchronize()
    lut_s = time.perf_counter() - t0
    prm.volume_max_stddev = np.float32(lmax)
    ctx.set_filter(args.filter)

    def timed(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.steps

    ms_alpha = timed(lambda: ctx.vct_render_brick(cam, light, prm, brick, capi.BRICK_ALPHA))
    ms_exact = timed(lambda: ctx.vct_render_brick(cam, light, prm, brick, capi.BRICK_EXACT, []))
    # the same window marched by the unlit sort-last kernel (k_rc1pass_brick: opacity pre-pass and exact pass)
    ms_rc_alpha = timed(lambda: ctx.rc1pass_brick_alpha(cam, brick, 0.5))
    ms_rc_exact = timed(lambda: ctx.rc1pass_render_brick_exact(cam, brick, [], 0.5))
    prm.count_samples = 1
    ctx.vct_render_brick(cam, light, prm, brick, capi.BRICK_EXACT, [])
    samples, taps = ctx.last_sample_count, ctx.last_aux_count
    free_b, total_b = torch.cuda.mem_get_info()
    print(json.dumps({
        "what": "one rank's share of config 5 on one GPU (per-rank kernels only; see the docstring)",
        "volume": f"{n}^3 {args.dtype} V-noise, device-generated", "frame": [W, H], "bricks": args.world, "rank": rank,
        "window_voxels": window, "halo_voxels": halo, "pyramid_levels_per_brick": n_levels, "filter": args.filter,
        "generate_s": gen_s, "upload_s": upload_s, "pyramid_s": pyr_s, "lut_s": lut_s, "max_stddev_window": lmax,
        "ms_alpha_pass": ms_alpha, "ms_shaded_pass": ms_exact, "ms_both": ms_alpha + ms_exact,
        "ms_rc1pass_alpha_pass": ms_rc_alpha, "ms_rc1pass_exact_pass": ms_rc_exact,
        "samples_owned": samples, "cone_taps": taps,
        "gsamples_per_s_shaded_pass": samples / (ms_exact * 1e-3) / 1e9, "gtaps_per_s": taps / (ms_exact * 1e-3) / 1e9,
        "hbm_used_gb": (total_b - free_b) / 1e9}))
    ctx.close()


if __name__ == "__main__":
    main()
