#!/usr/bin/env python
"""One small frame of every renderer and pre-pass through the C ABI: the workload compute-sanitizer runs over
(profiles/rNN_sanitizer_*.txt).  Sizes are tiny because memcheck / racecheck slow kernels down 10-100x.

  compute-sanitizer --tool memcheck  python tools/sanitize_small.py
  compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cpp_volume_rendering_b200 as vrb                      # noqa: E402
from cpp_volume_rendering_b200 import capi, synth            # noqa: E402
import bench                                                # noqa: E402


def main():
    n, W, H = 40, 72, 56
    for dt in ("u8", "u16"):
        vox = bench.make_volume(dict(volume="gauss_noise", dtype=dt, n=n))
        bpv = vox.dtype.itemsize
        rgbt, rgba, lut = bench.host_tf_arrays("bonsai", bpv)
        eye, center, up = synth.camera_state(0, n)
        cam = capi.make_camera(eye, center, up, W, H)
        ctx = vrb.Context(0)
        ctx.volume_upload(vox); ctx.tf_upload(rgbt, rgba); ctx.frame_resize(W, H)
        light = capi.default_lighting(light_pos=synth.light_position(n), forward=synth.camera_forward(eye, center))
        diag = float(np.sqrt(3.0) * n)
        sums = {}
        ctx.rc1pass_render(cam, 0.5, count_samples=True); sums["rc1pass"] = float(ctx.frame_read().sum())
        ctx.rc1pass_render(cam, 0.5, count_samples=True, skip_empty=True); sums["rc1pass_skip"] = float(ctx.frame_read().sum())
        for order in ("scan", "reference"):
            ctx.sat_set_order(order); ctx.sat_build(lut)
        p = capi.default_ebs_params(diag); p.count_samples = 1
        ctx.ebs_render(cam, light, p); sums["ebs"] = float(np.nan_to_num(ctx.frame_read()).sum())
        ctx.extcoef_build(1.0, (16, 16, 16))
        occ, _, _ = capi.host_cone_sampler(20.0, 1, 0.5 * diag, 0.35)
        sdw, _, _ = capi.host_cone_sampler(0.5, 0, 0.75 * diag, 1.0)
        ctx.dos_set_cones(occ, sdw)
        p = capi.default_dos_params(0.5, apply_shadow=True); p.count_samples = 1
        ctx.dos_render(cam, light, p); sums["dos"] = float(ctx.frame_read().sum())
        occ_r, sdw_r = capi.host_gt_ray_tables(8, 90.0, 8, 1.0)
        ctx.gt_set_rays(occ_r, sdw_r)
        p = capi.default_gt_params(diag, 8, 8); p.count_samples = 1
        ctx.gt_render(cam, light, p); sums["gt"] = float(ctx.frame_read().sum())
        ctx.vct_build(capi.host_opacity_by_density(synth.TFS["bonsai"], bpv))
        _, _, ms = ctx.vct_info()
        p = capi.default_vct_params(255.0 if bpv == 1 else 65535.0, ms); p.count_samples = 1
        ctx.vct_render(cam, light, p); sums["vct"] = float(ctx.frame_read().sum())
        ctx.gradient_build(1)
        ctx.close()
        print(dt, {k: round(v, 3) for k, v in sums.items()}, flush=True)
    print("sanitize_small: done")


if __name__ == "__main__":
    main()
