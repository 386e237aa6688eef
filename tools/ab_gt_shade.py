"""A/B of the rc1pcrtgt shade variants at cfg4 (256^3 u8 V-boxes, 64+64 rays, 1080p): ms/frame, counters, frame checksum."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import cpp_volume_rendering_b200 as vrb
from cpp_volume_rendering_b200 import capi, synth
import bench

torch.cuda.set_device(0)
STREAM = torch.cuda.Stream(device=0)
torch.cuda.set_stream(STREAM)
wl = bench.WORKLOADS["cfg4"]
vox = bench.make_volume(wl)
n = wl["n"]
rgbt, rgba, lut = bench.host_tf_arrays(wl["tf"], 1)
eye, center, up = synth.camera_state(0, n)
cam = capi.make_camera(eye, center, up, wl["W"], wl["H"])
occ_r, sdw_r = capi.host_gt_ray_tables(64, 90.0, 64, 1.0)
fwd = synth.camera_forward(eye, center)
light = capi.default_lighting(light_pos=synth.light_position(n), forward=tuple(-f for f in fwd))
variants = [dict(VRB_GT_SHADE="entry", VRB_GT_ILP="1", VRB_VOL_QUADS="1"), dict(VRB_GT_SHADE="entry", VRB_GT_ILP="2", VRB_VOL_QUADS="1"),
            dict(VRB_GT_SHADE="entry", VRB_GT_ILP="4", VRB_VOL_QUADS="1"), dict(VRB_GT_SHADE="entry", VRB_GT_ILP="2", VRB_VOL_QUADS="0"),
            dict(VRB_GT_SHADE="task", VRB_VOL_QUADS="1"), dict(VRB_GT_SHADE="task", VRB_VOL_QUADS="0")]
if len(sys.argv) > 1:
    variants = [variants[int(a)] for a in sys.argv[1:]]
ref = None
for v in variants:
    for k in ("VRB_GT_SHADE", "VRB_GT_ILP", "VRB_VOL_QUADS", "VRB_GT_KERNEL"):
        os.environ.pop(k, None)
    os.environ.update(v)
    ctx = vrb.Context(0)
    ctx.set_stream(STREAM.cuda_stream)
    ctx.volume_upload(vox); ctx.tf_upload(rgbt, rgba); ctx.frame_resize(wl["W"], wl["H"])
    ctx.gt_set_rays(occ_r, sdw_r)
    prm = capi.default_gt_params(float(np.sqrt(3.0) * n), 64, 64)
    prm.count_samples = 1
    ctx.gt_render(cam, light, prm)
    counts = (ctx.last_sample_count, ctx.last_aux_count)
    img = ctx.frame_read().copy()
    prm.count_samples = 0
    ctx.set_kernel_timing(True)
    ts = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(STREAM)
    for _ in range(2):
        ctx.gt_render(cam, light, prm)
        ts.append(ctx.last_kernel_ms()[0])
    e1.record(STREAM); torch.cuda.synchronize()
    same = None if ref is None else bool(np.array_equal(img.view(np.uint32), ref.view(np.uint32)))
    if ref is None:
        ref = img
    print(v, f"frame {e0.elapsed_time(e1) / 2:.1f} ms, shade {np.mean(ts):.1f} ms, counts {counts}, identical_to_first={same}, checksum {float(img.sum()):.3f}", flush=True)
    ctx.close()
