"""A/B of the two exact rc1pdosct kernels (k_dos_compact vs k_dos) on the GPU: frames must be bit-identical, counters equal."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import cpp_volume_rendering_b200 as vrb
from cpp_volume_rendering_b200 import capi, synth
import bench

def scene(n, dt, W, H, occ=(20.0, 1, 0.35), sdw=(0.5, 0, 1.0), pyr=(128, 128, 128), vol="gauss_noise", tf="bonsai", scale=None, **kw):
    wl = dict(volume=vol, dtype=dt, n=n)
    vox = bench.make_volume(wl)
    bpv = vox.dtype.itemsize
    rgbt, rgba, lut = bench.host_tf_arrays(tf, bpv)
    eye, center, up = synth.camera_state(0, n)
    cam = capi.make_camera(eye, center, up, W, H)
    ctx = vrb.Context(0)
    ctx.set_stream(STREAM.cuda_stream)
    if scale is not None:
        ctx.volume_upload(vox, scale=scale)
    else:
        ctx.volume_upload(vox)
    ctx.tf_upload(rgbt, rgba)
    ctx.frame_resize(W, H)
    diag = float(np.sqrt(3.0) * n)
    ctx.extcoef_build(1.0, pyr)
    o, _, _ = capi.host_cone_sampler(occ[0], occ[1], 0.5 * diag, occ[2])
    s, _, _ = capi.host_cone_sampler(sdw[0], sdw[1], 0.75 * diag, sdw[2])
    ctx.dos_set_cones(o, s)
    light = capi.default_lighting(light_pos=synth.light_position(n), forward=synth.camera_forward(eye, center))
    prm = capi.default_dos_params(0.5, apply_shadow=True)
    for k, v in kw.items():
        setattr(prm, k, v)
    return ctx, cam, light, prm

def run(ctx, cam, light, prm, kernel, reps):
    os.environ["VRB_DOS_KERNEL"] = kernel
    prm.count_samples = 1
    ctx.dos_render(cam, light, prm)
    counts = (ctx.last_sample_count, int(ctx.lib.vrb_last_aux_count(ctx.h)))
    img = ctx.frame_read().copy()
    prm.count_samples = 0
    ctx.dos_render(cam, light, prm)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(STREAM)
    for _ in range(reps):
        ctx.dos_render(cam, light, prm)
    e1.record(STREAM)
    torch.cuda.synchronize()
    return img, counts, e0.elapsed_time(e1) / reps

cases = [
    ("cfg3 512^3 u16 1080p", dict(n=512, dt="u16", W=1920, H=1080), 5),
    ("128^3 u8 7-ray AO", dict(n=128, dt="u8", W=640, H=360, occ=(30.0, 2, 0.35), pyr=(64, 64, 64)), 3),
    ("96^3 u8 non-pow2 pyramid 48x40x56, 3-ray shadow", dict(n=96, dt="u8", W=320, H=240, sdw=(5.0, 1, 1.0), pyr=(48, 40, 56)), 3),
    ("100^3 u16 non-pow2 volume, spot", dict(n=100, dt="u16", W=320, H=240, pyr=(32, 32, 32), type_of_shadow=1, spot_cos=0.9), 3),
    ("64^3 u8 directional, no AO", dict(n=64, dt="u8", W=320, H=240, pyr=(32, 32, 32), type_of_shadow=2, apply_occlusion=0), 3),
]
if len(sys.argv) > 1 and sys.argv[1] == "quick":
    cases = cases[:1]
torch.cuda.set_device(0)
STREAM = torch.cuda.Stream(device=0)
torch.cuda.set_stream(STREAM)
for name, kw, reps in cases:
    ctx, cam, light, prm = scene(**kw)
    d, cd, td = run(ctx, cam, light, prm, "deferred", reps)
    if len(sys.argv) > 1 and sys.argv[1] == "quick":
        print(f"{name}: deferred {td:.3f} ms, counts {cd}", flush=True)
        ctx.close()
        continue
    a, ca, ta = run(ctx, cam, light, prm, "compact", reps)
    b, cb, tb = run(ctx, cam, light, prm, "ray", reps)
    same = np.array_equal(a.view(np.uint32), b.view(np.uint32)) and np.array_equal(d.view(np.uint32), b.view(np.uint32))
    print(f"{name}: deferred {td:.3f} ms, compact {ta:.3f} ms, ray {tb:.3f} ms, speedup {tb / td:.2f}x, identical={same}, counts {cd} {ca} {cb}, "
          f"max|d|={float(np.nanmax(np.abs(d - b))):.3g}, checksum {float(np.nansum(d)):.4f}", flush=True)
    ctx.close()
