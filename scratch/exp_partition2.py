import sys, numpy as np, torch
sys.path.insert(0, '.')
import cpp_volume_rendering_b200 as vrb
from cpp_volume_rendering_b200 import capi, synth
import bench
wl = bench.WORKLOADS['cfg2']
vox = bench.make_volume(wl); n=wl['n']; W,H=wl['W'],wl['H']
rgbt, rgba, lut = bench.host_tf_arrays(wl['tf'], 1)
eye, center, up = synth.camera_state(0, n)
cam = capi.make_camera(eye, center, up, W, H)
ctx = vrb.Context(0)
s = torch.cuda.Stream(); torch.cuda.set_stream(s); ctx.set_stream(s.cuda_stream)
ctx.volume_upload(vox); ctx.tf_upload(rgbt, rgba); ctx.frame_resize(W,H); ctx.sat_build(lut)
light = capi.default_lighting(light_pos=synth.light_position(n), forward=synth.camera_forward(eye, center))
prm = capi.default_ebs_params(float(np.sqrt(3.0)*n))
def t(reps=5):
    for _ in range(2): ctx.ebs_render(cam, light, prm)
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(s)
    for _ in range(reps): ctx.ebs_render(cam, light, prm)
    e1.record(s); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps
out = ["full %.2f" % t()]
img = ctx.frame_read(); fin = np.isfinite(img)
out.append("checksum %.1f nonfinite %d" % (float(img[fin].sum()), int((~fin).sum())))
for nr in (2, 4, 8):
    ctx.set_partition(0, nr, 32, 32); out.append("1/%d %.2f" % (nr, t()))
ctx.set_partition(0,1)
prm.count_samples=1; ctx.ebs_render(cam, light, prm); out.append("samples %d queries %d" % (ctx.last_sample_count, ctx.last_aux_count))
print(" | ".join(out), flush=True)
