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
ctx.sat_set_order("scan")
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
for tile in (8, 16, 32, 64, 128):
    row = []
    for nr in (8,):
        ts = []
        for r in range(nr):
            ctx.set_partition(r, nr, tile, tile); ts.append(t(3))
        row.append("1/%d: rank times min %.2f max %.2f mean %.2f" % (nr, min(ts), max(ts), sum(ts)/len(ts)))
    print("tile %3d | %s" % (tile, " | ".join(row)), flush=True)
