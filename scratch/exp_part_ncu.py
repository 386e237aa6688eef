import sys, numpy as np, torch
sys.path.insert(0, '.')
import cpp_volume_rendering_b200 as vrb
from cpp_volume_rendering_b200 import capi, synth
import bench
nr = int(sys.argv[1]) if len(sys.argv) > 1 else 8
wl = bench.WORKLOADS['cfg2']
vox = bench.make_volume(wl); n=wl['n']; W,H=wl['W'],wl['H']
rgbt, rgba, lut = bench.host_tf_arrays(wl['tf'], 1)
eye, center, up = synth.camera_state(0, n)
cam = capi.make_camera(eye, center, up, W, H)
ctx = vrb.Context(0)
ctx.sat_set_order("scan")
ctx.volume_upload(vox); ctx.tf_upload(rgbt, rgba); ctx.frame_resize(W,H); ctx.sat_build(lut)
light = capi.default_lighting(light_pos=synth.light_position(n), forward=synth.camera_forward(eye, center))
prm = capi.default_ebs_params(float(np.sqrt(3.0)*n))
ctx.set_partition(0, nr, 32, 32)
for _ in range(4): ctx.ebs_render(cam, light, prm)
ctx.synchronize()
