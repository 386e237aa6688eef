"""rc1pass variants at cfg1: plain / skip_empty / hardware filter (+skip): ms per frame and error vs plain."""
import numpy as np, torch, time
from cpp_volume_rendering_b200 import capi, synth
from oracle import bind
ctx = capi.Context(0)
for n, W, H, tfname, vol in [(256, 768, 768, "bonsai", "gauss"), (512, 1920, 1080, "bonsai", "gauss"), (512, 1920, 1080, "ramp", "noise")]:
    vox = synth.volume_gauss(n) if vol == "gauss" else synth.volume_noise(n)
    tf = bind.TF(*synth.TFS[tfname])
    eye, center, up = synth.camera_state(0, n)
    ctx.volume_upload(vox, (1.0, 1.0, 1.0)); ctx.tf_upload(tf.floats_rgbt(), tf.floats_rgba()); ctx.frame_resize(W, H)
    cam = capi.make_camera(eye, center, up, W, H)
    base = None
    for mode, skip in [("exact", False), ("exact", True), ("hardware", False), ("hardware", True)]:
        ctx.set_filter(mode)
        ctx.rc1pass_render(cam, 0.5, count_samples=True, skip_empty=skip); ns = ctx.last_sample_count
        img = ctx.frame_read()
        if base is None: base = img
        for _ in range(3): ctx.rc1pass_render(cam, 0.5, skip_empty=skip)
        ctx.synchronize(); t0 = time.perf_counter()
        K = 20
        for _ in range(K): ctx.rc1pass_render(cam, 0.5, skip_empty=skip)
        ctx.synchronize(); ms = (time.perf_counter() - t0) / K * 1e3
        print(f"{vol}{n} {W}x{H} {tfname} {mode:8s} skip={int(skip)} {ms:7.3f} ms  samples {ns}  maxerr_vs_plain {np.abs(img-base).max():.6f}", flush=True)
    ctx.set_filter("exact")
