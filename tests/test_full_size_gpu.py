"""Parity at BASELINE.json's FULL sizes (the other parity tests use sizes the oracle finishes in a second).
Config 2 is the bench default: rc1pextbsd, 512^3 u8 V-noise, bonsai TF, 1920x1080.  The CPU oracle needs ~1 min for it
(volume synthesis, fp64 SAT, 4.8 M primary samples x 16 SAT box queries on all cores).

Configs 3, 4 and the single-GPU stand-in of config 5 (the workloads bench.py times: same volumes, transfer functions, cameras
and renderer defaults) are compared on EVERY k-th RAY of the full-resolution view: the frame is W/k x H/k pixels with the
aspect ratio of the W x H frame, so the rays are a regular subsample of the 1920x1080 view through the full-size volume
with the full-length cones / secondary rays; CUDA kernel and oracle render the same rays.  Tolerance: BASELINE.json's
(max abs 2/255, PSNR >= 50 dB on float RGBA); loop counts equal."""
import numpy as np
import pytest

from cpp_volume_rendering_b200 import capi, synth
from oracle import bind
from conftest import psnr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(built):
    c = capi.Context(0)
    yield c
    c.close()


def test_config2_full_size_ebs_matches_oracle(ctx):
    n, W, H, step = 512, 1920, 1080, 0.5
    vox = synth.volume_noise(n)
    tf = bind.TF(*synth.TF_BONSAI)
    eye, center, up = synth.camera_state(0, n)
    lut = tf.ext_lut(1)
    ctx.volume_upload(vox)
    ctx.tf_upload(tf.floats_rgbt(), tf.floats_rgba())
    ctx.sat_build(lut)
    ctx.frame_resize(W, H)
    prm = capi.default_ebs_params(float(np.sqrt(3.0) * n), step)
    prm.count_samples = 1
    light = capi.default_lighting(light_pos=synth.light_position(n), forward=synth.camera_forward(eye, center))
    ctx.ebs_render(capi.make_camera(eye, center, up, W, H), light, prm)
    img = ctx.frame_read()
    n_gpu = ctx.last_sample_count
    # the GPU SAT (reference-order wavefront build) must be the reference algorithm's SAT bit for bit: the marcher's box
    # queries are differences of these ~1e7-1e8-sized floats and amplify a single differing ulp
    sat_ref = bind.sat_build(vox, lut)
    sat_gpu = ctx.sat_read(vox.shape)
    assert sat_gpu.shape == sat_ref.shape
    neq = int((sat_gpu != sat_ref).sum())
    print(f"SAT texels differing from the reference-order build: {neq} of {sat_ref.size}")
    assert neq == 0
    del sat_gpu
    ref, ns = bind.ebs(vox, tf, sat_ref, bind.camera(eye, center, up, W, H), bind.copy_struct(light, bind.OrcLighting),
                       bind.copy_struct(prm, bind.OrcEbsParams), W, H, count=True)
    assert n_gpu == int(ns.sum())
    # fp32 cancellation in the SAT queries makes exp(-Stau) overflow RGBA16F at some pixels: the reference's own result
    # (SURVEY.md section 8a12).  Same pixels must be non-finite; the finite ones are held to the parity bar.
    fin_ref, fin_img = np.isfinite(ref), np.isfinite(img)
    print(f"non-finite values: oracle {int((~fin_ref).sum())}, gpu {int((~fin_img).sum())}, differing {int((fin_ref != fin_img).sum())}")
    both = fin_ref & fin_img
    d = np.abs(np.where(both, img, 0.0).astype(np.float64) - np.where(both, ref, 0.0).astype(np.float64))
    print(f"finite pixels: max abs err {d.max():.6f}, values over 2/255: {int((d > 2.0 / 255.0).sum())}, PSNR {psnr(np.where(both, img, 0.0), np.where(both, ref, 0.0)):.1f} dB")
    wi = np.unravel_index(int(np.argmax(d)), d.shape)
    print(f"worst value at (row, col, channel) = {wi}: gpu {img[wi[0], wi[1]]}, oracle {ref[wi[0], wi[1]]}")
    assert np.array_equal(fin_ref, fin_img), f"{int((fin_ref != fin_img).sum())} pixels differ in finiteness"
    assert 0 < int((~fin_ref).sum()) < ref.size // 100
    a = np.where(fin_ref, img, 0.0).astype(np.float64); b = np.where(fin_ref, ref, 0.0).astype(np.float64)
    # The same cancellation leaves a few finite but absurd values (|v| up to 1e4, outside the displayable [0,1] range) where
    # exp() of the noise blows up: there one ulp of expf() is worth more than 2/255 in absolute terms.  Displayable range:
    # BASELINE's absolute bar on the clamped image; out-of-range values: relative 1e-2.
    ac, bc = np.clip(a, 0.0, 1.0), np.clip(b, 0.0, 1.0)
    err = float(np.abs(ac - bc).max())
    assert err <= 2.0 / 255.0, err
    assert psnr(ac, bc) >= 50.0
    wild = np.abs(b) > 1.0
    assert int(wild.sum()) < b.size // 100
    if wild.any():
        rel = float((np.abs(a - b)[wild] / np.abs(b)[wild]).max())
        assert rel <= 1e-2, rel
    ctx.volume_upload(synth.volume_gauss(16))           # release the big buffers


def _sub_view(eye, center, up, W, H, k):
    """Cameras (C ABI, oracle) of every k-th ray per axis of the W x H view: a W/k x H/k frame with the full frame's aspect."""
    sw, sh = W // k, H // k
    gcam = capi.make_camera(eye, center, up, sw, sh)
    ocam = bind.camera(eye, center, up, sw, sh)
    gcam.aspect = ocam.aspect = np.float32(np.float32(W) / np.float32(H))
    return sw, sh, gcam, ocam


def test_config3_full_size_dos_matches_oracle(ctx):
    """rc1pdosct, 512^3 u16 V-gauss+noise, AO 20 deg / 3 rays + cone shadows 0.5 deg ON, 128^3 pyramid (dosrcrenderer.cpp:47-59,
    extcoefvolumegenerator.cpp:10-15), every 8th ray of the 1920x1080 view."""
    n, W, H, step, k = 512, 1920, 1080, 0.5, 8
    vox = synth.volume_gauss_noise(n, np.uint16)
    tf = bind.TF(*synth.TF_BONSAI)
    eye, center, up = synth.camera_state(0, n)
    sw, sh, gcam, ocam = _sub_view(eye, center, up, W, H, k)
    diag = float(np.sqrt(3.0) * n)
    po, ps = bind.cone_params(20.0, 1, 0.5 * diag, 0.35), bind.cone_params(0.5, 0, 0.75 * diag, 1.0)
    so, oo = bind.cone_sampler(po, 1.0)
    ss, os_ = bind.cone_sampler(ps, 1.0)
    ho, _, _ = capi.host_cone_sampler(20.0, 1, 0.5 * diag, 0.35)
    hs, _, _ = capi.host_cone_sampler(0.5, 0, 0.75 * diag, 1.0)
    res = (128, 128, 128)
    ctx.volume_upload(vox)
    ctx.tf_upload(tf.floats_rgbt(), tf.floats_rgba())
    ctx.extcoef_build(1.0, res)
    ctx.dos_set_cones(ho, hs)
    ctx.frame_resize(sw, sh)
    prm = capi.default_dos_params(step, apply_shadow=True)
    prm.count_samples = 1
    light = capi.default_lighting(light_pos=synth.light_position(n), forward=synth.camera_forward(eye, center))
    ctx.dos_render(gcam, light, prm)
    img = ctx.frame_read()
    n_gpu, taps_gpu = ctx.last_sample_count, int(ctx.lib.vrb_last_aux_count(ctx.h))
    # the GPU pyramid against the oracle's (fp16 levels filtered from fp16 levels: a few ulps, compounding per level)
    levels = ctx.extcoef_levels()
    pyr, dims = bind.extcoef_build(vox, tf, 1.0, res)
    off = 0
    for l, lev in enumerate(levels):
        w, h, d = (int(v) for v in dims[l])
        want = pyr[off:off + w * h * d].reshape(d, h, w)
        off += w * h * d
        tol = (2 + l) * 2.0 ** -10
        assert np.all(np.abs(lev - want) <= tol * np.maximum(np.abs(want), 2.0 ** -10)), (l, float(np.abs(lev - want).max()))
    ref, ns = bind.dos(vox, tf, pyr, dims, ocam, bind.copy_struct(light, bind.OrcLighting), bind.dos_cone(so, oo, po), bind.dos_cone(ss, os_, ps),
                       bind.copy_struct(prm, bind.OrcDosParams), sw, sh, count=True)
    hit = int((ns > 0).sum())
    print(f"config 3, every {k}th ray: {hit} hit rays, {int(ns.sum())} loop iterations (gpu {n_gpu}), {taps_gpu} cone taps on the GPU")
    assert hit > 5000 and ref[..., :3].max() > 0.05
    assert n_gpu == int(ns.sum())
    assert taps_gpu % (1 + 3 * 17 + 159) == 0 and taps_gpu > 0          # 18 AO sections (1 + 17 x 3 rays) + 159 shadow sections per shaded sample
    err = float(np.abs(img - ref).max())
    print(f"config 3: max abs err {err:.6f}, PSNR {psnr(img, ref):.1f} dB")
    assert np.isfinite(img).all()
    assert err <= 2.0 / 255.0, err
    assert psnr(img, ref) >= 50.0
    ctx.volume_upload(synth.volume_gauss(16))


def test_config4_full_size_gt_matches_oracle(ctx):
    """rc1pcrtgt, 256^3 u8 V-boxes, 64 occlusion + 64 shadow rays per sample marched to convergence (crtgtrenderer.cpp:35-50,
    272-325), every 16th ray of the 1920x1080 view (1/256 of the rays)."""
    n, W, H, k, rays = 256, 1920, 1080, 16, 64
    vox = synth.volume_boxes(n)
    tf = bind.TF(*synth.TF_RAMP)
    eye, center, up = synth.camera_state(0, n)
    sw, sh, gcam, ocam = _sub_view(eye, center, up, W, H, k)
    occ_r, sdw_r = capi.host_gt_ray_tables(rays, 90.0, rays, 1.0)
    ctx.volume_upload(vox)
    ctx.tf_upload(tf.floats_rgbt(), tf.floats_rgba())
    ctx.gt_set_rays(occ_r, sdw_r)
    ctx.frame_resize(sw, sh)
    fwd = synth.camera_forward(eye, center)
    light = capi.default_lighting(light_pos=synth.light_position(n), forward=tuple(-f for f in fwd))
    prm = capi.default_gt_params(float(np.sqrt(3.0) * n), rays, rays)
    prm.count_samples = 1
    ctx.gt_render(gcam, light, prm)
    img = ctx.frame_read()
    n_gpu, sec_gpu = ctx.last_sample_count, int(ctx.lib.vrb_last_aux_count(ctx.h))
    ref, ns, nsec = bind.gt(vox, tf, ocam, bind.copy_struct(light, bind.OrcLighting), bind.copy_struct(prm, bind.OrcGtParams), occ_r, sdw_r, sw, sh, count=True)
    hit = int((ns > 0).sum())
    print(f"config 4, every {k}th ray: {hit} hit rays, {int(ns.sum())} loop iterations (gpu {n_gpu}), secondary steps oracle {nsec} gpu {sec_gpu}")
    assert hit > 1000 and ref[..., :3].max() > 0.05
    assert n_gpu == int(ns.sum())
    assert abs(sec_gpu - nsec) <= max(16, nsec // 100000)     # a secondary ray ends on (1 - Vt) > 0.99: one ulp of expf() can move one step
    err = float(np.abs(img - ref).max())
    print(f"config 4: max abs err {err:.6f}, PSNR {psnr(img, ref):.1f} dB")
    assert np.isfinite(img).all()
    assert err <= 2.0 / 255.0, err
    assert psnr(img, ref) >= 50.0
    ctx.volume_upload(synth.volume_gauss(16))


def test_config5_single_gpu_full_size_vct_matches_oracle(ctx):
    """rc1pvctsg, 512^3 u16 V-noise (the single-GPU stand-in of config 5 that bench.py times), defaults of vctrenderer.cpp:33-43,
    every 8th ray of the 1920x1080 view.  The super-voxel pyramid must equal the oracle's bit for bit; the 65535 x ceil(maxStd)
    pre-integration LUT is O(65535^2 h) on the CPU, so three of its rows are checked against the oracle and the oracle
    marcher samples the LUT the GPU built."""
    n, W, H, step, k = 512, 1920, 1080, 0.5, 8
    vox = synth.volume_noise(n, np.uint16)
    tf = bind.TF(*synth.TF_BONSAI)
    eye, center, up = synth.camera_state(0, n)
    sw, sh, gcam, ocam = _sub_view(eye, center, up, W, H, k)
    opc = capi.host_opacity_by_density(synth.TF_BONSAI, 2)
    ctx.volume_upload(vox)
    ctx.tf_upload(tf.floats_rgbt(), tf.floats_rgba())
    ctx.vct_build(opc)
    ctx.frame_resize(sw, sh)
    _, (lw, lh), ms = ctx.vct_info()
    prm = capi.default_vct_params(65535.0, ms, step)
    prm.count_samples = 1
    light = capi.default_lighting(light_pos=synth.light_position(n))
    ctx.vct_render(gcam, light, prm)
    img = ctx.frame_read()
    n_gpu = ctx.last_sample_count
    levels, lut, ms2 = ctx.vct_read()
    olev, odims, oms = bind.vct_supervoxels(vox)
    assert len(levels) == len(olev) and np.float32(oms) == np.float32(ms)
    for a, b in zip(levels, olev):
        assert a.shape == b.shape and np.array_equal(a, b)
    rows = (0, min(3, lut.shape[0]))
    olut = bind.vct_preintegration(opc, 65535, oms, rows)
    assert np.abs(lut[rows[0]:rows[1]] - olut).max() <= 2.0 ** -11 * max(1e-3, float(olut.max()))
    del levels
    ref, ns = bind.vct(vox, tf, olev, odims, lut, ocam, bind.copy_struct(light, bind.OrcLighting), bind.copy_struct(prm, bind.OrcVctParams), sw, sh, count=True)
    hit = int((ns > 0).sum())
    print(f"config 5 (1 GPU), every {k}th ray: {hit} hit rays, {int(ns.sum())} loop iterations (gpu {n_gpu})")
    assert hit > 5000 and ref[..., :3].max() > 0.05
    assert n_gpu == int(ns.sum())
    err = float(np.abs(img - ref).max())
    print(f"config 5 (1 GPU): max abs err {err:.6f}, PSNR {psnr(img, ref):.1f} dB")
    assert np.isfinite(img).all()
    assert err <= 2.0 / 255.0, err
    assert psnr(img, ref) >= 50.0
    ctx.volume_upload(synth.volume_gauss(16))
