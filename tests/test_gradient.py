"""Gradient texture + gradient Blinn-Phong branch of every ShadeSample.

CPU: the oracle's gradient generators pinned BIT FOR BIT against the reference's own libs/volvis_utils/utils.cpp
(GenerateSobelFeldmanGradientTexture, GenerateGradientTexture, GenerateRTexture compiled into oracle/_ref), known
answers of the Phong branch.  GPU: vrb_gradient_build against the oracle, the Phong branch of the five marchers against
the oracle at BASELINE's tolerance, the C++ host mirror path."""
import ctypes as C

import numpy as np
import pytest

from cpp_volume_rendering_b200 import capi, synth
from oracle import bind
from conftest import assert_image_parity


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


VOLS = [("noise24-u8", lambda: synth.volume_noise(24)), ("boxes20-u8", lambda: synth.volume_boxes(20)),
        ("gauss18-u16", lambda: synth.volume_gauss(18, np.uint16)),
        ("noise-ragged-u16", lambda: synth.volume_noise(24, np.uint16)[:9, :14, :21].copy()),
        ("single-voxel", lambda: np.full((1, 1, 1), 200, np.uint8))]


# ---------------------------------------------------------------------------------------------------------------- CPU
@pytest.mark.parametrize("name,mk", VOLS, ids=[v[0] for v in VOLS])
@pytest.mark.parametrize("mode", [1, 2])
def test_oracle_gradient_is_bit_identical_to_the_reference_generators(name, mk, mode):
    if bind.ref() is None:
        pytest.skip("oracle/_ref/libref.so was not built (no /root/reference here)")
    vox = mk()
    got = bind.gradient_build(vox, mode)
    want = bind.gradient_build(vox, mode, use_ref=True)
    assert got.shape == vox.shape + (3,)
    assert np.array_equal(got, want), (name, mode, float(np.abs(got - want).max()))
    if mode == 2:                                              # normalised or exactly zero (NaN -> 0)
        ln = np.sqrt((got.astype(np.float64) ** 2).sum(-1))
        assert np.all((np.abs(ln - 1.0) < 2e-3) | (ln == 0.0))
    if name == "single-voxel":
        assert np.array_equal(got, np.zeros((1, 1, 1, 3), np.float32))   # symmetric stencil over zeros outside


def test_oracle_volume_texels_match_the_reference_rtexture():
    r = bind.ref()
    if r is None:
        pytest.skip("oracle/_ref/libref.so was not built")
    for _, mk in VOLS[:4]:
        vox = np.ascontiguousarray(mk())
        d, h, w = vox.shape
        out = np.empty(vox.shape, np.float32)
        assert r.ref_volume_rtexture(_p(vox), w, h, d, vox.dtype.itemsize, _p(out)) == 1
        assert np.array_equal(out.astype(np.float16).astype(np.float32), bind.volume_r16f(vox))   # GL_R16F upload rounding


def test_oracle_shader_sobel_follows_the_cpu_sobel():
    """sobelfeldman_generator.comp runs the same stencil in fp32 on the fp16 texels: close to, not equal to, the CPU one."""
    vox = synth.volume_noise(20)
    cpu, gl = bind.gradient_build(vox, 1), bind.gradient_build(vox, 3)
    assert not np.array_equal(cpu, gl)
    assert np.abs(cpu - gl).max() <= 16.0 * 2.0 ** -11 * 4      # 26 taps of fp16-rounded inputs (|v| <= 1, weights <= 4)


def _scene(n=28, W=72, H=64, cam_id=0, tfname="bonsai", vol="gauss"):
    vox = {"gauss": synth.volume_gauss, "noise": synth.volume_noise, "boxes": synth.volume_boxes}[vol](n)
    tf = bind.TF(*synth.TFS[tfname])
    eye, center, up = synth.camera_state(cam_id, n)
    light = capi.default_lighting(light_pos=synth.light_position(n), forward=synth.camera_forward(eye, center))
    return vox, tf, (eye, center, up), light


def test_oracle_phong_known_answers():
    vox, tf, (eye, center, up), light = _scene()
    W, H = 72, 64
    cam = bind.camera(eye, center, up, W, H)
    plain = bind.rc1pass(vox, tf, cam, W, H)
    ol = bind.copy_struct(light, bind.OrcLighting)
    # apply_phong = 0: the lit entry point is the plain renderer
    assert np.array_equal(bind.rc1pass_lit(vox, tf, cam, ol, W, H), plain)
    try:
        # a zero gradient leaves the colour untouched (ray_marching_1p.comp:55)
        bind.set_gradient(np.zeros(vox.shape + (3,), np.float32))
        ol.apply_phong = 1
        assert np.array_equal(bind.rc1pass_lit(vox, tf, cam, ol, W, H), plain)
        # ka = 1, kd = ks = 0: clr * (1 + 0) + 0
        bind.set_gradient(bind.gradient_build(vox, 1))
        ol.ka, ol.kd, ol.ks = 1.0, 0.0, 0.0
        assert np.array_equal(bind.rc1pass_lit(vox, tf, cam, ol, W, H), plain)
        # defaults: darker ambient + diffuse, alpha unchanged, image differs
        ol.ka, ol.kd, ol.ks = 0.5, 0.5, 0.8
        lit = bind.rc1pass_lit(vox, tf, cam, ol, W, H)
        assert np.array_equal(lit[..., 3], plain[..., 3]) and np.abs(lit[..., :3] - plain[..., :3]).max() > 0.05
    finally:
        bind.set_gradient(None)
    # apply_phong without a bound gradient texture is an error in the oracle too
    o = bind.orc()
    out = np.zeros((H, W, 4), np.float32)
    tex = bind.volume_r16f(vox)
    G = np.array(vox.shape[::-1], np.float32)
    rgbt = tf.texture_rgbt()
    assert o.orc_rc1pass_render_lit(_p(tex), vox.shape[2], vox.shape[1], vox.shape[0], _p(G), _p(rgbt), tf.n, C.byref(cam), C.c_float(0.5),
                                    W, H, _p(out), None, C.byref(ol)) == -2


def test_oracle_lit_renderers_zero_gradient_is_the_unlit_image():
    """ApplyPhongShading == 1 with a zero gradient leaves L = clr in every lit ShadeSample (no occlusion, no shadow): the
    image is then the plain single-pass one (same compositing; the sample position is scaled by 1/G instead of divided)."""
    n, W, H = 24, 56, 48
    vox, tf, (eye, center, up), light = _scene(n, W, H)
    cam = bind.camera(eye, center, up, W, H)
    plain = bind.rc1pass(vox, tf, cam, W, H)
    ol = bind.copy_struct(light, bind.OrcLighting)
    ol.apply_phong = 1
    diag = float(np.sqrt(3.0) * n)
    bind.set_gradient(np.zeros(vox.shape + (3,), np.float32))
    try:
        opc = capi.host_opacity_by_density(synth.TFS["bonsai"], 1)
        olev, odims, oms = bind.vct_supervoxels(vox)
        olut = bind.vct_preintegration(opc, 255, oms)
        prm = capi.default_vct_params(255.0, np.float32(oms), 0.5)
        img = bind.vct(vox, tf, olev, odims, olut, cam, ol, bind.copy_struct(prm, bind.OrcVctParams), W, H)
        assert np.abs(img - plain).max() <= 2e-3
        lut = tf.ext_lut(1)
        eprm = capi.default_ebs_params(diag, 0.5)
        img = bind.ebs(vox, tf, bind.sat_build(vox, lut), cam, ol, bind.copy_struct(eprm, bind.OrcEbsParams), W, H)
        assert np.abs(img - plain).max() <= 2e-3
        # and with a real gradient the branch is live in each of them
        bind.set_gradient(bind.gradient_build(vox, 2))
        img2 = bind.ebs(vox, tf, bind.sat_build(vox, lut), cam, ol, bind.copy_struct(eprm, bind.OrcEbsParams), W, H)
        assert np.abs(img2 - plain).max() > 0.02 and np.array_equal(img2[..., 3], img[..., 3])
    finally:
        bind.set_gradient(None)


# ---------------------------------------------------------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def ctx(built):
    c = capi.Context(0)
    yield c
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name,mk", VOLS, ids=[v[0] for v in VOLS])
def test_gradient_build_matches_oracle(ctx, name, mk):
    vox = mk()
    ctx.volume_upload(vox)
    assert ctx.gradient_mode() == capi.GRADIENT_NONE          # a new volume drops the gradient texture
    with pytest.raises(capi.VrbError, match="no gradient"):
        ctx.gradient_read(vox.shape)
    for mode in (capi.GRADIENT_SOBEL_FELDMAN, capi.GRADIENT_FINITE_DIFFERENCES):
        ctx.gradient_build(mode)
        assert ctx.gradient_mode() == mode
        got, want = ctx.gradient_read(vox.shape), bind.gradient_build(vox, mode)
        assert np.array_equal(got, want), (name, mode, float(np.abs(got - want).max()))     # same fp64 operation order
    ctx.gradient_build(capi.GRADIENT_COMPUTE_SHADER_SOBEL)
    got, want = ctx.gradient_read(vox.shape), bind.gradient_build(vox, 3)
    # the oracle samples texel centres through the GL_LINEAR arithmetic (weights ~1e-7), the kernel fetches the texels
    assert np.abs(got - want).max() <= 2.0 ** -10 * max(1.0, float(np.abs(want).max()))
    assert (got != want).mean() < 0.01
    ctx.gradient_build(capi.GRADIENT_NONE)
    assert ctx.gradient_mode() == capi.GRADIENT_NONE


RC1_CASES = [("gauss40-bonsai-sobel", "gauss", 40, "bonsai", 0, 1), ("noise36-ramp-fd", "noise", 36, "ramp", 4, 2),
             ("boxes36-sparse-shader", "boxes", 36, "sparse", 1, 3)]


@pytest.mark.gpu
@pytest.mark.parametrize("filt", ["exact", "hardware"])
@pytest.mark.parametrize("name,vol,n,tfname,cam_id,mode", RC1_CASES, ids=[c[0] for c in RC1_CASES])
def test_rc1pass_blinn_phong_matches_oracle(ctx, name, vol, n, tfname, cam_id, mode, filt):
    W, H = 112, 96
    vox, tf, (eye, center, up), light = _scene(n, W, H, cam_id, tfname, vol)
    ctx.volume_upload(vox); ctx.tf_upload(tf.floats_rgbt(), tf.floats_rgba()); ctx.frame_resize(W, H)
    cam = capi.make_camera(eye, center, up, W, H)
    light.apply_phong = 1
    with pytest.raises(capi.VrbError, match="gradient texture"):
        ctx.rc1pass_render_lit(cam, light)
    ctx.gradient_build(mode)
    ctx.set_filter(filt)
    try:
        ctx.rc1pass_render_lit(cam, light, 0.5, count_samples=True)
        img = ctx.frame_read().copy()
        nsamp = ctx.last_sample_count
        light.apply_phong = 0
        ctx.rc1pass_render_lit(cam, light, 0.5)                 # == vrb_rc1pass_render
        plain = ctx.frame_read().copy()
        ctx.rc1pass_render(cam, 0.5)
        assert np.array_equal(plain, ctx.frame_read())
    finally:
        ctx.set_filter("exact")
    light.apply_phong = 1
    bind.set_gradient(bind.gradient_build(vox, mode))
    try:
        ref, ns = bind.rc1pass_lit(vox, tf, bind.camera(eye, center, up, W, H), bind.copy_struct(light, bind.OrcLighting), W, H, count=True)
    finally:
        bind.set_gradient(None)
    assert np.abs(ref[..., :3] - plain[..., :3]).max() > 0.03, "the Phong branch must change the image for the test to mean anything"
    bounds = dict(max_abs=8.0 / 255.0) if (filt == "hardware" and vol == "boxes") else {}
    assert_image_parity(img, ref, what=f"rc1pass phong {name} [{filt}]", **bounds)
    assert abs(nsamp - int(ns.sum())) <= (int(ns.sum()) // 200 if filt == "hardware" else max(2, int(ns.sum()) // 100000))


@pytest.mark.gpu
@pytest.mark.parametrize("renderer", ["dos", "ebs", "vct", "gt"])
@pytest.mark.parametrize("mode", [1, 2])
def test_lit_renderers_blinn_phong_matches_oracle(ctx, renderer, mode):
    """ApplyPhongShading branch of rc1pdosct / rc1pextbsd / rc1pvctsg / rc1pcrtgt ShadeSample."""
    n, W, H = 36, 88, 80
    vox, tf, (eye, center, up), light = _scene(n, W, H, 0, "bonsai", "gauss")
    ctx.volume_upload(vox); ctx.tf_upload(tf.floats_rgbt(), tf.floats_rgba()); ctx.frame_resize(W, H)
    ctx.gradient_build(mode)
    cam = capi.make_camera(eye, center, up, W, H)
    ocam = bind.camera(eye, center, up, W, H)
    grad = bind.gradient_build(vox, mode)
    diag = float(np.sqrt(3.0) * n)

    def both(render, oracle):
        light.apply_phong = 0
        render(); plain = ctx.frame_read().copy()
        light.apply_phong = 1
        render(); img = ctx.frame_read().copy()
        bind.set_gradient(grad)
        try:
            ref = oracle(bind.copy_struct(light, bind.OrcLighting))
        finally:
            bind.set_gradient(None)
        assert np.abs(ref[..., :3] - plain[..., :3]).max() > 0.02
        assert_image_parity(img, ref, what=f"{renderer} phong mode {mode}")

    if renderer == "dos":
        from test_dos import _cones
        (oc, sc), (ho, hs) = _cones(diag)
        res = (32, 32, 32)
        ctx.extcoef_build(1.0, res); ctx.dos_set_cones(ho, hs)
        prm = capi.default_dos_params(0.5, spot_angle_deg=20.0)
        prm.apply_shadow = 1
        pyr, dims = bind.extcoef_build(vox, tf, 1.0, res)
        both(lambda: ctx.dos_render(cam, light, prm),
             lambda ol: bind.dos(vox, tf, pyr, dims, ocam, ol, oc, sc, bind.copy_struct(prm, bind.OrcDosParams), W, H))
    elif renderer == "ebs":
        lut = tf.ext_lut(1)
        ctx.sat_build(lut)
        prm = capi.default_ebs_params(diag, 0.5)
        sat_ref = bind.sat_build(vox, lut)
        both(lambda: ctx.ebs_render(cam, light, prm),
             lambda ol: bind.ebs(vox, tf, sat_ref, ocam, ol, bind.copy_struct(prm, bind.OrcEbsParams), W, H))
    elif renderer == "vct":
        opc = capi.host_opacity_by_density(synth.TFS["bonsai"], 1)
        ctx.vct_build(opc)
        _, _, ms = ctx.vct_info()
        prm = capi.default_vct_params(255.0, ms, 0.5)
        olev, odims, oms = bind.vct_supervoxels(vox)
        olut = bind.vct_preintegration(opc, 255, oms)
        both(lambda: ctx.vct_render(cam, light, prm),
             lambda ol: bind.vct(vox, tf, olev, odims, olut, ocam, ol, bind.copy_struct(prm, bind.OrcVctParams), W, H))
    else:
        occ, sdw = capi.host_gt_ray_tables(6, 90.0, 5, 10.0)
        prm = capi.default_gt_params(diag, 6, 5, 0.5)
        ctx.gt_set_rays(occ, sdw)
        both(lambda: ctx.gt_render(cam, light, prm),
             lambda ol: bind.gt(vox, tf, ocam, ol, bind.copy_struct(prm, bind.OrcGtParams), occ, sdw, W, H))


@pytest.mark.gpu
def test_gradient_shading_through_cpp_host_mirror(built):
    """DataManager::SetCurrentGradient / UpdateStructuredGradientTexture + the renderers' "Apply Gradient Shading" flag."""
    h = capi.load_host()
    h.vrbh_gradient_name.restype = C.c_char_p
    n, W, H = 36, 96, 80
    vox, tf, (eye, center, up), light0 = _scene(n, W, H)
    rgb, a = synth.TF_BONSAI
    assert h.vrbh_init(0) == 0, h.vrbh_last_error()
    try:
        assert h.vrbh_set_gradient(0) == 0                        # before the volume exists: only the model is recorded
        assert h.vrbh_gradient_name() == b"Sobel-Feldman"
        assert h.vrbh_set_volume(_p(vox), n, n, n, 1, C.c_double(1.0), C.c_double(1.0), C.c_double(1.0)) == 0, h.vrbh_last_error()
        assert h.vrbh_set_tf_points(_p(np.ascontiguousarray(rgb)), len(rgb), _p(np.ascontiguousarray(a)), len(a), 255, 0) == 0
        assert h.vrbh_bind_data() == 0, h.vrbh_last_error()
        assert h.vrbh_reshape(W, H) == 0
        e = np.array(eye, np.float32); c = np.array(center, np.float32); u = np.array(up, np.float32)
        h.vrbh_set_camera(_p(e), _p(c), _p(u))
        lp = np.array(synth.light_position(n), np.float32)
        h.vrbh_set_light_position(_p(lp))
        h.vrbh_update_light_camera_vectors()
        assert h.vrbh_set_renderer(b"s_1rc") == 0, h.vrbh_last_error()
        img = np.zeros((H, W, 4), np.float32)

        def frame():
            assert h.vrbh_display() == 0, h.vrbh_last_error()
            assert h.vrbh_read_rgba(_p(img), img.size) == 0, h.vrbh_last_error()
            return img.copy()
        plain = frame()
        assert h.vrbh_set_param(b"ApplyGradientShading", C.c_double(1.0)) == 0
        lit = frame()
        light = capi.Lighting()
        h.vrbh_get_lighting(C.byref(light))
        light.apply_phong = 1
        bind.set_gradient(bind.gradient_build(vox, 1))
        try:
            ref = bind.rc1pass_lit(vox, tf, bind.camera(eye, center, up, W, H), bind.copy_struct(light, bind.OrcLighting), W, H,
                                   step=float(np.float32(0.5)))
        finally:
            bind.set_gradient(None)
        assert np.abs(lit - plain).max() > 0.03
        assert_image_parity(lit, ref, what="host mirror s_1rc + gradient shading")
        # gradient "None": the flag stays set but ApplyGradientPhongShading falls back to 0 (rc1prenderer.cpp:112)
        assert h.vrbh_set_gradient(3) == 0
        assert np.array_equal(frame(), plain)
        # the lit renderers take the same flag
        assert h.vrbh_set_gradient(1) == 0 and h.vrbh_gradient_name() == b"Finite Diferences"
        assert h.vrbh_set_renderer(b"s_1rc_vct") == 0, h.vrbh_last_error()
        a0 = frame()
        assert h.vrbh_set_param(b"ApplyGradientShading", C.c_double(1.0)) == 0
        a1 = frame()
        assert np.abs(a1 - a0).max() > 0.02
    finally:
        h.vrbh_shutdown()
