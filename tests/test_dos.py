"""Directional occlusion / cone shadows (rc1pdosct).
CPU: the cone section schedule (oracle restatement and C++ host mirror) against the REFERENCE's own
conegaussiansampler.cpp compiled into oracle/_ref; pyramid known answers.
GPU: extinction pyramid and the DOS marcher against the oracle."""
import ctypes as C

import numpy as np
import pytest

from cpp_volume_rendering_b200 import capi, synth
from oracle import bind
from conftest import assert_image_parity, hardware_filter_bounds

needs_ref = pytest.mark.skipif(bind.ref() is None, reason="oracle/_ref/libref.so not built and /root/reference absent")

# (half angle, max packing, covered distance, ui weight): the renderer's defaults for 512^3 / 64^3 plus stress cases
CONE_CASES = [(20.0, 1, 0.5 * np.sqrt(3) * 512, 0.35), (0.5, 0, 0.75 * np.sqrt(3) * 512, 1.0), (20.0, 1, 0.5 * np.sqrt(3) * 64, 0.35),
              (30.0, 2, 100.0, 1.0), (45.0, 2, 300.0, 0.5), (10.0, 0, 50.0, 1.0), (89.5, 2, 40.0, 1.0), (5.0, 1, 5.0, 1.0)]


@needs_ref
@pytest.mark.parametrize("ang,pack,cov,w", CONE_CASES)
def test_cone_schedule_oracle_vs_reference(ang, pack, cov, w):
    p = bind.cone_params(ang, pack, cov, w)
    for sigma0 in (1.0, 0.5, 2.0):
        s, o = bind.cone_sampler(p, sigma0)
        sr, orf = bind.cone_sampler(p, sigma0, use_ref=True)
        assert o.n_sections == orf.n_sections and list(o.counts) == list(orf.counts)
        assert np.array_equal(s, sr)                     # section table: bit-identical floats
        assert bytes(o) == bytes(orf)                    # ray axes and adjacent weights too


@pytest.mark.parametrize("ang,pack,cov,w", CONE_CASES)
def test_cone_schedule_host_mirror_vs_oracle(built, ang, pack, cov, w):
    p = bind.cone_params(ang, pack, cov, w)
    s, o = bind.cone_sampler(p, 1.0)
    cs, sec, adj3 = capi.host_cone_sampler(ang, pack, cov, w)
    assert cs.n_sections == o.n_sections and list(cs.integration_samples) == list(o.counts)
    assert np.array_equal(sec, s)
    assert np.array_equal(np.ctypeslib.as_array(cs.ray_axes), np.ctypeslib.as_array(o.ray_axes))
    assert cs.ray7_adj_weight == o.ray7_adj_weight and np.float32(adj3) == o.ray3_adj_weight


def test_cone_schedule_known_properties():
    """Defaults at 512^3: 18 AO sections (1 + 17 three-ray) and 159 shadow sections."""
    s, o = bind.cone_sampler(bind.cone_params(20.0, 1, 0.5 * np.sqrt(3) * 512, 0.35), 1.0)
    assert o.n_sections == 18 and list(o.counts) == [1, 17, 0]
    s2, o2 = bind.cone_sampler(bind.cone_params(0.5, 0, 0.75 * np.sqrt(3) * 512, 1.0), 1.0)
    assert o2.n_sections == 159 and list(o2.counts) == [159, 0, 0]
    for sec in (s, s2):
        assert np.all(sec[:, 1] == np.round(sec[:, 1])) and np.all(np.diff(sec[:, 1]) >= 0)   # integer, non-decreasing mips
        assert sec[-1, 0] == 0.0 and np.all(sec[:-1, 0] > 0)                                  # last interval has length 0
        assert np.all(sec[1:, 2] == 0.5 * sec[:-1, 0])                                       # d_integral = previous interval / 2
    ax = np.ctypeslib.as_array(o.ray_axes)
    assert np.allclose(np.linalg.norm(ax, axis=1), 1.0, atol=1e-6)
    assert np.allclose(ax[3], [0, 0, 1])


def test_pyramid_uniform_volume_known_answer():
    """Uniform density, opacity a: far from the border every level holds -log(1-a)."""
    vox = np.full((24, 24, 24), 255, np.uint8)
    tf = bind.TF(*synth.TF_RAMP)                       # opacity 0.8 at 255
    pyr, dims = bind.extcoef_build(vox, tf, 1.0, (16, 16, 16))
    assert [tuple(d) for d in dims] == [(16, 16, 16), (8, 8, 8), (4, 4, 4), (2, 2, 2), (1, 1, 1)]
    l0 = pyr[:16 ** 3].reshape(16, 16, 16)
    a16 = np.float32(np.float16(np.float32(0.8)))
    want = np.float32(np.float16(-np.log(np.float32(1.0) - np.float32(np.float16(a16)))))
    assert np.allclose(l0[4:12, 4:12, 4:12], want, rtol=2e-3)
    assert l0[0, 0, 0] < l0[8, 8, 8]                   # taps outside the volume add 0 but still count in sum(w)


# ---------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ctx(built):
    c = capi.Context(0)
    yield c
    c.close()


def _cones(n_diag, occ=(20.0, 1, 0.35), sdw=(0.5, 0, 1.0)):
    po = bind.cone_params(occ[0], occ[1], 0.5 * n_diag, occ[2])
    ps = bind.cone_params(sdw[0], sdw[1], 0.75 * n_diag, sdw[2])
    so, oo = bind.cone_sampler(po, 1.0)
    ss, os_ = bind.cone_sampler(ps, 1.0)
    ho, _, _ = capi.host_cone_sampler(occ[0], occ[1], 0.5 * n_diag, occ[2])
    hs, _, _ = capi.host_cone_sampler(sdw[0], sdw[1], 0.75 * n_diag, sdw[2])
    return (bind.dos_cone(so, oo, po), bind.dos_cone(ss, os_, ps)), (ho, hs)


@pytest.mark.gpu
@pytest.mark.parametrize("shape,dt,res,tfname", [((40, 40, 40), np.uint8, (32, 32, 32), "bonsai"),
                                                 ((24, 36, 48), np.uint16, (16, 24, 32), "ramp"),
                                                 ((20, 20, 20), np.uint8, None, "sparse")])
def test_extinction_pyramid_matches_oracle(ctx, shape, dt, res, tfname):
    vox = (synth.volume_noise(max(shape), dt)[:shape[0], :shape[1], :shape[2]]).copy()
    tf = bind.TF(*synth.TFS[tfname])
    ctx.volume_upload(vox)
    ctx.tf_upload(tf.floats_rgbt(), tf.floats_rgba())
    ctx.extcoef_build(1.0, res)
    levels = ctx.extcoef_levels()
    r = res if res else (shape[2], shape[1], shape[0])
    pyr, dims = bind.extcoef_build(vox, tf, 1.0, r)
    assert len(levels) == len(dims)
    off = 0
    for l, lev in enumerate(levels):
        w, h, d = (int(v) for v in dims[l])
        want = pyr[off:off + w * h * d].reshape(d, h, w)
        off += w * h * d
        assert lev.shape == want.shape
        # every level is stored fp16 and filtered from the fp16 level above: allow a few fp16 ulps, compounding per level
        tol = (2 + l) * 2.0 ** -10
        assert np.all(np.abs(lev - want) <= tol * np.maximum(np.abs(want), 2.0 ** -10)), (l, float(np.abs(lev - want).max()))
        assert np.mean(lev == want) > 0.9


DOS_CASES = [
    ("gauss48-ao", lambda: synth.volume_gauss(48), "bonsai", 0, 112, 112, 0.5, dict()),
    ("noise40-ao+shadow", lambda: synth.volume_noise(40), "ramp", 4, 96, 96, 0.5, dict(apply_shadow=1)),
    ("boxes40-shadow-only-directional", lambda: synth.volume_boxes(40), "sparse", 1, 96, 80, 0.6,
     dict(apply_shadow=1, apply_occlusion=0, type_of_shadow=2)),
    ("gauss40-spot", lambda: synth.volume_gauss(40), "bonsai", 0, 96, 96, 0.5, dict(apply_shadow=1, type_of_shadow=1)),
    ("gauss36-u16-7rays", lambda: synth.volume_gauss(36, np.uint16), "thin", 2, 80, 80, 0.9, dict(apply_shadow=1, occ=(40.0, 2, 0.5), sdw=(10.0, 1, 1.0))),
]


@pytest.mark.gpu
@pytest.mark.parametrize("filt", ["exact", "hardware"])
@pytest.mark.parametrize("name,mk,tfname,cam_id,W,H,step,opts", DOS_CASES, ids=[c[0] for c in DOS_CASES])
def test_dos_matches_oracle(ctx, name, mk, tfname, cam_id, W, H, step, opts, filt):
    """filt = exact: software fp32 blends, loop counts equal the oracle's; hardware: texture units (the reference's
    own GL_LINEAR_MIPMAP_LINEAR sampling), same image within the tolerance."""
    opts = dict(opts)
    vox = mk()
    n = vox.shape[0]
    tf = bind.TF(*synth.TFS[tfname])
    eye, center, up = synth.camera_state(cam_id, n)
    diag = float(np.sqrt(3.0) * n)
    (oc, sc), (ho, hs) = _cones(diag, opts.pop("occ", (20.0, 1, 0.35)), opts.pop("sdw", (0.5, 0, 1.0)))
    res = (32, 32, 32)
    ctx.volume_upload(vox)
    ctx.tf_upload(tf.floats_rgbt(), tf.floats_rgba())
    ctx.extcoef_build(1.0, res)
    ctx.dos_set_cones(ho, hs)
    ctx.frame_resize(W, H)
    prm = capi.default_dos_params(step, spot_angle_deg=20.0)
    for k, v in opts.items():
        setattr(prm, k, v)
    prm.count_samples = 1
    fwd = synth.camera_forward(eye, center)
    light = capi.default_lighting(light_pos=synth.light_position(n), forward=fwd, up=(0.0, 1.0, 0.0), right=(1.0, 0.0, 0.0))
    ctx.set_filter(filt)
    try:
        ctx.dos_render(capi.make_camera(eye, center, up, W, H), light, prm)
        img = ctx.frame_read()
    finally:
        ctx.set_filter("exact")
    # oracle on the GPU-built pyramid's own oracle counterpart
    pyr, dims = bind.extcoef_build(vox, tf, 1.0, res)
    ref, ns = bind.dos(vox, tf, pyr, dims, bind.camera(eye, center, up, W, H), bind.copy_struct(light, bind.OrcLighting), oc, sc,
                       bind.copy_struct(prm, bind.OrcDosParams), W, H, count=True)
    assert (ns > 0).sum() > 100 and ref[..., :3].max() > 0.01
    assert_image_parity(img, ref, what=f"{name} [{filt}]", **(hardware_filter_bounds(name) if filt == "hardware" else {}))
    slack = max(2, int(ns.sum()) // 100000) if filt == "exact" else int(ns.sum()) // 200
    assert abs(ctx.last_sample_count - int(ns.sum())) <= slack


@pytest.mark.gpu
def test_dos_through_cpp_host_mirror(ctx, built):
    h = capi.load_host()
    n, W, H = 40, 96, 96
    vox = synth.volume_gauss(n)
    rgb, a = synth.TF_BONSAI
    assert h.vrbh_init(0) == 0, h.vrbh_last_error()
    try:
        p = lambda x: x.ctypes.data_as(C.c_void_p)
        assert h.vrbh_set_volume(p(vox), n, n, n, 1, 1.0, 1.0, 1.0) == 0, h.vrbh_last_error()
        assert h.vrbh_set_tf_points(p(np.ascontiguousarray(rgb)), len(rgb), p(np.ascontiguousarray(a)), len(a), 255, 0) == 0
        assert h.vrbh_bind_data() == 0, h.vrbh_last_error()
        assert h.vrbh_reshape(W, H) == 0
        eye, center, up = synth.camera_state(0, n)
        e = np.array(eye, np.float32); c = np.array(center, np.float32); u = np.array(up, np.float32)
        h.vrbh_set_camera(p(e), p(c), p(u))
        h.vrbh_update_light_camera_vectors()
        assert h.vrbh_set_renderer(b"s_1rc_dos") == 0, h.vrbh_last_error()      # Init: pyramid (128^3) + cone schedules
        assert h.vrbh_display() == 0, h.vrbh_last_error()
        img = np.zeros((H, W, 4), np.float32)
        assert h.vrbh_read_rgba(p(img), img.size) == 0, h.vrbh_last_error()
        light = capi.Lighting()
        h.vrbh_get_lighting(C.byref(light))
        tf = bind.TF(rgb, a)
        (oc, sc), _ = _cones(float(np.sqrt(3.0) * n))
        pyr, dims = bind.extcoef_build(vox, tf, 1.0, (128, 128, 128))
        prm = capi.default_dos_params(0.5, spot_angle_deg=4.0)
        ref = bind.dos(vox, tf, pyr, dims, bind.camera(eye, center, up, W, H), bind.copy_struct(light, bind.OrcLighting), oc, sc,
                       bind.copy_struct(prm, bind.OrcDosParams), W, H)
        assert_image_parity(img, ref, what="DOS host mirror (defaults: AO on, shadow off, 128^3 pyramid)")
    finally:
        h.vrbh_shutdown()


# ---- object-space light cache (preillumination.cpp, rc1pdosct/lightcachecomputation.comp, obj_ray_marching.comp) ------
def _camera_up(eye, center, up):
    """v_up of Camera::GetCameraVectors (camera.cpp:336-341): forward = -dir, right = up x forward, up = forward x right."""
    e = np.asarray(eye, np.float32); c = np.asarray(center, np.float32); u = np.asarray(up, np.float32)
    d = c - e
    d = d / np.sqrt(np.sum(d * d, dtype=np.float32))
    f = -d
    r = np.cross(u, f).astype(np.float32); r /= np.sqrt(np.sum(r * r, dtype=np.float32))
    v = np.cross(f, r).astype(np.float32); v /= np.sqrt(np.sum(v * v, dtype=np.float32))
    return tuple(float(x) for x in v)


LC_CASES = [
    ("gauss40-ao", lambda: synth.volume_gauss(40), "bonsai", 0, (16, 16, 16), dict()),
    ("noise40-ao+shadow", lambda: synth.volume_noise(40), "ramp", 4, (12, 10, 8), dict(apply_shadow=1)),
    ("gauss36-shadow-only-directional", lambda: synth.volume_gauss(36), "bonsai", 1, (8, 8, 8), dict(apply_shadow=1, apply_occlusion=0, type_of_shadow=2)),
]


@pytest.mark.gpu
@pytest.mark.parametrize("name,mk,tfname,cam_id,res,opts", LC_CASES, ids=[c[0] for c in LC_CASES])
def test_dos_light_cache_and_object_space_march_match_oracle(ctx, name, mk, tfname, cam_id, res, opts):
    vox = mk()
    n = vox.shape[0]
    W, H, step = 96, 96, 0.5
    tf = bind.TF(*synth.TFS[tfname])
    eye, center, up = synth.camera_state(cam_id, n)
    eye_up = _camera_up(eye, center, up)
    (oc, sc), (ho, hs) = _cones(float(np.sqrt(3.0) * n))
    pres = (32, 32, 32)
    ctx.volume_upload(vox)
    ctx.tf_upload(tf.floats_rgbt(), tf.floats_rgba())
    ctx.extcoef_build(1.0, pres)
    ctx.dos_set_cones(ho, hs)
    ctx.frame_resize(W, H)
    prm = capi.default_dos_params(step, spot_angle_deg=20.0)
    for k, v in opts.items():
        setattr(prm, k, v)
    fwd = synth.camera_forward(eye, center)
    light = capi.default_lighting(light_pos=synth.light_position(n), forward=fwd, up=(0.0, 1.0, 0.0), right=(1.0, 0.0, 0.0))
    ctx.dos_light_cache_build(eye, eye_up, light, prm, res)
    got = ctx.light_cache_read()
    pyr, dims = bind.extcoef_build(vox, tf, 1.0, pres)
    want = bind.dos_light_cache(vox.shape, pyr, dims, eye, eye_up, bind.copy_struct(light, bind.OrcLighting), oc, sc,
                                bind.copy_struct(prm, bind.OrcDosParams), res)
    assert got.shape == want.shape == (res[2], res[1], res[0], 2)
    assert 0.0 <= want.min() and want.max() <= 1.0 and want.std() > 1e-3
    # fp16 texels computed from slightly different pyramid roundings: a few fp16 ulps
    assert np.abs(got - want).max() <= 4 * 2.0 ** -11, float(np.abs(got - want).max())
    # the march over the cache
    cam = capi.make_camera(eye, center, up, W, H)
    ctx.obj_march_render(cam, light, step, prm.apply_occlusion, prm.apply_shadow, count_samples=True)
    img = ctx.frame_read()
    ref, ns = bind.obj_march(vox, tf, bind.camera(eye, center, up, W, H), light.ka, light.kd, prm.apply_occlusion, prm.apply_shadow,
                             step, want, W, H, count=True)
    assert (ns > 0).sum() > 100 and ref[..., :3].max() > 0.01
    assert_image_parity(img, ref, what=name + " (object-space march)")
    assert abs(ctx.last_sample_count - int(ns.sum())) <= max(2, int(ns.sum()) // 100000)


@pytest.mark.gpu
def test_obj_march_without_shading_composites_nothing(ctx):
    """obj_ray_marching.comp:312: a sample is composited only if ApplyOcclusion or ApplyShadow is on."""
    vox = synth.volume_gauss(32)
    tf = bind.TF(*synth.TF_BONSAI)
    eye, center, up = synth.camera_state(0, 32)
    (oc, sc), (ho, hs) = _cones(float(np.sqrt(3.0) * 32))
    ctx.volume_upload(vox); ctx.tf_upload(tf.floats_rgbt(), tf.floats_rgba()); ctx.extcoef_build(1.0, (16, 16, 16)); ctx.dos_set_cones(ho, hs)
    ctx.frame_resize(64, 64)
    light = capi.default_lighting(light_pos=synth.light_position(32), forward=(0, 0, 1), up=(0.0, 1.0, 0.0), right=(1.0, 0.0, 0.0))
    ctx.dos_light_cache_build(eye, _camera_up(eye, center, up), light, capi.default_dos_params(0.5), (8, 8, 8))
    ctx.obj_march_render(capi.make_camera(eye, center, up, 64, 64), light, 0.5, 0, 0)
    assert float(np.abs(ctx.frame_read()).max()) == 0.0


@pytest.mark.gpu
def test_dos_light_cache_through_cpp_host_mirror(ctx, built):
    """RC1PConeTracingDirOcclusionShading with UsePreIllumination: Update rebuilds the cache, Redraw runs the object-space march."""
    h = capi.load_host()
    n, W, H = 40, 96, 96
    vox = synth.volume_gauss(n)
    rgb, a = synth.TF_BONSAI
    assert h.vrbh_init(0) == 0, h.vrbh_last_error()
    try:
        p = lambda x: x.ctypes.data_as(C.c_void_p)
        assert h.vrbh_set_volume(p(vox), n, n, n, 1, 1.0, 1.0, 1.0) == 0, h.vrbh_last_error()
        assert h.vrbh_set_tf_points(p(np.ascontiguousarray(rgb)), len(rgb), p(np.ascontiguousarray(a)), len(a), 255, 0) == 0
        assert h.vrbh_bind_data() == 0, h.vrbh_last_error()
        assert h.vrbh_reshape(W, H) == 0
        eye, center, up = synth.camera_state(1, n)
        e = np.array(eye, np.float32); c = np.array(center, np.float32); u = np.array(up, np.float32)
        h.vrbh_set_camera(p(e), p(c), p(u))
        h.vrbh_update_light_camera_vectors()
        assert h.vrbh_set_renderer(b"s_1rc_dos") == 0, h.vrbh_last_error()
        h.vrbh_set_param.argtypes = [C.c_char_p, C.c_double]
        assert h.vrbh_set_param(b"UsePreIllumination", 1.0) == 0, h.vrbh_last_error()
        assert h.vrbh_display() == 0, h.vrbh_last_error()
        img = np.zeros((H, W, 4), np.float32)
        assert h.vrbh_read_rgba(p(img), img.size) == 0, h.vrbh_last_error()
        f3 = np.zeros(3, np.float32); u3 = np.zeros(3, np.float32); r3 = np.zeros(3, np.float32)
        h.vrbh_get_camera_vectors(p(f3), p(u3), p(r3))
        assert np.allclose(u3, _camera_up(eye, center, up), atol=1e-6)
        light = capi.Lighting()
        h.vrbh_get_lighting(C.byref(light))
        tf = bind.TF(rgb, a)
        (oc, sc), _ = _cones(float(np.sqrt(3.0) * n))
        pyr, dims = bind.extcoef_build(vox, tf, 1.0, (128, 128, 128))
        prm = capi.default_dos_params(0.5, spot_angle_deg=4.0)
        cache = bind.dos_light_cache(vox.shape, pyr, dims, eye, tuple(float(x) for x in u3), bind.copy_struct(light, bind.OrcLighting), oc, sc,
                                     bind.copy_struct(prm, bind.OrcDosParams), (32, 32, 32))
        ref = bind.obj_march(vox, tf, bind.camera(eye, center, up, W, H), light.ka, light.kd, 1, 0, 0.5, cache, W, H)
        assert ref[..., :3].max() > 0.01
        assert_image_parity(img, ref, what="DOS host mirror with the light cache (32^3)")
    finally:
        h.vrbh_shutdown()


def test_light_cache_oracle_known_answers():
    """Transparent medium: every cone integrates 0 extinction, so the cache holds exp(0) = 1 for both channels; with both
    terms switched off the shader's defaults (Idao = Idcs = 1.0, lightcachecomputation.comp:533-534) are stored."""
    n = 16
    vox = synth.volume_gauss(n)
    tf = bind.TF(*synth.TF_ZERO)
    (oc, sc), _ = _cones(float(np.sqrt(3.0) * n))
    pyr, dims = bind.extcoef_build(vox, tf, 1.0, (8, 8, 8))
    assert float(np.abs(pyr).max()) == 0.0
    light = bind.copy_struct(capi.default_lighting(light_pos=(30.0, 10.0, 50.0), forward=(0, 0, 1), up=(0, 1, 0), right=(1, 0, 0)), bind.OrcLighting)
    prm = capi.default_dos_params(0.5); prm.apply_shadow = 1
    cache = bind.dos_light_cache(vox.shape, pyr, dims, (20.0, 20.0, 40.0), (0.0, 1.0, 0.0), light, oc, sc, bind.copy_struct(prm, bind.OrcDosParams), (4, 5, 6))
    assert cache.shape == (6, 5, 4, 2) and np.all(cache == 1.0)
    prm.apply_shadow = 0; prm.apply_occlusion = 0
    cache = bind.dos_light_cache(vox.shape, pyr, dims, (20.0, 20.0, 40.0), (0.0, 1.0, 0.0), light, oc, sc, bind.copy_struct(prm, bind.OrcDosParams), (4, 4, 4))
    assert np.all(cache == 1.0)
    # an absorbing medium darkens the cache, more so at the centre than at the corner facing the eye
    tf2 = bind.TF(*synth.TF_RAMP)
    pyr2, dims2 = bind.extcoef_build(vox, tf2, 1.0, (8, 8, 8))
    prm.apply_occlusion = 1
    c2 = bind.dos_light_cache(vox.shape, pyr2, dims2, (20.0, 20.0, 40.0), (0.0, 1.0, 0.0), light, oc, sc, bind.copy_struct(prm, bind.OrcDosParams), (4, 4, 4))
    assert 0.0 < c2[..., 0].min() < c2[..., 0].max() <= 1.0
    assert c2[1, 1, 1, 0] < c2[3, 3, 3, 0]
