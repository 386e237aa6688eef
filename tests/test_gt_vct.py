"""Cone ground truth (rc1pcrtgt) and voxel cone tracing (rc1pvctsg): CPU known answers for the oracle, GPU parity of
the pre-passes and marchers against the oracle."""
import ctypes as C

import numpy as np
import pytest

from cpp_volume_rendering_b200 import capi, synth
from oracle import bind
from conftest import assert_image_parity, hardware_filter_bounds


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


# ---------------------------------------------------------------------------------------------------------------- CPU
def test_gt_ray_tables_are_unit_vectors_inside_the_cone(built):
    occ, sdw = capi.host_gt_ray_tables(64, 90.0, 16, 1.0)
    assert occ.shape == (64, 3) and sdw.shape == (16, 3)
    assert np.allclose(np.linalg.norm(occ, axis=1), 1.0, atol=1e-5)
    # theta = U * pi * aperture / 180 from the +z axis (crtgtrenderer.cpp:139-144)
    assert np.all(np.arccos(np.clip(occ[:, 2], -1, 1)) <= np.pi * 90.0 / 180.0 + 1e-5)
    assert np.all(np.arccos(np.clip(sdw[:, 2], -1, 1)) <= np.pi * 1.0 / 180.0 + 1e-4)
    # regeneration restarts the engine: same table again (engine is constructed inside Update)
    occ2, _ = capi.host_gt_ray_tables(64, 90.0, 16, 1.0)
    assert np.array_equal(occ, occ2)


def test_gt_oracle_transparent_surroundings_known_answer():
    """A TF that is opaque only above the densities present around the sample: every secondary ray sees tau = 0, so
    occlusion = shadow = 1 and the GT image equals rc1pass up to the per-sample fp16 state rounding."""
    n, W, H = 24, 40, 40
    vox = synth.volume_gauss(n)
    tf = bind.TF(*synth.TF_THIN)
    eye, center, up = synth.camera_state(0, n)
    cam = bind.camera(eye, center, up, W, H)
    zero_ext = (synth.TF_THIN[0], np.array([[0.0, 0], [0.0, 255]], np.float64))
    occ, sdw = capi.host_gt_ray_tables(4, 90.0, 4, 1.0)
    light = bind.copy_struct(capi.default_lighting(light_pos=(50.0, 80.0, 120.0)), bind.OrcLighting)
    prm = bind.copy_struct(capi.default_gt_params(float(np.sqrt(3) * n), 4, 4), bind.OrcGtParams)
    img0 = bind.gt(vox, bind.TF(*zero_ext), cam, light, prm, occ, sdw, W, H)
    assert np.all(img0 == 0)                       # nothing is shaded when the TF is fully transparent
    img = bind.gt(vox, tf, cam, light, prm, occ, sdw, W, H)
    ref = bind.rc1pass(vox, tf, cam, W, H, 0.5)
    assert img[..., 3].max() > 0.05
    assert np.all(img[..., :3] <= ref[..., :3] + 2e-3)      # lighting only darkens
    assert np.abs(img[..., 3] - ref[..., 3]).max() < 4e-3   # alpha is independent of the lighting; fp16 state rounding only


def test_vct_supervoxels_known_answer():
    rng = np.random.default_rng(21)
    vox = rng.integers(0, 256, (8, 12, 16)).astype(np.uint8)
    levels, dims, ms = bind.vct_supervoxels(vox)
    assert [tuple(d) for d in dims] == [(16, 12, 8), (8, 6, 4), (4, 3, 2), (2, 1, 1)]     # halve until a dimension hits 0
    l0 = levels[0]
    assert np.array_equal(l0[..., 0], vox.astype(np.float32)) and np.all(l0[..., 1] == 0)   # u8: mean = v exactly
    m0 = vox.astype(np.float64)
    blk = m0.reshape(4, 2, 6, 2, 8, 2).transpose(0, 2, 4, 1, 3, 5).reshape(4, 6, 8, 8)
    mean1, std1 = blk.mean(-1), blk.std(-1)
    with np.errstate(over="ignore"):
        assert np.array_equal(levels[1][..., 0], mean1.astype(np.float32).astype(np.float16).astype(np.float32))
        assert np.allclose(levels[1][..., 1], std1.astype(np.float32).astype(np.float16).astype(np.float32), rtol=1e-3)
    assert ms >= std1.max() - 1e-9
    # 16-bit data: the mean is still scaled to 0..255
    v16 = (vox.astype(np.uint16) * 257)
    l16, _, _ = bind.vct_supervoxels(v16)
    assert np.allclose(l16[0][..., 0], vox.astype(np.float32), atol=0.13)


def test_vct_preintegration_known_answer():
    tf = bind.TF(*synth.TF_RAMP)
    opc = np.array([tf.get_opc(i, 255.0) for i in range(256)], np.float32)
    lut = bind.vct_preintegration(opc, 255, 20.3)
    assert lut.shape == (21, 255)
    with np.errstate(over="ignore"):
        assert np.array_equal(lut[0], opc[:255].astype(np.float16).astype(np.float32))    # stddev 0 row = GetOpc(mean)
    # a linear opacity ramp is reproduced by a symmetric Gaussian away from the ends
    assert np.allclose(lut[5, 40:200], opc[40:200], atol=2e-3)


# ---------------------------------------------------------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def ctx(built):
    c = capi.Context(0)
    yield c
    c.close()


GT_CASES = [
    ("gauss32-occ8-sdw8", lambda: synth.volume_gauss(32), "bonsai", 0, 64, 64, 0.5, 8, 8, dict()),
    ("noise28-occ-only", lambda: synth.volume_noise(28), "ramp", 4, 56, 48, 0.5, 6, 2, dict(apply_shadow=0)),
    ("boxes32-shadow-directional", lambda: synth.volume_boxes(32), "sparse", 1, 56, 56, 0.7, 2, 6, dict(apply_occlusion=0, shadow_type=2)),
    ("gauss24-u16-odd-step", lambda: synth.volume_gauss(24, np.uint16), "thin", 2, 48, 48, 0.3, 4, 4, dict()),
]


@pytest.mark.gpu
@pytest.mark.parametrize("filt", ["exact", "hardware"])
@pytest.mark.parametrize("name,mk,tfname,cam_id,W,H,step,nocc,nsdw,opts", GT_CASES, ids=[c[0] for c in GT_CASES])
def test_gt_matches_oracle(ctx, name, mk, tfname, cam_id, W, H, step, nocc, nsdw, opts, filt):
    vox = mk()
    n = vox.shape[0]
    tf = bind.TF(*synth.TFS[tfname])
    eye, center, up = synth.camera_state(cam_id, n)
    occ, sdw = capi.host_gt_ray_tables(nocc, 90.0, nsdw, 10.0)
    prm = capi.default_gt_params(float(np.sqrt(3.0) * n), nocc, nsdw, step)
    for k, v in opts.items():
        setattr(prm, k, v)
    prm.count_samples = 1
    fwd = synth.camera_forward(eye, center)
    light = capi.default_lighting(light_pos=synth.light_position(n), forward=tuple(-f for f in fwd), up=(0.0, 1.0, 0.0), right=(1.0, 0.0, 0.0))
    ctx.volume_upload(vox)
    ctx.tf_upload(tf.floats_rgbt(), tf.floats_rgba())
    ctx.frame_resize(W, H)
    ctx.gt_set_rays(occ, sdw)
    ctx.set_filter(filt)
    try:
        ctx.gt_render(capi.make_camera(eye, center, up, W, H), light, prm)
    finally:
        ctx.set_filter("exact")
    n_samples, n_aux = ctx.last_sample_count, ctx.last_aux_count
    img = ctx.frame_read()
    ref, ns, nsec = bind.gt(vox, tf, bind.camera(eye, center, up, W, H), bind.copy_struct(light, bind.OrcLighting),
                            bind.copy_struct(prm, bind.OrcGtParams), occ, sdw, W, H, count=True)
    assert (ns > 0).sum() > 100 and ref[..., :3].max() > 0.01
    hw = filt == "hardware"
    assert_image_parity(img, ref, what=f"{name} [{filt}]", **(hardware_filter_bounds(name) if hw else {}))
    assert abs(n_samples - int(ns.sum())) <= (int(ns.sum()) // 200 if hw else max(2, int(ns.sum()) // 10000))
    assert abs(n_aux - nsec) <= (nsec // 100 if hw else max(16, nsec // 1000))


@pytest.mark.gpu
@pytest.mark.parametrize("shape,dt", [((16, 16, 16), np.uint8), ((12, 20, 24), np.uint8), ((8, 8, 8), np.uint16)])
def test_vct_prepass_matches_oracle(ctx, shape, dt):
    vox = synth.volume_noise(max(shape), dt)[:shape[0], :shape[1], :shape[2]].copy()
    bpv = vox.dtype.itemsize
    tf = bind.TF(*synth.TF_BONSAI)
    mx = 255 if bpv == 1 else 65535
    opc_host = capi.host_opacity_by_density(synth.TF_BONSAI, bpv)
    opc_orc = np.array([tf.get_opc(i, float(mx)) for i in range(0, mx + 1, 1 if bpv == 1 else 257)], np.float32)
    assert np.array_equal(opc_host[::1 if bpv == 1 else 257], opc_orc)
    ctx.volume_upload(vox)
    ctx.vct_build(opc_host)
    levels, lut, ms = ctx.vct_read()
    olev, odims, oms = bind.vct_supervoxels(vox)
    assert len(levels) == len(olev) and np.float32(oms) == np.float32(ms)
    for a, b in zip(levels, olev):
        assert a.shape == b.shape and np.array_equal(a, b)          # fp64 arithmetic in the same order: bit-identical
    if bpv == 1:
        olut = bind.vct_preintegration(opc_host, mx, oms)
        assert lut.shape == olut.shape
        assert np.abs(lut - olut).max() <= 2.0 ** -11 * max(1e-3, float(olut.max()))     # exp() rounding only
    else:
        rows = (1, 3)                                                # the CPU LUT is O(65535^2 h): check two rows
        olut = bind.vct_preintegration(opc_host, mx, oms, rows)
        assert np.abs(lut[rows[0]:rows[1]] - olut).max() <= 2.0 ** -11 * max(1e-3, float(olut.max()))


VCT_CASES = [
    ("gauss48-bonsai", lambda: synth.volume_gauss(48), "bonsai", 0, 112, 112, 0.5, dict()),
    ("noise40-ramp", lambda: synth.volume_noise(40), "ramp", 4, 96, 96, 0.5, dict(cone_step_increase_rate=1.1)),
    ("boxes48-sparse-nocorr", lambda: synth.volume_boxes(48), "sparse", 1, 96, 80, 0.6, dict(apply_opacity_correction=0)),
    ("gauss32-shadow-only", lambda: synth.volume_gauss(32), "bonsai", 5, 80, 80, 0.5, dict(apply_occlusion=0, cone_number_of_samples=20)),
]


@pytest.mark.gpu
@pytest.mark.parametrize("filt", ["exact", "hardware"])
@pytest.mark.parametrize("name,mk,tfname,cam_id,W,H,step,opts", VCT_CASES, ids=[c[0] for c in VCT_CASES])
def test_vct_matches_oracle(ctx, name, mk, tfname, cam_id, W, H, step, opts, filt):
    vox = mk()
    n = vox.shape[0]
    tf = bind.TF(*synth.TFS[tfname])
    eye, center, up = synth.camera_state(cam_id, n)
    opc = capi.host_opacity_by_density(synth.TFS[tfname], 1)
    ctx.volume_upload(vox)
    ctx.tf_upload(tf.floats_rgbt(), tf.floats_rgba())
    ctx.vct_build(opc)
    ctx.frame_resize(W, H)
    _, (lw, lh), ms = ctx.vct_info()
    prm = capi.default_vct_params(255.0, ms, step)
    for k, v in opts.items():
        setattr(prm, k, v)
    prm.count_samples = 1
    light = capi.default_lighting(light_pos=synth.light_position(n))
    ctx.set_filter(filt)
    try:
        ctx.vct_render(capi.make_camera(eye, center, up, W, H), light, prm)
    finally:
        ctx.set_filter("exact")
    n_samples = ctx.last_sample_count
    img = ctx.frame_read()
    olev, odims, oms = bind.vct_supervoxels(vox)
    olut = bind.vct_preintegration(opc, 255, oms)
    ref, ns = bind.vct(vox, tf, olev, odims, olut, bind.camera(eye, center, up, W, H), bind.copy_struct(light, bind.OrcLighting),
                       bind.copy_struct(prm, bind.OrcVctParams), W, H, count=True)
    assert (ns > 0).sum() > 100 and ref[..., :3].max() > 0.01
    hw = filt == "hardware"
    assert_image_parity(img, ref, what=f"{name} [{filt}]", **(hardware_filter_bounds(name) if hw else {}))
    assert abs(n_samples - int(ns.sum())) <= (int(ns.sum()) // 200 if hw else max(2, int(ns.sum()) // 100000))


@pytest.mark.gpu
@pytest.mark.parametrize("abbr,setup", [("s_1rc_gt_c", [(b"ApplyConeOcclusion", 1), (b"OccNumberOfSampledRays", 4), (b"ApplyConeShadow", 1), (b"SdwNumberOfSampledRays", 3)]),
                                        ("s_1rc_vct", [])])
def test_gt_and_vct_through_cpp_host_mirror(ctx, built, abbr, setup):
    h = capi.load_host()
    n, W, H = 32, 64, 64
    vox = synth.volume_gauss(n)
    rgb, a = synth.TF_BONSAI
    assert h.vrbh_init(0) == 0, h.vrbh_last_error()
    try:
        assert h.vrbh_set_volume(_p(vox), n, n, n, 1, 1.0, 1.0, 1.0) == 0, h.vrbh_last_error()
        assert h.vrbh_set_tf_points(_p(np.ascontiguousarray(rgb)), len(rgb), _p(np.ascontiguousarray(a)), len(a), 255, 0) == 0
        assert h.vrbh_bind_data() == 0, h.vrbh_last_error()
        assert h.vrbh_reshape(W, H) == 0
        eye, center, up = synth.camera_state(0, n)
        e = np.array(eye, np.float32); c = np.array(center, np.float32); u = np.array(up, np.float32)
        h.vrbh_set_camera(_p(e), _p(c), _p(u))
        lp = np.array(synth.light_position(n), np.float32)
        h.vrbh_set_light_position(_p(lp))
        h.vrbh_update_light_camera_vectors()
        assert h.vrbh_set_renderer(abbr.encode()) == 0, h.vrbh_last_error()
        for k, v in setup:
            assert h.vrbh_set_param(k, float(v)) == 0, h.vrbh_last_error()
        assert h.vrbh_display() == 0, h.vrbh_last_error()
        img = np.zeros((H, W, 4), np.float32)
        assert h.vrbh_read_rgba(_p(img), img.size) == 0, h.vrbh_last_error()
        light = capi.Lighting()
        h.vrbh_get_lighting(C.byref(light))
        tf = bind.TF(rgb, a)
        ocam = bind.camera(eye, center, up, W, H)
        if abbr == "s_1rc_gt_c":
            occ, sdw = capi.host_gt_ray_tables(4, 90.0, 3, 1.0)
            prm = capi.default_gt_params(float(np.sqrt(3.0) * n), 4, 3, 0.5)
            for i in range(3):
                light.light_forward[i] = -light.light_forward[i]
            ref = bind.gt(vox, tf, ocam, bind.copy_struct(light, bind.OrcLighting), bind.copy_struct(prm, bind.OrcGtParams), occ, sdw, W, H)
        else:
            opc = capi.host_opacity_by_density(synth.TF_BONSAI, 1)
            olev, odims, oms = bind.vct_supervoxels(vox)
            olut = bind.vct_preintegration(opc, 255, oms)
            prm = capi.default_vct_params(255.0, np.float32(oms), 0.5)
            ref = bind.vct(vox, tf, olev, odims, olut, ocam, bind.copy_struct(light, bind.OrcLighting), bind.copy_struct(prm, bind.OrcVctParams), W, H)
        assert ref[..., 3].max() > 0.05
        assert_image_parity(img, ref, what=abbr + " host mirror")
    finally:
        h.vrbh_shutdown()


@pytest.mark.gpu
@pytest.mark.parametrize("filt", ["exact", "hardware"])
def test_vct_light_cache_and_object_space_march_match_oracle(ctx, filt):
    """rc1pvctsg/lightcachecomputation.comp (K13: cone without the leave-the-volume cut, Iao = 1) + obj_ray_marching.comp."""
    n, W, H, step, res = 40, 96, 96, 0.5, (16, 12, 8)
    vox = synth.volume_gauss(n)
    tf = bind.TF(*synth.TF_BONSAI)
    eye, center, up = synth.camera_state(0, n)
    opc = capi.host_opacity_by_density(synth.TF_BONSAI, 1)
    ctx.volume_upload(vox)
    ctx.tf_upload(tf.floats_rgbt(), tf.floats_rgba())
    ctx.vct_build(opc)
    ctx.frame_resize(W, H)
    _, _, ms = ctx.vct_info()
    prm = capi.default_vct_params(255.0, ms, step)
    light = capi.default_lighting(light_pos=synth.light_position(n))
    ctx.set_filter(filt)
    try:
        ctx.vct_light_cache_build(light, prm, res)
    finally:
        ctx.set_filter("exact")
    got = ctx.light_cache_read()
    olev, odims, oms = bind.vct_supervoxels(vox)
    olut = bind.vct_preintegration(opc, 255, oms)
    want = bind.vct_light_cache(vox.shape, olev, odims, olut, bind.copy_struct(light, bind.OrcLighting), bind.copy_struct(prm, bind.OrcVctParams), res)
    assert got.shape == want.shape == (res[2], res[1], res[0], 2)
    assert np.all(want[..., 0] == 1.0) and want[..., 1].std() > 1e-3
    # hardware: 50 fixed-point-filtered taps multiply up per cache voxel; the image below is held to the parity bar
    tol = 2.0 ** -9 if filt == "exact" else 4.0 / 255.0
    assert np.abs(got - want).max() <= tol, float(np.abs(got - want).max())
    ctx.obj_march_render(capi.make_camera(eye, center, up, W, H), light, step, prm.apply_occlusion, prm.apply_shadow, count_samples=True)
    img = ctx.frame_read()
    ref, ns = bind.obj_march(vox, tf, bind.camera(eye, center, up, W, H), light.ka, light.kd, prm.apply_occlusion, prm.apply_shadow,
                             step, want, W, H, count=True)
    assert ref[..., :3].max() > 0.01
    assert_image_parity(img, ref, what=f"VCT light cache march [{filt}]")
