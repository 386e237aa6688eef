"""GPU parity tests of the extinction-based-shading path: the three-pass SAT build and the EBS marcher against the
oracle (whose SAT recurrence is pinned bit-for-bit against the reference's own header in tests/test_oracle_ref.py)."""
import ctypes as C

import numpy as np
import pytest

from cpp_volume_rendering_b200 import capi, synth
from cpp_volume_rendering_b200 import dist as vdist
from oracle import bind
from conftest import assert_image_parity

pytestmark = pytest.mark.gpu

# fp64 accumulation in a different association than the reference's 7-term recurrence: the float32 results may
# differ by a rounding step of the largest prefix sums.  Stated tolerance: |d| <= 1e-6 * S_max.
SAT_REL_TOL = 1e-6


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def ctx(built):
    c = capi.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("shape,dt", [((4, 4, 4), np.uint8), ((33, 20, 47), np.uint8), ((64, 64, 64), np.uint8),
                                      ((17, 40, 35), np.uint16), ((1, 1, 1), np.uint8), ((1, 70, 3), np.uint16)])
@pytest.mark.parametrize("order", ["reference", "scan"])
def test_sat_float_mode_matches_oracle(ctx, shape, dt, order):
    """order = reference (default): BuildSAT's own fp64 recurrence as anti-diagonal wavefronts, BIT-identical floats;
    order = scan: three separable fp64 scans, another association of the same sums, within 1e-6 S_max."""
    rng = np.random.default_rng(11)
    vox = rng.integers(0, np.iinfo(dt).max + 1, shape).astype(dt)
    tf = bind.TF(*synth.TF_BONSAI)
    lut = tf.ext_lut(vox.dtype.itemsize)
    ctx.volume_upload(vox)
    assert ctx.sat_get_order() == "reference"
    ctx.sat_set_order(order)
    try:
        ctx.sat_build(lut)
    finally:
        ctx.sat_set_order("reference")
    got = ctx.sat_read(shape)
    want, w64 = bind.sat_build(vox, lut, want_f64=True)
    assert got.shape == want.shape
    smax = float(want.max())
    if order == "reference":
        assert np.array_equal(got, want), f"{int((got != want).sum())} texels differ"
    assert np.abs(got.astype(np.float64) - want.astype(np.float64)).max() <= SAT_REL_TOL * max(smax, 1e-30)
    # the zero border: first planes are exactly 0, last planes repeat the previous ones
    assert np.all(got[0] == 0) and np.all(got[:, 0] == 0) and np.all(got[:, :, 0] == 0)
    assert np.array_equal(got[-1], got[-2]) and np.array_equal(got[:, -1], got[:, -2]) and np.array_equal(got[:, :, -1], got[:, :, -2])
    # exact total against numpy's fp64 sum of the same float inputs
    tot = float(lut[vox].astype(np.float64).sum())
    assert abs(float(got[-1, -1, -1]) - tot) <= 1e-6 * max(tot, 1e-30)


def test_sat_all_ones_and_dyadic_inputs_are_exact(ctx):
    """Inputs that are exactly representable sums (all ones; dyadic rationals) leave no room for association effects:
    the float SAT must be bit-identical to the reference recurrence."""
    vox = np.ones((20, 31, 45), np.uint8)
    lut = np.zeros(256, np.float32); lut[1] = 1.0
    ctx.volume_upload(vox)
    ctx.sat_build(lut)
    got = ctx.sat_read(vox.shape)
    assert np.array_equal(got, bind.sat_build(vox, lut))
    z, y, x = np.meshgrid(np.arange(22), np.arange(33), np.arange(47), indexing="ij")
    assert np.array_equal(got, (np.minimum(x, 45) * np.minimum(y, 31) * np.minimum(z, 20)).astype(np.float32))
    rng = np.random.default_rng(12)
    vox = rng.integers(0, 256, (24, 24, 24)).astype(np.uint8)
    lut = (np.arange(256) / 64.0).astype(np.float32)          # multiples of 2^-6: every partial sum is exact in fp64
    ctx.volume_upload(vox)
    ctx.sat_build(lut)
    assert np.array_equal(ctx.sat_read(vox.shape), bind.sat_build(vox, lut))


@pytest.mark.parametrize("shape,dt", [((16, 16, 16), np.uint8), ((37, 21, 50), np.uint8), ((9, 33, 65), np.uint16)])
def test_sat_integer_mode_is_bit_exact(ctx, shape, dt):
    rng = np.random.default_rng(13)
    vox = rng.integers(0, np.iinfo(dt).max + 1, shape).astype(dt)
    lut = rng.integers(0, 1 << 24, 256 if dt == np.uint8 else 65536).astype(np.uint32)
    ctx.volume_upload(vox)
    got = ctx.sat_build_u64(lut, shape)
    assert np.array_equal(got, bind.sat_build_u64(vox, lut))
    assert np.array_equal(got, lut[vox].astype(np.uint64).cumsum(0).cumsum(1).cumsum(2))


def test_sat_512_cubed_properties(ctx):
    """Config-2 size: too slow for the CPU recurrence inside a test, so check size-independent properties:
    integer mode against numpy cumsum on a slab, float mode via box sums recovered by inclusion-exclusion."""
    n = 512
    rng = np.random.default_rng(14)
    vox = rng.integers(0, 256, (n, n, n), dtype=np.uint8)
    ctx.volume_upload(vox)
    lut = np.arange(256, dtype=np.uint32)                      # SAT of the voxel values themselves
    got = ctx.sat_build_u64(lut, vox.shape)
    assert int(got[-1, -1, -1]) == int(vox.sum(dtype=np.uint64))
    for z in (0, 1, 255, 511):                                  # checksum of checksums: plane totals
        assert int(got[z, -1, -1]) == int(vox[: z + 1].sum(dtype=np.uint64))
    sub = vox[:40, :50, :60].astype(np.uint64).cumsum(0).cumsum(1).cumsum(2)
    assert np.array_equal(got[:40, :50, :60], sub)
    del got
    flut = (np.arange(256) / 128.0).astype(np.float32)         # dyadic: exact in fp64, so float SAT = float(exact)
    ctx.sat_build(flut)
    sat = ctx.sat_read(vox.shape)
    assert float(sat[-1, -1, -1]) == float(np.float32(float(vox.sum(dtype=np.uint64)) / 128.0))
    ex = vox[:33, :47, :29].astype(np.float64).cumsum(0).cumsum(1).cumsum(2) / 128.0
    assert np.array_equal(sat[1:34, 1:48, 1:30], ex.astype(np.float32))
    ctx.volume_upload(synth.volume_gauss(16))                   # free the big buffers


def _ebs_case(ctx, vox, tfname, cam_id, W, H, step, params_mod=None, scale=(1.0, 1.0, 1.0), light_pos=None):
    d, h, w = vox.shape
    n = max(w, h, d)
    tf = bind.TF(*synth.TFS[tfname])
    eye, center, up = synth.camera_state(cam_id, n)
    lut = tf.ext_lut(vox.dtype.itemsize)
    ctx.volume_upload(vox, scale)
    ctx.tf_upload(tf.floats_rgbt(), tf.floats_rgba())
    ctx.sat_build(lut)
    ctx.frame_resize(W, H)
    diag = float(np.sqrt((w * scale[0]) ** 2 + (h * scale[1]) ** 2 + (d * scale[2]) ** 2))
    prm = capi.default_ebs_params(diag, step)
    if params_mod:
        params_mod(prm)
    light = capi.default_lighting(light_pos=light_pos or synth.light_position(n), forward=synth.camera_forward(eye, center))
    prm.count_samples = 1
    ctx.ebs_render(capi.make_camera(eye, center, up, W, H), light, prm)
    img = ctx.frame_read()
    sat_ref = bind.sat_build(vox, lut)
    ref, ns = bind.ebs(vox, tf, sat_ref, bind.camera(eye, center, up, W, H), bind.copy_struct(light, bind.OrcLighting),
                       bind.copy_struct(prm, bind.OrcEbsParams), W, H, scale, count=True)
    assert (ns > 0).sum() > 100
    assert abs(ctx.last_sample_count - int(ns.sum())) <= max(2, int(ns.sum()) // 100000)
    return img, ref


EBS_CASES = [
    ("gauss48-bonsai-default", lambda: synth.volume_gauss(48), "bonsai", 0, 128, 128, 0.5, None),
    ("noise40-ramp-cam4", lambda: synth.volume_noise(40), "ramp", 4, 112, 96, 0.5, None),
    ("boxes48-sparse-cam1", lambda: synth.volume_boxes(48), "sparse", 1, 96, 96, 0.6, None),
    ("gauss40-ao-only", lambda: synth.volume_gauss(40), "bonsai", 0, 96, 96, 0.5, lambda p: setattr(p, "apply_shadow", 0)),
    ("gauss40-shadow-only", lambda: synth.volume_gauss(40), "bonsai", 5, 96, 96, 0.5, lambda p: setattr(p, "apply_occlusion", 0)),
    ("gauss40-directional", lambda: synth.volume_gauss(40), "ramp", 2, 96, 96, 0.5, lambda p: setattr(p, "type_of_shadow", 1)),
    ("noise32-u16-5shells", lambda: synth.volume_noise(32, np.uint16), "thin", 3, 96, 80, 0.8, lambda p: setattr(p, "amb_occ_shells", 5)),
]


@pytest.mark.parametrize("name,mk,tfname,cam_id,W,H,step,mod", EBS_CASES, ids=[c[0] for c in EBS_CASES])
def test_ebs_matches_oracle(ctx, name, mk, tfname, cam_id, W, H, step, mod):
    img, ref = _ebs_case(ctx, mk(), tfname, cam_id, W, H, step, mod)
    assert_image_parity(img, ref, what=name)


@pytest.mark.parametrize("lp", [(300.0, 10.0, 5.0), (-5.0, -400.0, 20.0), (10.0, 20.0, -350.0)])
def test_ebs_all_three_dominant_light_axes(ctx, lp):
    """ConeXAxis / ConeYAxis / ConeZAxis are chosen per sample from the dominant component of the light direction."""
    img, ref = _ebs_case(ctx, synth.volume_gauss(40), "bonsai", 0, 96, 96, 0.5, None, light_pos=lp)
    assert_image_parity(img, ref, what=f"light {lp}")


def test_ebs_anisotropic_voxels(ctx):
    vox = synth.volume_noise(40)[:24, :32, :]
    img, ref = _ebs_case(ctx, vox, "ramp", 0, 112, 80, 0.5, None, scale=(1.0, 1.25, 2.0))
    assert_image_parity(img, ref, what="anisotropic")


def test_ebs_through_cpp_host_mirror(ctx, built):
    h = capi.load_host()
    n, W, H = 40, 96, 96
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
        h.vrbh_update_light_camera_vectors()                    # renderingmanager.cpp:168
        assert h.vrbh_set_renderer(b"s_1rc_eb") == 0, h.vrbh_last_error()   # Init: TF textures + SAT build
        assert h.vrbh_display() == 0, h.vrbh_last_error()
        img = np.zeros((H, W, 4), np.float32)
        assert h.vrbh_read_rgba(_p(img), img.size) == 0, h.vrbh_last_error()
        light = capi.Lighting()
        h.vrbh_get_lighting(C.byref(light))
        tf = bind.TF(rgb, a)
        prm = capi.default_ebs_params(float(np.sqrt(3.0) * n), 0.5)
        sat_ref = bind.sat_build(vox, tf.ext_lut(1))
        ref = bind.ebs(vox, tf, sat_ref, bind.camera(eye, center, up, W, H), bind.copy_struct(light, bind.OrcLighting),
                       bind.copy_struct(prm, bind.OrcEbsParams), W, H)
        assert_image_parity(img, ref, what="EBS host mirror")
        assert h.vrbh_set_param(b"AmbOccShells", 4) == 0 and h.vrbh_set_param(b"ApplyShadow", 0) == 0
        assert h.vrbh_display() == 0
        assert h.vrbh_read_rgba(_p(img), img.size) == 0
        prm.amb_occ_shells = 4; prm.apply_shadow = 0
        ref = bind.ebs(vox, tf, sat_ref, bind.camera(eye, center, up, W, H), bind.copy_struct(light, bind.OrcLighting),
                       bind.copy_struct(prm, bind.OrcEbsParams), W, H)
        assert_image_parity(img, ref, what="EBS host mirror, 4 shells, no shadow")
    finally:
        h.vrbh_shutdown()


@pytest.mark.parametrize("mod", [None, lambda p: setattr(p, "apply_shadow", 0), lambda p: setattr(p, "type_of_shadow", 1)],
                         ids=["ao+shadow", "ao-only", "directional"])
def test_ebs_light_cache_and_object_space_march_match_oracle(ctx, mod):
    """rc1pextbsd/lightcachecomputation.comp (K9) + obj_ray_marching.comp (K7): the SAT queries are the marcher's own, so the
    cache must agree to fp16 rounding and the image to the parity tolerance."""
    n, W, H, step, res = 40, 96, 96, 0.5, (12, 10, 14)
    vox = synth.volume_gauss(n)
    tf = bind.TF(*synth.TF_BONSAI)
    eye, center, up = synth.camera_state(0, n)
    lut = tf.ext_lut(1)
    ctx.volume_upload(vox)
    ctx.tf_upload(tf.floats_rgbt(), tf.floats_rgba())
    ctx.sat_build(lut)
    ctx.frame_resize(W, H)
    prm = capi.default_ebs_params(float(np.sqrt(3.0) * n), step)
    if mod:
        mod(prm)
    light = capi.default_lighting(light_pos=synth.light_position(n), forward=synth.camera_forward(eye, center))
    ctx.ebs_light_cache_build(light, prm, res)
    got = ctx.light_cache_read()
    sat_ref = bind.sat_build(vox, lut)
    want = bind.ebs_light_cache(vox.shape, sat_ref, bind.copy_struct(light, bind.OrcLighting), bind.copy_struct(prm, bind.OrcEbsParams), res)
    assert got.shape == want.shape == (res[2], res[1], res[0], 2)
    assert want[..., 0].std() > 1e-3
    assert np.abs(got - want).max() <= 2.0 ** -10 * max(1.0, float(np.abs(want).max())), float(np.abs(got - want).max())
    ctx.obj_march_render(capi.make_camera(eye, center, up, W, H), light, step, prm.apply_occlusion, prm.apply_shadow, count_samples=True)
    img = ctx.frame_read()
    ref, ns = bind.obj_march(vox, tf, bind.camera(eye, center, up, W, H), light.ka, light.kd, prm.apply_occlusion, prm.apply_shadow,
                             step, want, W, H, count=True)
    assert ref[..., :3].max() > 0.01
    assert_image_parity(img, ref, what="EBS light cache march")
    assert abs(ctx.last_sample_count - int(ns.sum())) <= max(2, int(ns.sum()) // 100000)


def test_sharded_sat_slabs_equal_the_single_gpu_scan(ctx):
    """SURVEY.md section 8e, SAT row: z-slabs scanned independently + the prefix of the slabs' last fp64 planes.  Two and three
    "ranks" are played by contexts on this one GPU (the all-gather is a host-side sum here; the NCCL version is
    dist.sat_build_sharded, run by tools/sat_sharded_run.py on 2+ GPUs).  Integer-valued extinction: every partial sum is
    exact in fp64, so the table must equal the single-context scan bit for bit; real extinction: the fp64 association differs,
    the float texels may differ by one ulp."""
    import torch
    n = 40
    vox = synth.volume_gauss_noise(n, np.uint8)
    for lut, exact in ((np.arange(256, dtype=np.float32) % 7.0, True), (bind.TF(*synth.TF_BONSAI).ext_lut(1).astype(np.float32), False)):
        lut = np.nan_to_num(lut, posinf=50.0).astype(np.float32)
        ctx.volume_upload(vox)
        ctx.sat_set_order("scan")
        ctx.sat_build(lut)
        want = ctx.sat_read(vox.shape)
        ctx.sat_set_order("reference")
        for world in (2, 3):
            bounds = vdist.slab_bounds(n + 2, world)
            ranks = [capi.Context(0) for _ in range(world)]
            planes = []
            for r, c in enumerate(ranks):
                c.volume_upload(vox)
                c.sat_build_slab(lut, *bounds[r])
                p, cnt = c.sat_slab_plane()
                planes.append(vdist._device_tensor(p, cnt, "<f8", torch.device("cuda", 0)).clone())
            got = np.empty_like(want)
            for r, c in enumerate(ranks):
                prefix = None
                if r > 0:
                    prefix = planes[0].clone()
                    for k in range(1, r):
                        prefix += planes[k]
                torch.cuda.synchronize()
                c.sat_finish_slab(prefix.data_ptr() if prefix is not None else None)
                sat = c.sat_read(vox.shape)
                lo, hi = bounds[r]
                got[lo:hi] = sat[lo:hi]
                c.close()
            if exact:
                assert np.array_equal(got, want), (world, float(np.abs(got - want).max()))
            else:
                ulp = np.spacing(np.abs(want).astype(np.float32))
                assert np.all(np.abs(got - want) <= ulp), (world, float((np.abs(got - want) / np.maximum(ulp, 1e-30)).max()))
