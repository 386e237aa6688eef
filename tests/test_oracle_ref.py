"""CPU tests: the oracle restatement against the REFERENCE's own CPU code (oracle/_ref/libref.so, compiled in place
from /root/reference) and against analytic known answers.  The reference ships no tests or golden vectors
(SURVEY.md F1), so these are the pins that exist."""
import ctypes as C

import numpy as np
import pytest

from cpp_volume_rendering_b200 import synth
from oracle import bind


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


needs_ref = pytest.mark.skipif(bind.ref() is None, reason="oracle/_ref/libref.so not built and /root/reference absent")

TF_CASES = [("bonsai", 255, 0), ("ramp", 255, 0), ("sparse", 255, 0), ("thin", 255, 0), ("ramp", 255, 1)]


def test_fp16_rounding_matches_numpy():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.standard_normal(20000).astype(np.float32) * s for s in (1e-8, 1e-5, 1e-3, 1.0, 100.0, 7e4)])
    x = np.concatenate([x, np.array([0.0, -0.0, 65504.0, 65519.9, 65520.0, 1e9, 2.0 ** -24, 2.0 ** -25, 3 * 2.0 ** -26,
                                     6.1e-5, np.inf, -np.inf], np.float32)])
    out = np.empty_like(x)
    bind.orc().orc_round_f16_array(_p(x), _p(out), x.size)
    with np.errstate(over="ignore"):
        want = x.astype(np.float16).astype(np.float32)
    assert np.array_equal(out.view(np.uint32), want.view(np.uint32))
    # every half bit pattern round-trips
    bits = np.arange(65536, dtype=np.uint16)
    f = bits.view(np.float16).astype(np.float32)
    finite = np.isfinite(f)
    back = np.array([bind.orc().orc_f32_to_f16_bits(float(v)) for v in f[finite][::7]], np.uint16)
    assert np.array_equal(back, bits[finite][::7])


def test_volume_texel_values():
    """half(float(double(v)/255)) and /65535 for every possible voxel value (structuredgridvolume.cpp:130-139)."""
    v8 = np.arange(256, dtype=np.uint8)
    got = bind.volume_r16f(v8.reshape(1, 1, -1)).ravel()
    want = (v8.astype(np.float64) / 255.0).astype(np.float32).astype(np.float16).astype(np.float32)
    assert np.array_equal(got, want)
    v16 = np.arange(65536, dtype=np.uint16)
    got = bind.volume_r16f(v16.reshape(1, 1, -1)).ravel()
    want = (v16.astype(np.float64) / 65535.0).astype(np.float32).astype(np.float16).astype(np.float32)
    assert np.array_equal(got, want)


@needs_ref
def test_volume_normalized_sample_vs_reference():
    rng = np.random.default_rng(1)
    for dt, bpv in ((np.uint8, 1), (np.uint16, 2)):
        vox = rng.integers(0, np.iinfo(dt).max + 1, (5, 6, 7)).astype(dt)
        for (x, y, z) in [(0, 0, 0), (6, 5, 4), (3, 2, 1), (-1, 0, 0), (7, 0, 0), (0, 6, 0), (0, 0, 5)]:
            r = bind.ref().ref_volume_normalized_sample(_p(vox), 7, 6, 5, bpv, x, y, z)
            inside = 0 <= x < 7 and 0 <= y < 6 and 0 <= z < 5
            want = float(vox[z, y, x]) / float(np.iinfo(dt).max) if inside else 0.0
            assert r == want


@needs_ref
@pytest.mark.parametrize("name,maxd,ext", TF_CASES)
def test_transfer_function_vs_reference(name, maxd, ext):
    rgb, a = synth.TFS[name]
    tf = bind.TF(rgb, a, maxd, ext)
    r = bind.ref()
    h = r.ref_tf_create(_p(np.ascontiguousarray(rgb)), len(rgb), _p(np.ascontiguousarray(a)), len(a), maxd, ext)
    try:
        # table at integer iso values (Get with max_data_value < 0)
        for i in range(maxd + 1):
            out = np.zeros(4, np.float32)
            r.ref_tf_get(h, float(i), -1.0, _p(out))
            assert np.array_equal(out, tf.table[i].astype(np.float32)), i
        # GetExtN / GetOpcN / GetOpc on the values the renderers use, plus off-grid ones
        xs = [v / 255.0 for v in range(256)] + [v / 65535.0 for v in range(0, 65536, 257)] + [0.1234, 0.5, 0.99999, 1.0]
        for x in xs:
            assert tf.get_extn(x) == r.ref_tf_get_extn(h, x) or (np.isnan(tf.get_extn(x)) and np.isnan(r.ref_tf_get_extn(h, x)))
            assert tf.get_opcn(x) == r.ref_tf_get_opcn(h, x)
        for v in (0.0, 17.0, 128.5, 254.999, 255.0):
            assert tf.get_opc(v, 255.0) == r.ref_tf_get_opc(h, v, 255.0)
        # the GL_FLOAT client arrays of GenerateTexture_1D_RGBt / _RGBA
        buf = np.zeros((maxd + 1, 4), np.float32)
        assert r.ref_tf_texture_rgbt(h, _p(buf), buf.size) == buf.size
        assert np.array_equal(buf, tf.floats_rgbt())
        assert r.ref_tf_texture_rgba(h, _p(buf), buf.size) == buf.size
        assert np.array_equal(buf, tf.floats_rgba())
    finally:
        r.ref_tf_destroy(h)


def test_tf_texture_is_fp16_of_floats():
    tf = bind.TF(*synth.TF_BONSAI)
    with np.errstate(over="ignore"):
        assert np.array_equal(tf.texture_rgbt(), tf.floats_rgbt().astype(np.float16).astype(np.float32))
        assert np.array_equal(tf.texture_rgba(), tf.floats_rgba().astype(np.float16).astype(np.float32))
    # .a of RGBt is extinction log(1/(1-a)), of RGBA opacity
    assert abs(tf.floats_rgbt()[255, 3] - np.log(1.0 / (1.0 - 0.8))) < 1e-6
    assert abs(tf.floats_rgba()[255, 3] - 0.8) < 1e-7
    assert tf.floats_rgbt()[20, 3] == 0.0


@needs_ref
@pytest.mark.parametrize("shape", [(4, 4, 4), (9, 7, 5), (3, 17, 11)])
def test_sat_restatement_bit_exact_vs_reference_header(shape):
    """orc_sat_build follows SummedAreaTable3D<double>::BuildSAT's recurrence: the doubles must be identical."""
    rng = np.random.default_rng(2)
    d, h, w = shape
    vox = rng.integers(0, 256, shape).astype(np.uint8)
    tf = bind.TF(*synth.TF_BONSAI)
    lut = tf.ext_lut(1)
    f32, f64 = bind.sat_build(vox, lut, want_f64=True)
    bordered = np.zeros((d + 2, h + 2, w + 2), np.float64)
    bordered[1:-1, 1:-1, 1:-1] = lut[vox].astype(np.float64)
    rf32 = np.empty_like(f32); rf64 = np.empty_like(f64)
    bind.ref().ref_sat3d_double(_p(bordered), w + 2, h + 2, d + 2, _p(rf32), _p(rf64))
    assert np.array_equal(f64.view(np.uint64), rf64.view(np.uint64))
    assert np.array_equal(f32.view(np.uint32), rf32.view(np.uint32))
    # and through the reference-style fill loop used for the CPU baseline
    r2 = np.empty_like(f32)
    bind.ref().ref_sat3d_from_volume(_p(vox), w, h, d, 1, _p(lut), _p(r2))
    assert np.array_equal(r2, rf32)


@needs_ref
def test_sat_integer_vs_reference_template():
    rng = np.random.default_rng(3)
    vox = rng.integers(0, 65536, (6, 5, 9)).astype(np.uint16)
    lut = rng.integers(0, 1 << 20, 65536).astype(np.uint32)
    got = bind.sat_build_u64(vox, lut)
    src = lut[vox].astype(np.uint64)
    want = np.empty_like(src)
    bind.ref().ref_sat3d_u64(_p(src), 9, 5, 6, _p(want))
    assert np.array_equal(got, want)
    assert np.array_equal(got, src.cumsum(0).cumsum(1).cumsum(2))


def test_sat_all_ones_known_answer():
    """SAT of all-ones is (x+1)(y+1)(z+1)."""
    vox = np.ones((5, 6, 7), np.uint8)
    lut = np.zeros(256, np.float32); lut[1] = 1.0
    sat = bind.sat_build(vox, lut)
    z, y, x = np.meshgrid(np.arange(7), np.arange(8), np.arange(9), indexing="ij")
    want = np.minimum(x, 7) * np.minimum(y, 6) * np.minimum(z, 5)
    assert np.array_equal(sat, want.astype(np.float32))


def test_rc1pass_homogeneous_volume_known_answer():
    """Constant density d => every sample has the same (rgb, tau): alpha = 1 - exp(-tau * D) along each ray until the
    0.99 cut."""
    n, W, H = 32, 40, 40
    vox = np.full((n, n, n), 200, np.uint8)
    tf = bind.TF(*synth.TF_THIN)
    eye, center, up = (10.0, 20.0, 90.0), (0, 0, 0), (0, 1, 0)
    cam = bind.camera(eye, center, up, W, H)
    img, ns = bind.rc1pass(vox, tf, cam, W, H, 0.5, count=True)
    hit = ns > 0
    assert hit.sum() > 200
    # texel value and TF lookup as the sampler sees them
    d = np.float32(np.float16(np.float32(200.0 / 255.0)))
    u = d * 256 - 0.5
    i0 = int(np.floor(u)); f = u - i0
    t = tf.texture_rgbt()
    tau = float(t[i0, 3] * (1 - f) + t[i0 + 1, 3] * f)
    # ray length D per pixel from the sample count: (ns-1)*0.5 < D <= ns*0.5 ; alpha is monotone in D
    a_lo = 1.0 - np.exp(-tau * (ns[hit] - 1) * 0.5)
    a_hi = 1.0 - np.exp(-tau * ns[hit] * 0.5)
    a = img[..., 3][hit]
    assert np.all(a >= a_lo - 2e-3) and np.all(a <= a_hi + 2e-3)
    # colour is premultiplied: rgb = alpha * tf.rgb
    rgb = t[i0, :3] * (1 - f) + t[i0 + 1, :3] * f
    assert np.allclose(img[..., :3][hit], a[:, None] * rgb[None, :], atol=2e-3)
    # pixels that miss the box stay (0,0,0,0)
    assert np.all(img[~hit] == 0.0)


def test_rc1pass_early_termination_and_zero_tf():
    n, W, H = 32, 32, 32
    vox = synth.volume_gauss(n)
    eye, center, up = synth.camera_state(0, n)
    cam = bind.camera(eye, center, up, W, H)
    img0, ns0 = bind.rc1pass(vox, bind.TF(*synth.TF_ZERO), cam, W, H, 0.5, count=True)
    assert np.all(img0 == 0.0)
    dense = (np.array([[1, 1, 1, 0], [1, 1, 1, 255]], np.float64), np.array([[0.9, 0], [0.9, 255]], np.float64))
    img1, ns1 = bind.rc1pass(np.full_like(vox, 255), bind.TF(*dense), cam, W, H, 0.5, count=True)
    hit = ns0 > 0
    assert np.all(ns1[hit] <= ns0[hit]) and ns1[hit].max() <= 5      # a = 0.68 per step -> stops after 5 steps
    long_rays = ns0 >= 5                                              # rays long enough to reach the 0.99 cut
    assert long_rays.sum() > 50 and np.all(img1[..., 3][long_rays] > 0.99)


# ---------------------------------------------------------------------------------------------------------------- VCT pre-passes
@pytest.mark.parametrize("shape,dt,tfname", [((8, 12, 16), np.uint8, "bonsai"), ((16, 16, 16), np.uint8, "ramp"), ((9, 7, 10), np.uint8, "sparse"),
                                             ((8, 8, 8), np.uint16, "bonsai")])
def test_vct_prepasses_match_the_reference_code(shape, dt, tfname):
    """VCTPreProcessing::PreProcessSuperVoxels / PreProcessPreIntegrationTable (preprocessingstages.cpp:35-202), the
    reference's own CPU code compiled in place: the float arrays it hands to glTexImage3D (GL_RG16F, one per level) and to
    Texture2D::SetData (GL_R16F) must equal the oracle's after the same fp16 rounding, and so must the maximum deviation."""
    r = bind.ref()
    if r is None or not hasattr(r, "ref_vct_preprocess"):
        pytest.skip("oracle/_ref/libref.so with preprocessingstages.cpp is not available")
    from cpp_volume_rendering_b200 import synth
    rng = np.random.default_rng(31)
    vox = rng.integers(0, np.iinfo(dt).max + 1, shape).astype(dt)
    d, h, w = vox.shape
    bpv = vox.dtype.itemsize
    rgb, a = synth.TFS[tfname]
    mx = 255 if bpv == 1 else 65535
    rgb_s = np.ascontiguousarray(rgb, np.float64).copy(); a_s = np.ascontiguousarray(a, np.float64).copy()
    if bpv == 2:                                                   # control points over the 16-bit range
        rgb_s[:, 3] *= 257.0; a_s[:, 1] *= 257.0
    tf = bind.TF(rgb_s, a_s, mx)
    r.ref_tf_create.restype = C.c_void_p
    rtf = r.ref_tf_create(_p(rgb_s), len(rgb_s), _p(a_s), len(a_s), mx, 0)
    r.ref_vct_preprocess.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_ulonglong, C.c_void_p, C.c_int,
                                     C.POINTER(C.c_double), C.c_void_p, C.c_ulonglong, C.c_void_p]
    want_lut = bpv == 1                                            # the 16-bit LUT is O(65535^2 h): levels only
    olev, odims, oms = bind.vct_supervoxels(vox)
    buf = np.zeros(vox.size * 2 * 2 + 64, np.float32)
    dims = np.zeros((16, 3), np.int32)
    ms = C.c_double(0.0)
    lut = np.zeros(256 * (int(np.ceil(oms)) + 2), np.float32)
    lut_wh = np.zeros(2, np.int32)
    n = r.ref_vct_preprocess(_p(vox), w, h, d, bpv, C.c_void_p(rtf), _p(buf), buf.size, _p(dims), 16, C.byref(ms),
                             _p(lut) if want_lut else None, lut.size, _p(lut_wh))
    r.ref_tf_destroy(C.c_void_p(rtf))
    assert n == len(olev), (n, len(olev))
    assert np.array_equal(dims[:n], odims)
    assert ms.value == oms                                          # fp64, same operation order
    off = 0
    with np.errstate(over="ignore"):
        for l in range(n):
            lw, lh, ld = (int(v) for v in dims[l])
            got = buf[off:off + lw * lh * ld * 2].reshape(ld, lh, lw, 2).astype(np.float16).astype(np.float32)
            off += lw * lh * ld * 2
            assert np.array_equal(got, olev[l]), (l, float(np.abs(got - olev[l]).max()))
        if want_lut:
            opc = np.array([tf.get_opc(i, float(mx)) for i in range(mx + 1)], np.float32)
            olut = bind.vct_preintegration(opc, mx, oms)
            assert (int(lut_wh[0]), int(lut_wh[1])) == (olut.shape[1], olut.shape[0])
            got = lut[:olut.size].reshape(olut.shape).astype(np.float16).astype(np.float32)
            assert np.array_equal(got, olut), float(np.abs(got - olut).max())


@needs_ref
def test_look_at_matches_the_vendored_glm():
    """Camera::LookAt is glm::lookAt of the reference's vendored glm 0.9.5.3: the oracle's restatement (and through
    tests/test_host_cpu.py the host mirror's) must give the same 16 floats for the reference's camera states and random ones."""
    r = bind.ref()
    if not hasattr(r, "ref_glm_look_at"):
        pytest.skip("libref.so predates ref_glm_look_at")
    r.ref_glm_look_at.argtypes = [C.c_void_p] * 4
    rng = np.random.default_rng(4)
    cases = [synth.camera_state(i, 256) for i in range(len(synth.CAMERA_STATES_256))]
    cases += [(tuple(rng.standard_normal(3) * 300), tuple(rng.standard_normal(3) * 20), tuple(rng.standard_normal(3))) for _ in range(200)]
    for eye, center, up in cases:
        e = np.asarray(eye, np.float32); c = np.asarray(center, np.float32); u = np.asarray(up, np.float32)
        a = np.zeros(16, np.float32); b = np.zeros(16, np.float32)
        r.ref_glm_look_at(_p(e), _p(c), _p(u), _p(a))
        bind.orc().orc_look_at(_p(e), _p(c), _p(u), _p(b))
        assert np.array_equal(a, b), (eye, center, up, a, b)
