"""The oracle against the REFERENCE'S OWN GLSL SHADERS, executed on the CPU.

oracle/_ref/librefglsl.so holds the reference's compute shaders compiled as C++ from where they lie under
/root/reference (oracle/glsl_cpu/: qualifier-only rewriting + a GLSL execution environment).  Each test plays the
reference's renderer class -- it sets the uniforms by name like Update() / CreateRenderingPass() do, binds textures
and the output image, dispatches -- and requires the oracle's restatement of the same shader to produce the same
fp16 image.  Samplers and built-ins are shared with the oracle (oracle_common.h), so what is pinned is the shader
LOGIC: ray set-up, loop structure, compositing, cone / shell / SAT-box walks, termination, quirks.
Expected: bit-identical (checked as max abs diff == 0 where nothing else is stated).  CPU only."""

import numpy as np
import pytest

from cpp_volume_rendering_b200 import capi, synth
from oracle import bind, refglsl


@pytest.fixture(scope="module")
def rg(built):
    if refglsl.lib() is None:
        pytest.skip("oracle/_ref/librefglsl.so is not available (needs /root/reference to build)")
    return refglsl


_pyramid_levels = refglsl.pyramid_levels


def _same(img, ref, what):
    assert np.array_equal(np.isfinite(img), np.isfinite(ref)), f"{what}: finite masks differ"
    fin = np.isfinite(ref)
    diff = float(np.abs(img[fin] - ref[fin]).max()) if fin.any() else 0.0
    assert diff == 0.0, f"{what}: max abs diff {diff} over {(np.abs(img - ref) > 0).sum()} values"
    assert (ref[..., 3] > 0).sum() > 50, f"{what}: the case renders (almost) nothing"


class _gradient:
    """Bind TexVolumeGradient for the oracle for the duration of a call."""

    def __init__(self, vox, mode):
        self.grad = bind.gradient_build(vox, mode) if mode else None

    def __enter__(self):
        bind.set_gradient(self.grad)
        return self.grad

    def __exit__(self, *a):
        bind.set_gradient(None)


# ---------------------------------------------------------------------------------------------------------------- rc1pass
RC1_CASES = [
    ("gauss32-bonsai", lambda: synth.volume_gauss(32), "bonsai", 0, 72, 56, 0.5, (1.0, 1.0, 1.0), 0),
    ("noise28-ramp-cam4", lambda: synth.volume_noise(28), "ramp", 4, 64, 64, 0.3, (1.0, 1.0, 1.0), 0),
    ("boxes24-sparse-scaled", lambda: synth.volume_boxes(24), "sparse", 2, 48, 64, 0.7, (1.0, 0.5, 2.0), 0),
    ("gauss24-u16-thin-phong", lambda: synth.volume_gauss(24, np.uint16), "thin", 1, 56, 56, 0.5, (1.0, 1.0, 1.0), 1),
    ("noise24-bonsai-phong-fd", lambda: synth.volume_noise(24), "bonsai", 5, 56, 48, 0.5, (1.0, 1.0, 1.0), 2),
]


@pytest.mark.parametrize("name,mk,tfname,cam_id,W,H,step,scale,phong", RC1_CASES, ids=[c[0] for c in RC1_CASES])
def test_rc1pass_oracle_equals_reference_shader(rg, name, mk, tfname, cam_id, W, H, step, scale, phong):
    vox = mk()
    n = vox.shape[0]
    tf = bind.TF(*synth.TFS[tfname])
    eye, center, up = synth.camera_state(cam_id, n)
    cam = bind.camera(eye, center, up, W, H)
    light = bind.copy_struct(capi.default_lighting(light_pos=synth.light_position(n)), bind.OrcLighting)
    light.apply_phong = 1 if phong else 0
    with _gradient(vox, phong) as grad:
        ref = bind.rc1pass_lit(vox, tf, cam, light, W, H, step, scale)
    _same(rg.run_rc1pass(vox, tf, cam, light, W, H, step, scale, grad), ref, name)


# ---------------------------------------------------------------------------------------------------------------- iso
@pytest.mark.parametrize("phong", [0, 1])
def test_iso_oracle_equals_reference_shader(rg, phong):
    n, W, H = 32, 72, 64
    z, y, x = np.mgrid[0:n, 0:n, 0:n].astype(np.float64)
    r = np.sqrt((x - n / 2 + 0.5) ** 2 + (y - n / 2 + 0.5) ** 2 + (z - n / 2 + 0.5) ** 2)
    vox = np.clip(255.0 * (1.0 - r / (0.5 * n)), 0, 255).astype(np.uint8)
    eye, center, up = synth.camera_state(3, n)
    cam = bind.camera(eye, center, up, W, H)
    light = bind.copy_struct(capi.default_lighting(light_pos=synth.light_position(n)), bind.OrcLighting)
    light.apply_phong = phong
    prm = bind.copy_struct(capi.default_iso_params(), bind.OrcIsoParams)
    prm.isovalue = 0.4
    with _gradient(vox, phong) as grad:
        ref = bind.iso(vox, cam, light, prm, W, H)
    _same(rg.run_iso(vox, cam, light, prm, W, H, grad), ref, f"iso phong={phong}")


# ---------------------------------------------------------------------------------------------------------------- EBS
EBS_CASES = [
    ("gauss32-defaults", lambda: synth.volume_gauss(32), "bonsai", 0, 64, 48, dict(), 0),
    ("noise32-defaults", lambda: synth.volume_noise(32), "bonsai", 0, 64, 48, dict(), 0),
    ("noise28-ao-only-8shells", lambda: synth.volume_noise(28), "ramp", 4, 56, 56, dict(apply_shadow=0, amb_occ_shells=8, amb_occ_radius=1.5), 0),
    ("boxes28-shadow-only-wide", lambda: synth.volume_boxes(28), "sparse", 1, 56, 48,
     dict(apply_occlusion=0, sdw_cone_angle_rad=np.float32(np.deg2rad(8.0)), sdw_sample_interval=3.0), 0),
    ("gauss24-u16-directional-light", lambda: synth.volume_gauss(24, np.uint16), "thin", 2, 48, 48, dict(type_of_shadow=1), 0),
    ("gauss24-light-from-x", lambda: synth.volume_gauss(24), "bonsai", 3, 48, 48, dict(), 0),
    ("noise24-phong", lambda: synth.volume_noise(24), "bonsai", 0, 48, 40, dict(), 1),
]


@pytest.mark.parametrize("name,mk,tfname,cam_id,W,H,opts,phong", EBS_CASES, ids=[c[0] for c in EBS_CASES])
def test_ebs_oracle_equals_reference_shader(rg, name, mk, tfname, cam_id, W, H, opts, phong):
    vox = mk()
    n = vox.shape[0]
    tf = bind.TF(*synth.TFS[tfname])
    eye, center, up = synth.camera_state(cam_id, n)
    cam = bind.camera(eye, center, up, W, H)
    sat = bind.sat_build(vox, tf.ext_lut(vox.dtype.itemsize))
    lpos = synth.light_position(n) if "light-from-x" not in name else (3.0 * n, 0.2 * n, -0.1 * n)
    light = bind.copy_struct(capi.default_lighting(light_pos=lpos, forward=synth.camera_forward(eye, center)), bind.OrcLighting)
    light.apply_phong = phong
    prm = bind.copy_struct(capi.default_ebs_params(float(np.sqrt(3.0) * n)), bind.OrcEbsParams)
    for k, v in opts.items():
        setattr(prm, k, v)
    with _gradient(vox, phong) as grad:
        ref = bind.ebs(vox, tf, sat, cam, light, prm, W, H)
    _same(rg.run_ebs(vox, tf, sat, cam, light, prm, W, H, grad), ref, name)


# ---------------------------------------------------------------------------------------------------------------- DOS
def _dos_cones(diag, occ, sdw):
    po = bind.cone_params(occ[0], occ[1], 0.5 * diag, occ[2])
    ps = bind.cone_params(sdw[0], sdw[1], 0.75 * diag, sdw[2])
    so, oo = bind.cone_sampler(po, 1.0)
    ss, os_ = bind.cone_sampler(ps, 1.0)
    return bind.dos_cone(so, oo, po), bind.dos_cone(ss, os_, ps)


DOS_CASES = [
    ("gauss32-ao", lambda: synth.volume_gauss(32), "bonsai", 0, 64, 56, 0.5, dict(), 0),
    ("noise28-ao+shadow", lambda: synth.volume_noise(28), "ramp", 4, 56, 56, 0.5, dict(apply_shadow=1), 0),
    ("boxes28-shadow-only-directional", lambda: synth.volume_boxes(28), "sparse", 1, 56, 48, 0.6, dict(apply_shadow=1, apply_occlusion=0, type_of_shadow=2), 0),
    ("gauss28-spot", lambda: synth.volume_gauss(28), "bonsai", 0, 56, 56, 0.5, dict(apply_shadow=1, type_of_shadow=1), 0),
    ("gauss24-u16-7rays", lambda: synth.volume_gauss(24, np.uint16), "thin", 2, 48, 48, 0.9, dict(apply_shadow=1, occ=(40.0, 2, 0.5), sdw=(10.0, 1, 1.0)), 0),
    ("noise24-phong", lambda: synth.volume_noise(24), "bonsai", 0, 48, 40, 0.5, dict(apply_shadow=1), 1),
]


@pytest.mark.parametrize("name,mk,tfname,cam_id,W,H,step,opts,phong", DOS_CASES, ids=[c[0] for c in DOS_CASES])
def test_dos_oracle_equals_reference_shader(rg, name, mk, tfname, cam_id, W, H, step, opts, phong):
    opts = dict(opts)
    vox = mk()
    n = vox.shape[0]
    tf = bind.TF(*synth.TFS[tfname])
    eye, center, up = synth.camera_state(cam_id, n)
    cam = bind.camera(eye, center, up, W, H)
    occ, sdw = _dos_cones(float(np.sqrt(3.0) * n), opts.pop("occ", (20.0, 1, 0.35)), opts.pop("sdw", (0.5, 0, 1.0)))
    prm = bind.copy_struct(capi.default_dos_params(step, spot_angle_deg=20.0), bind.OrcDosParams)
    for k, v in opts.items():
        setattr(prm, k, v)
    light = bind.copy_struct(capi.default_lighting(light_pos=synth.light_position(n), forward=synth.camera_forward(eye, center),
                                                   up=(0.0, 1.0, 0.0), right=(1.0, 0.0, 0.0)), bind.OrcLighting)
    light.apply_phong = phong
    pyr, dims = bind.extcoef_build(vox, tf, 1.0, (16, 16, 16))
    with _gradient(vox, phong) as grad:
        ref = bind.dos(vox, tf, pyr, dims, cam, light, occ, sdw, prm, W, H)
    _same(rg.run_dos(vox, tf, pyr, dims, cam, light, occ, sdw, prm, W, H, grad), ref, name)


# ---------------------------------------------------------------------------------------------------------------- VCT
VCT_CASES = [
    ("gauss32-bonsai", lambda: synth.volume_gauss(32), "bonsai", 0, 64, 56, 0.5, dict(), 0),
    ("noise28-ramp-rate", lambda: synth.volume_noise(28), "ramp", 4, 56, 56, 0.5, dict(cone_step_increase_rate=1.1), 0),
    ("boxes32-sparse-nocorr", lambda: synth.volume_boxes(32), "sparse", 1, 56, 48, 0.6, dict(apply_opacity_correction=0), 0),
    ("gauss24-shadow-only", lambda: synth.volume_gauss(24), "bonsai", 5, 48, 48, 0.5, dict(apply_occlusion=0, cone_number_of_samples=20), 0),
    ("noise24-phong", lambda: synth.volume_noise(24), "bonsai", 0, 48, 40, 0.5, dict(), 1),
]


@pytest.mark.parametrize("name,mk,tfname,cam_id,W,H,step,opts,phong", VCT_CASES, ids=[c[0] for c in VCT_CASES])
def test_vct_oracle_equals_reference_shader(rg, name, mk, tfname, cam_id, W, H, step, opts, phong):
    vox = mk()
    n = vox.shape[0]
    tf = bind.TF(*synth.TFS[tfname])
    eye, center, up = synth.camera_state(cam_id, n)
    cam = bind.camera(eye, center, up, W, H)
    opc = np.array([tf.get_opc(i, 255.0) for i in range(256)], np.float32)
    levels, dims, ms = bind.vct_supervoxels(vox)
    lut = bind.vct_preintegration(opc, 255, ms)
    prm = bind.copy_struct(capi.default_vct_params(255.0, ms, step), bind.OrcVctParams)
    for k, v in opts.items():
        setattr(prm, k, v)
    light = bind.copy_struct(capi.default_lighting(light_pos=synth.light_position(n)), bind.OrcLighting)
    light.apply_phong = phong
    with _gradient(vox, phong) as grad:
        ref = bind.vct(vox, tf, levels, dims, lut, cam, light, prm, W, H)
    _same(rg.run_vct(vox, tf, levels, lut, cam, light, prm, W, H, grad), ref, name)


# ---------------------------------------------------------------------------------------------------------------- GT
GT_CASES = [
    ("gauss24-occ4-sdw4", lambda: synth.volume_gauss(24), "bonsai", 0, 48, 40, 0.5, 4, 4, dict(), 0),
    ("noise20-occ-only", lambda: synth.volume_noise(20), "ramp", 4, 40, 40, 0.5, 5, 2, dict(apply_shadow=0), 0),
    ("boxes24-shadow-directional", lambda: synth.volume_boxes(24), "sparse", 1, 40, 40, 0.7, 2, 4, dict(apply_occlusion=0, shadow_type=2), 0),
    ("gauss20-spot", lambda: synth.volume_gauss(20), "bonsai", 0, 40, 32, 0.5, 2, 3, dict(shadow_type=1), 0),
    ("gauss20-u16-phong", lambda: synth.volume_gauss(20, np.uint16), "thin", 2, 40, 40, 0.4, 3, 3, dict(), 1),
]


@pytest.mark.parametrize("name,mk,tfname,cam_id,W,H,step,nocc,nsdw,opts,phong", GT_CASES, ids=[c[0] for c in GT_CASES])
def test_gt_oracle_equals_reference_shader_driven_to_convergence(rg, name, mk, tfname, cam_id, W, H, step, nocc, nsdw, opts, phong):
    """The reference converges this renderer by re-dispatching gt_ray_marching.comp (one primary sample per dispatch and
    pixel, colour and ray parameter kept in an rgba16f and an rg16f image) until every pixel's state says done
    (crtgtrenderer.cpp:262-325).  The same loop is run here on the shader; the oracle returns the converged frame.

    A reference quirk this harness found: the ray parameter is kept in an rg16f image, and when its fp16 rounding lands
    at or beyond D the next dispatch's loop `for (s = ifrag.x; s < D;)` runs zero times and never writes the "done" flag
    (gt_ray_marching.comp:413-470).  The frame is converged (a further dispatch changes nothing) but m_frame_outdated
    stays true for ever.  run_gt stops at that fixed point and reports the pixels; the oracle and the product end such rays."""
    vox = mk()
    n = vox.shape[0]
    tf = bind.TF(*synth.TFS[tfname])
    eye, center, up = synth.camera_state(cam_id, n)
    cam = bind.camera(eye, center, up, W, H)
    occ, sdw = capi.host_gt_ray_tables(nocc, 90.0, nsdw, 10.0)
    prm = bind.copy_struct(capi.default_gt_params(float(np.sqrt(3.0) * n), nocc, nsdw, step), bind.OrcGtParams)
    for k, v in opts.items():
        setattr(prm, k, v)
    fwd = synth.camera_forward(eye, center)
    light = bind.copy_struct(capi.default_lighting(light_pos=synth.light_position(n), forward=tuple(-f for f in fwd), up=(0.0, 1.0, 0.0),
                                                   right=(1.0, 0.0, 0.0)), bind.OrcLighting)
    light.apply_phong = phong
    with _gradient(vox, phong) as grad:
        ref = bind.gt(vox, tf, cam, light, prm, occ, sdw, W, H)
    out, dispatches, stalled = rg.run_gt(vox, tf, cam, light, prm, occ, sdw, W, H, grad)
    assert dispatches > 10 and stalled <= 0.02 * W * H, (dispatches, stalled)
    _same(out, ref, f"{name} after {dispatches} dispatches")


@pytest.mark.parametrize("cam_id,wh,shape,scale", [(0, (64, 48), (20, 20, 20), (1.0, 1.0, 1.0)), (3, (50, 70), (12, 20, 30), (1.0, 0.5, 2.0)),
                                                   (1, (40, 40), (16, 16, 16), (1.0, 1.0, 1.0)), (5, (33, 57), (9, 31, 14), (2.0, 1.0, 1.25))])
def test_gt_bounding_box_placeholder_oracle_equals_reference_shader(rg, cam_id, wh, shape, scale):
    """RedrawCube (crtgtrenderer.cpp:327-338, vol_intersection.comp): what rc1pcrtgt shows while "Show Generated Frame
    Texture" is off, its default."""
    W, H = wh
    eye, center, up = synth.camera_state(cam_id, max(shape))
    cam = bind.camera(eye, center, up, W, H)
    a, b = rg.run_gt_cube(shape, cam, W, H, scale), bind.gt_cube(shape, cam, W, H, scale)
    assert np.array_equal(a, b) and (b[..., 3] > 0).sum() > 100
    assert set(np.unique(b)) <= {0.0, 1.0} and np.all(b[..., :3].sum(-1) == b[..., 3])      # one face colour per hit pixel
    inside = bind.camera((1.0, 2.0, 3.0), (0, 0, -40), (0, 1, 0), W, H)                       # eye inside the box: tnear clamps to 0
    assert np.array_equal(rg.run_gt_cube(shape, inside, W, H, scale), bind.gt_cube(shape, inside, W, H, scale))


# ---------------------------------------------------------------------------------------------------------------- pyramid
@pytest.mark.parametrize("shape,dt,res,tfname,sigma0", [((20, 20, 20), np.uint8, (16, 16, 16), "bonsai", 1.0),
                                                       ((12, 18, 24), np.uint16, (8, 12, 16), "ramp", 1.0),
                                                       ((16, 16, 16), np.uint8, (16, 16, 16), "sparse", 1.5)])
def test_extinction_pyramid_oracle_equals_reference_shaders(rg, shape, dt, res, tfname, sigma0):
    """ExtinctionCoefficientVolume::GenerateExtinctionCoefficientVolumeAnySize + TransformTexOpacityToExtinction
    (extcoefvolumegenerator.cpp:230-408) replayed on the reference's three compute shaders (refglsl.run_extcoef_pyramid)."""
    vox = synth.volume_noise(max(shape), dt)[:shape[0], :shape[1], :shape[2]].copy()
    tf = bind.TF(*synth.TFS[tfname])
    pyr, dims = bind.extcoef_build(vox, tf, sigma0, res)
    want = _pyramid_levels(pyr, dims)
    assert [tuple(int(v) for v in d) for d in dims] == rg.mip_dims(*res)
    levels = rg.run_extcoef_pyramid(vox, tf, sigma0, res)
    for i, (a, b) in enumerate(zip(levels, want)):
        assert a.shape == b.shape
        assert np.array_equal(a, b), (i, float(np.abs(a - b).max()), int((a != b).sum()))
    assert float(levels[0].max()) > 0.01


@pytest.mark.parametrize("seed", [9, 10])
def test_same_size_extinction_pyramid_oracle_equals_reference_shaders(rg, seed):
    """GenerateExtinctionCoefficientVolumeSameSize (extcoefvolumegenerator.cpp:92-228, the branch taken without a custom
    resolution): gen_extcoefvol_samesize.comp places its taps with the volume's own VoxelSize, which for arbitrary voxel scales
    is not bit-identical to grid size / resolution -- the any-size arithmetic differs by one fp16 ulp in rare texels (found by
    this sweep; oracle and k_extcoef_level now carry the same-size form)."""
    rng = np.random.default_rng(seed)
    for _ in range(10):
        shape = tuple(int(v) for v in rng.integers(6, 15, 3))
        scale = tuple(float(np.float32(v)) for v in rng.uniform(0.2, 3.0, 3))
        vox = np.ascontiguousarray(synth.volume_noise(max(shape))[:shape[0], :shape[1], :shape[2]])
        tf = bind.TF(*synth.TFS[str(rng.choice(["bonsai", "ramp", "sparse"]))])
        pyr, dims = bind.extcoef_build(vox, tf, 1.0, None, scale)
        for i, (a, b) in enumerate(zip(rg.run_extcoef_pyramid_samesize(vox, tf, 1.0, scale), _pyramid_levels(pyr, dims))):
            assert np.array_equal(a, b), (shape, scale, i, int((a != b).sum()))


# ---------------------------------------------------------------------------------------------------------------- Sobel
@pytest.mark.parametrize("shape,dt", [((12, 14, 16), np.uint8), ((10, 10, 10), np.uint16)])
def test_compute_shader_sobel_gradient_oracle_equals_reference_shader(rg, shape, dt):
    """DataManager::GenerateStructuredGradientTexture, compute-shader branch (datamanager.cpp:623-717)."""
    vox = synth.volume_noise(max(shape), dt)[:shape[0], :shape[1], :shape[2]].copy()
    want = bind.gradient_build(vox, bind.GRADIENT_COMPUTE_SHADER_SOBEL)
    got = rg.run_sobel(vox)
    assert np.array_equal(got, want), float(np.abs(got - want).max())
    assert float(np.abs(want).max()) > 0.05


# ---------------------------------------------------------------------------------------------------------------- light caches
LC_CASES = [
    ("gauss28-ao", lambda: synth.volume_gauss(28), "bonsai", 0, (10, 10, 10), dict()),
    ("noise28-ao+shadow", lambda: synth.volume_noise(28), "ramp", 4, (12, 10, 8), dict(apply_shadow=1)),
    ("gauss24-shadow-only-directional", lambda: synth.volume_gauss(24), "bonsai", 1, (8, 8, 8), dict(apply_shadow=1, apply_occlusion=0, type_of_shadow=2)),
    ("gauss24-spot-7rays", lambda: synth.volume_gauss(24), "bonsai", 2, (8, 8, 8), dict(apply_shadow=1, type_of_shadow=1, occ=(40.0, 2, 0.5), sdw=(10.0, 1, 1.0))),
]


@pytest.mark.parametrize("name,mk,tfname,cam_id,res,opts", LC_CASES, ids=[c[0] for c in LC_CASES])
def test_dos_light_cache_and_object_space_march_equal_reference_shaders(rg, name, mk, tfname, cam_id, res, opts):
    """K6 rc1pdosct/lightcachecomputation.comp as PreComputeLightCache dispatches it (dosrcrenderer.cpp:555-657,700-735),
    then _common_shaders/obj_ray_marching.comp over the cache (Update :134-141, CreateRenderingPass :659-700)."""
    opts = dict(opts)
    vox = mk()
    n = vox.shape[0]
    W, H, step = 56, 48, 0.5
    tf = bind.TF(*synth.TFS[tfname])
    eye, center, up = synth.camera_state(cam_id, n)
    _, up_v, _ = rg.camera_vectors(eye, center, up)
    occ, sdw = _dos_cones(float(np.sqrt(3.0) * n), opts.pop("occ", (20.0, 1, 0.35)), opts.pop("sdw", (0.5, 0, 1.0)))
    prm = bind.copy_struct(capi.default_dos_params(step, spot_angle_deg=20.0), bind.OrcDosParams)
    for k, v in opts.items():
        setattr(prm, k, v)
    light = bind.copy_struct(capi.default_lighting(light_pos=synth.light_position(n), forward=synth.camera_forward(eye, center),
                                                   up=(0.0, 1.0, 0.0), right=(1.0, 0.0, 0.0)), bind.OrcLighting)
    pyr, dims = bind.extcoef_build(vox, tf, 1.0, (16, 16, 16))
    want = bind.dos_light_cache(vox.shape, pyr, dims, eye, tuple(float(v) for v in up_v), light, occ, sdw, prm, res)
    cache = rg.run_dos_light_cache(vox, tf, pyr, dims, eye, center, up, light, occ, sdw, prm, res)
    assert np.array_equal(cache, want), float(np.abs(cache - want).max())
    assert 0.0 <= want.min() and want.max() <= 1.0 and want.std() > 1e-3
    # the march over the cache
    cam = bind.camera(eye, center, up, W, H)
    ref = bind.obj_march(vox, tf, cam, light.ka, light.kd, prm.apply_occlusion, prm.apply_shadow, step, want, W, H)
    _same(rg.run_obj(vox, tf, cam, light, prm.apply_occlusion, prm.apply_shadow, step, cache, W, H), ref, name + " (object-space march)")


@pytest.mark.parametrize("occ,sdw,mode", [(1, 1, 1), (1, 0, 2), (0, 1, 1)])
def test_object_space_march_with_gradient_phong_equals_reference_shader(rg, occ, sdw, mode):
    """obj_ray_marching.comp's ApplyPhongShading branch (:236-257) over a light cache (values need not come from a cone
    shader for this: a smooth synthetic RG cache exercises the same code)."""
    n, W, H, step = 24, 56, 48, 0.5
    vox = synth.volume_noise(n)
    tf = bind.TF(*synth.TF_BONSAI)
    eye, center, up = synth.camera_state(4, n)
    cam = bind.camera(eye, center, up, W, H)
    light = bind.copy_struct(capi.default_lighting(light_pos=synth.light_position(n)), bind.OrcLighting)
    light.apply_phong = 1
    z, y, x = np.mgrid[0:6, 0:8, 0:10].astype(np.float32)
    cache = np.stack([0.3 + 0.07 * x, 1.0 - 0.1 * y + 0.02 * z], -1).astype(np.float16).astype(np.float32)
    with _gradient(vox, mode) as grad:
        ref = bind.obj_march_lit(vox, tf, cam, light, occ, sdw, step, cache, W, H)
    _same(rg.run_obj(vox, tf, cam, light, occ, sdw, step, cache, W, H, grad), ref, f"obj phong occ={occ} sdw={sdw}")
    light.apply_phong = 0
    plain = bind.obj_march(vox, tf, cam, light.ka, light.kd, occ, sdw, step, cache, W, H)
    if sdw:                                                      # without the shadow term kd = ks = 0 and the branch reduces to the plain mix
        assert np.abs(plain - ref).max() > 1e-3


@pytest.mark.parametrize("name,mk,tfname,res,opts", [
    ("gauss24", lambda: synth.volume_gauss(24), "bonsai", (10, 10, 10), dict()),
    ("noise24-ao-only", lambda: synth.volume_noise(24), "ramp", (8, 6, 12), dict(apply_shadow=0, amb_occ_shells=6)),
    ("boxes24-shadow-only-directional", lambda: synth.volume_boxes(24), "sparse", (8, 8, 8), dict(apply_occlusion=0, type_of_shadow=1)),
], ids=lambda v: v if isinstance(v, str) else None)
def test_ebs_light_cache_oracle_equals_reference_shader(rg, name, mk, tfname, res, opts):
    """K9 rc1pextbsd/lightcachecomputation.comp as PreComputeLightCache dispatches it (ebsrenderer.cpp:441-555)."""
    vox = mk()
    n = vox.shape[0]
    tf = bind.TF(*synth.TFS[tfname])
    eye, center, up = synth.camera_state(0, n)
    sat = bind.sat_build(vox, tf.ext_lut(vox.dtype.itemsize))
    light = bind.copy_struct(capi.default_lighting(light_pos=synth.light_position(n), forward=synth.camera_forward(eye, center)), bind.OrcLighting)
    prm = bind.copy_struct(capi.default_ebs_params(float(np.sqrt(3.0) * n)), bind.OrcEbsParams)
    for k, v in opts.items():
        setattr(prm, k, v)
    want = bind.ebs_light_cache(vox.shape, sat, light, prm, res)
    cache = rg.run_ebs_light_cache(vox, tf, sat, eye, light, prm, res)
    fin = np.isfinite(want)
    assert np.array_equal(np.isfinite(cache), fin) and np.array_equal(cache[fin], want[fin]), float(np.abs(cache[fin] - want[fin]).max())
    assert want[fin].std() > 1e-3


@pytest.mark.parametrize("name,mk,tfname,res,opts", [
    ("gauss24", lambda: synth.volume_gauss(24), "bonsai", (10, 10, 10), dict()),
    ("boxes24-rate-20steps", lambda: synth.volume_boxes(24), "sparse", (8, 6, 12), dict(cone_step_increase_rate=1.1, cone_number_of_samples=20)),
], ids=lambda v: v if isinstance(v, str) else None)
def test_vct_light_cache_oracle_equals_reference_shader(rg, name, mk, tfname, res, opts):
    """K13 rc1pvctsg/lightcachecomputation.comp as PreComputeLightCache dispatches it (vctrenderer.cpp:393-515)."""
    vox = mk()
    n = vox.shape[0]
    tf = bind.TF(*synth.TFS[tfname])
    opc = np.array([tf.get_opc(i, 255.0) for i in range(256)], np.float32)
    levels, dims, ms = bind.vct_supervoxels(vox)
    lut = bind.vct_preintegration(opc, 255, ms)
    prm = bind.copy_struct(capi.default_vct_params(255.0, ms, 0.5), bind.OrcVctParams)
    for k, v in opts.items():
        setattr(prm, k, v)
    light = bind.copy_struct(capi.default_lighting(light_pos=synth.light_position(n)), bind.OrcLighting)
    want = bind.vct_light_cache(vox.shape, levels, dims, lut, light, prm, res)
    cache = rg.run_vct_light_cache(vox, tf, levels, lut, light, prm, res)
    assert np.array_equal(cache, want), float(np.abs(cache - want).max())
    assert want[..., 1].std() > 1e-3


# ---------------------------------------------------------------------------------------------------------------- frame filters
def _frame(rng, w, h):
    img = rng.random((h, w, 4)).astype(np.float32)
    img[..., :3] *= img[..., 3:4]
    return img.astype(np.float16).astype(np.float32)


def test_multisample_filter_oracle_equals_reference_shader(rg):
    src = _frame(np.random.default_rng(5), 64, 48)
    assert np.array_equal(rg.run_frame_filter(src, 32, 24, 1), bind.frame_filter(src, 32, 24, 1))


@pytest.mark.parametrize("kernel", range(6), ids=refglsl.FILTER_KERNELS)
@pytest.mark.parametrize("direction,src_wh,dst_wh", [("down", (66, 50), (33, 25)), ("down", (40, 56), (24, 30)),
                                                    ("up", (24, 30), (48, 60)), ("up", (31, 17), (50, 40))])
def test_scaling_filters_oracle_equals_reference_shaders(rg, kernel, direction, src_wh, dst_wh):
    """DrawHigherResolutionWithDownScale / DrawLowerResolutionWithUpScale (renderoutputframe.cpp:305-540): the kernel file
    linked with down- / upscaling_filter.comp; cardinal kernels run their digital filter on the result (down) or, in
    place, on the rendered frame first (up)."""
    src = _frame(np.random.default_rng(100 + kernel), *src_wh)
    pass_id = 2 if direction == "down" else 3
    want = bind.frame_filter(src, dst_wh[0], dst_wh[1], pass_id, kernel)
    out = rg.run_frame_filter(src, dst_wh[0], dst_wh[1], pass_id, kernel)
    assert np.array_equal(np.isfinite(out), np.isfinite(want))
    assert np.array_equal(out, want, equal_nan=True), float(np.nanmax(np.abs(out - want)))


# ---------------------------------------------------------------------------------------------------------------- full size
def test_config1_full_size_oracle_equals_reference_shader(rg):
    """BASELINE config 1 as benchmarked (256^3 u8 V-gauss + bonsai TF at 768 x 768, step 0.5): 7.8e7 loop iterations of the
    reference's ray_marching_1p.comp on the CPU; the oracle's frame must be bit-identical."""
    n, W, H = 256, 768, 768
    vox = synth.volume_gauss(n)
    tf = bind.TF(*synth.TF_BONSAI)
    eye, center, up = synth.camera_state(0, n)
    cam = bind.camera(eye, center, up, W, H)
    ref, ns = bind.rc1pass(vox, tf, cam, W, H, 0.5, count=True)
    img = rg.run_rc1pass(vox, tf, cam, bind.OrcLighting(), W, H, 0.5)
    assert (ns > 0).sum() == 240599                       # rays that hit the box at config 1
    _same(img, ref, "config 1")


# ---------------------------------------------------------------------------------------------------------------- random sweep
@pytest.mark.parametrize("seed", [11, 12, 13])
def test_random_configurations_oracle_equals_reference_shaders(rg, seed):
    """Ten random scenes per seed through all five marchers: ragged u8 / u16 volumes, anisotropic voxel scales, cameras
    outside / inside the volume / looking away from it, random lights, step sizes, shadow types, cone set-ups, shell
    counts, gradient modes.  Every frame of the oracle must equal the reference shader's, value for value."""
    rng = np.random.default_rng(seed)
    N = 10
    failures = []
    def cmp(name, a, b, info):
        fin = np.isfinite(a) & np.isfinite(b)
        same_mask = np.array_equal(np.isfinite(a), np.isfinite(b))
        d = float(np.abs(a[fin] - b[fin]).max()) if fin.any() else 0.0
        if d != 0.0 or not same_mask:
            failures.append((name, d, int((a[fin] != b[fin]).sum()), same_mask, info))
    for it in range(N):
        shape = tuple(int(v) for v in rng.integers(10, 28, 3))
        dt = np.uint16 if rng.random() < 0.3 else np.uint8
        kind = rng.choice(["gauss", "noise", "boxes"])
        n = max(shape)
        vox = {"gauss": synth.volume_gauss, "noise": synth.volume_noise, "boxes": synth.volume_boxes}[kind](n, dt) if kind != "boxes" else synth.volume_boxes(n)
        if kind == "boxes" and dt == np.uint16: vox = (vox.astype(np.uint16) * 257)
        vox = np.ascontiguousarray(vox[:shape[0], :shape[1], :shape[2]])
        tfname = rng.choice(["bonsai", "ramp", "sparse", "thin"])
        tf = bind.TF(*synth.TFS[tfname])
        scale = (1.0, 1.0, 1.0) if rng.random() < 0.4 else tuple(float(v) for v in rng.choice([0.5, 1.0, 1.25, 2.0], 3))
        G = np.array([shape[2] * scale[0], shape[1] * scale[1], shape[0] * scale[2]])
        diag = float(np.sqrt((G ** 2).sum()))
        # camera: outside, inside, or looking away
        mode = rng.choice(["out", "in", "away"], p=[0.7, 0.2, 0.1])
        dirv = rng.standard_normal(3); dirv /= np.linalg.norm(dirv)
        eye = tuple(dirv * diag * (rng.uniform(0.7, 1.6) if mode != "in" else rng.uniform(0.0, 0.25)))
        center = tuple(rng.standard_normal(3) * 0.1 * diag) if mode != "away" else tuple(np.array(eye) * 2.0)
        up = (0.0, 1.0, 0.0) if abs(dirv[1]) < 0.9 else (0.0, 0.0, 1.0)
        W, H = int(rng.integers(20, 48)), int(rng.integers(20, 48))
        cam = bind.camera(eye, center, up, W, H)
        step = float(rng.choice([0.25, 0.5, 0.7, 1.3]))
        lpos = tuple(rng.standard_normal(3) * diag * rng.uniform(0.2, 2.0))
        fwd = synth.camera_forward(eye, center)
        phong = int(rng.integers(0, 3))
        info = dict(it=it, shape=shape, dt=dt.__name__, kind=str(kind), tf=str(tfname), scale=scale, mode=str(mode), W=W, H=H, step=step, phong=phong)
        grad = bind.gradient_build(vox, phong) if phong else None
        light = bind.copy_struct(capi.default_lighting(light_pos=lpos, forward=fwd, up=(0.0, 1.0, 0.0), right=(1.0, 0.0, 0.0)), bind.OrcLighting)
        light.apply_phong = 1 if phong else 0
        bind.set_gradient(grad)
        try:
            # rc1pass
            cmp("rc1pass", rg.run_rc1pass(vox, tf, cam, light, W, H, step, scale, grad), bind.rc1pass_lit(vox, tf, cam, light, W, H, step, scale), info)
            # ebs
            sat = bind.sat_build(vox, tf.ext_lut(vox.dtype.itemsize))
            prm = bind.copy_struct(capi.default_ebs_params(diag, step), bind.OrcEbsParams)
            prm.type_of_shadow = int(rng.integers(0, 2)); prm.amb_occ_shells = int(rng.integers(1, 12)); prm.amb_occ_radius = float(rng.choice([0.5, 1.0, 2.0]))
            prm.sdw_cone_angle_rad = np.float32(np.deg2rad(rng.choice([0.5, 1.0, 5.0, 15.0]))); prm.apply_occlusion = int(rng.random() < 0.8); prm.apply_shadow = int(rng.random() < 0.8)
            cmp("ebs", rg.run_ebs(vox, tf, sat, cam, light, prm, W, H, grad, scale), bind.ebs(vox, tf, sat, cam, light, prm, W, H, scale), dict(info, tos=prm.type_of_shadow, occ=prm.apply_occlusion, sdw=prm.apply_shadow))
            # dos
            occ, sdw = _dos_cones(diag, (float(rng.choice([10.0, 20.0, 40.0])), int(rng.integers(0, 3)), 0.35), (float(rng.choice([0.5, 5.0, 10.0])), int(rng.integers(0, 3)), 1.0))
            dprm = bind.copy_struct(capi.default_dos_params(step, spot_angle_deg=20.0), bind.OrcDosParams)
            dprm.apply_shadow = int(rng.random() < 0.7); dprm.apply_occlusion = int(rng.random() < 0.8); dprm.type_of_shadow = int(rng.integers(0, 3))
            pyr, dims = bind.extcoef_build(vox, tf, 1.0, (8, 8, 8), scale)
            cmp("dos", rg.run_dos(vox, tf, pyr, dims, cam, light, occ, sdw, dprm, W, H, grad, scale), bind.dos(vox, tf, pyr, dims, cam, light, occ, sdw, dprm, W, H, scale),
                dict(info, tos=dprm.type_of_shadow, occ=dprm.apply_occlusion, sdw=dprm.apply_shadow))
            # vct (8-bit only: LUT cost)
            if dt == np.uint8:
                opc = np.array([tf.get_opc(i, 255.0) for i in range(256)], np.float32)
                levels, vdims, ms = bind.vct_supervoxels(vox)
                if ms > 0.5:
                    lut = bind.vct_preintegration(opc, 255, ms)
                    vprm = bind.copy_struct(capi.default_vct_params(255.0, ms, step), bind.OrcVctParams)
                    vprm.cone_number_of_samples = int(rng.integers(5, 40)); vprm.cone_step_increase_rate = float(rng.choice([1.0, 1.1, 1.3]))
                    cmp("vct", rg.run_vct(vox, tf, levels, lut, cam, light, vprm, W, H, grad, scale), bind.vct(vox, tf, levels, vdims, lut, cam, light, vprm, W, H, scale), info)
            # gt
            nocc, nsdw = int(rng.integers(1, 5)), int(rng.integers(1, 5))
            orays, srays = capi.host_gt_ray_tables(nocc, 90.0, nsdw, 10.0)
            gprm = bind.copy_struct(capi.default_gt_params(diag, nocc, nsdw, step), bind.OrcGtParams)
            gprm.shadow_type = int(rng.integers(0, 3)); gprm.apply_occlusion = int(rng.random() < 0.8); gprm.apply_shadow = int(rng.random() < 0.8)
            glight = bind.copy_struct(capi.default_lighting(light_pos=lpos, forward=tuple(-f for f in fwd), up=(0.0, 1.0, 0.0), right=(1.0, 0.0, 0.0)), bind.OrcLighting)
            glight.apply_phong = light.apply_phong
            out, disp, stalled = rg.run_gt(vox, tf, cam, glight, gprm, orays, srays, W, H, grad, scale=scale)
            cmp("gt", out, bind.gt(vox, tf, cam, glight, gprm, orays, srays, W, H, scale), dict(info, st=gprm.shadow_type, disp=disp, stalled=stalled))
        finally:
            bind.set_gradient(None)
    assert not failures, failures



@pytest.mark.parametrize("seed", [21, 22])
def test_random_prepasses_and_light_caches_oracle_equals_reference_shaders(rg, seed):
    """Eight random scenes per seed through the pyramid shaders, the three light-cache shaders and the Sobel generator:
    ragged u8 / u16 volumes, anisotropic voxel scales, odd cache and pyramid resolutions, random eyes / lights / cone set-ups."""
    rng = np.random.default_rng(seed)
    N = 8
    failures = []
    def cmp(name, a, b, info):
        fin = np.isfinite(a) & np.isfinite(b)
        if not np.array_equal(np.isfinite(a), np.isfinite(b)) or not np.array_equal(a[fin], b[fin]):
            failures.append((name, info))
    for it in range(N):
        shape = tuple(int(v) for v in rng.integers(8, 22, 3))
        dt = np.uint16 if rng.random() < 0.3 else np.uint8
        vox = np.ascontiguousarray(synth.volume_noise(max(shape), dt)[:shape[0], :shape[1], :shape[2]]) if rng.random() < 0.5 else np.ascontiguousarray(synth.volume_gauss(max(shape), dt)[:shape[0], :shape[1], :shape[2]])
        tfname = str(rng.choice(["bonsai", "ramp", "sparse", "thin"])); tf = bind.TF(*synth.TFS[tfname])
        scale = (1.0, 1.0, 1.0) if rng.random() < 0.3 else tuple(float(v) for v in rng.choice([0.5, 1.0, 1.25, 2.0], 3))
        G = np.array([shape[2] * scale[0], shape[1] * scale[1], shape[0] * scale[2]]); diag = float(np.sqrt((G ** 2).sum()))
        res = tuple(int(v) for v in rng.integers(3, 9, 3)); pres = tuple(int(v) for v in rng.choice([4, 6, 8, 12], 3))
        sigma0 = float(rng.choice([1.0, 1.5, 0.75]))
        info = dict(it=it, shape=shape, dt=dt.__name__, tf=tfname, scale=scale, res=res, pres=pres, sigma0=sigma0)
        pyr, dims = bind.extcoef_build(vox, tf, sigma0, pres, scale)
        levels = rg.run_extcoef_pyramid(vox, tf, sigma0, pres, scale)
        for i, (a, b) in enumerate(zip(levels, rg.pyramid_levels(pyr, dims))): cmp(f"pyramid L{i}", a, b, info)
        dirv = rng.standard_normal(3); dirv /= np.linalg.norm(dirv)
        eye = tuple(dirv * diag * rng.uniform(0.6, 1.5)); center = (0.0, 0.0, 0.0); up = (0.0, 1.0, 0.0) if abs(dirv[1]) < 0.9 else (0.0, 0.0, 1.0)
        lpos = tuple(rng.standard_normal(3) * diag * rng.uniform(0.3, 2.0))
        light = bind.copy_struct(capi.default_lighting(light_pos=lpos, forward=synth.camera_forward(eye, center), up=(0.0, 1.0, 0.0), right=(1.0, 0.0, 0.0)), bind.OrcLighting)
        occ, sdw = _dos_cones(diag, (float(rng.choice([10.0, 20.0, 40.0])), int(rng.integers(0, 3)), 0.35), (float(rng.choice([0.5, 5.0])), int(rng.integers(0, 3)), 1.0))
        dprm = bind.copy_struct(capi.default_dos_params(0.5, spot_angle_deg=20.0), bind.OrcDosParams)
        dprm.apply_shadow = int(rng.random() < 0.7); dprm.apply_occlusion = int(rng.random() < 0.8); dprm.type_of_shadow = int(rng.integers(0, 3))
        _, up_v, _ = rg.camera_vectors(eye, center, up)
        cmp("dos_lc", rg.run_dos_light_cache(vox, tf, pyr, dims, eye, center, up, light, occ, sdw, dprm, res, scale),
            bind.dos_light_cache(vox.shape, pyr, dims, eye, tuple(float(v) for v in up_v), light, occ, sdw, dprm, res, scale), info)
        sat = bind.sat_build(vox, tf.ext_lut(vox.dtype.itemsize))
        eprm = bind.copy_struct(capi.default_ebs_params(diag), bind.OrcEbsParams)
        eprm.type_of_shadow = int(rng.integers(0, 2)); eprm.amb_occ_shells = int(rng.integers(1, 10)); eprm.apply_occlusion = int(rng.random() < 0.8); eprm.apply_shadow = int(rng.random() < 0.8)
        cmp("ebs_lc", rg.run_ebs_light_cache(vox, tf, sat, eye, light, eprm, res, scale), bind.ebs_light_cache(vox.shape, sat, light, eprm, res, scale), info)
        if dt == np.uint8:
            lv, vdims, ms = bind.vct_supervoxels(vox)
            if ms > 0.5:
                opc = np.array([tf.get_opc(i, 255.0) for i in range(256)], np.float32)
                lut = bind.vct_preintegration(opc, 255, ms)
                vprm = bind.copy_struct(capi.default_vct_params(255.0, ms, 0.5), bind.OrcVctParams)
                vprm.cone_number_of_samples = int(rng.integers(5, 40))
                cmp("vct_lc", rg.run_vct_light_cache(vox, tf, lv, lut, light, vprm, res, 2.0, scale), bind.vct_light_cache(vox.shape, lv, vdims, lut, light, vprm, res, scale), info)
        cmp("sobel", rg.run_sobel(vox), bind.gradient_build(vox, 3), info)
    assert not failures, failures

