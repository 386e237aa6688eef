"""The CUDA kernels against the REFERENCE'S OWN GLSL SHADERS (executed on the CPU, oracle/_ref/librefglsl.so): the
closest thing to "compare against the reference's own renderer on identical volume / TF / camera inputs" (BASELINE.json)
that can exist without a GL implementation.  One case per renderer, the inputs of the first case of the corresponding
oracle parity test; tolerance is BASELINE.json's (max abs <= 2/255, PSNR >= 50 dB on float RGBA).

Every case has a CPU half (`-m "not gpu"`): the reference shader's image must equal the oracle's bit for bit, which is
what ties the two kinds of GPU parity test together.  The file sorts last so that `-x` runs everything else first."""
import numpy as np
import pytest

from cpp_volume_rendering_b200 import capi, synth
from oracle import bind, refglsl
from conftest import assert_image_parity


@pytest.fixture(scope="module")
def rg(built):
    if refglsl.lib() is None:
        pytest.skip("oracle/_ref/librefglsl.so is not available (it is built where /root/reference exists and travels with the tree)")
    return refglsl


@pytest.fixture(scope="module")
def ctx(built):
    c = capi.Context(0)
    yield c
    c.close()


class Case:
    """Inputs shared by the CPU half and the GPU half of one renderer's test."""

    def __init__(self, kind):
        self.kind = kind
        if kind == "rc1pass":
            self.n, self.W, self.H, self.step, tfname, cam_id = 64, 160, 160, 0.5, "bonsai", 0
        elif kind == "ebs":
            self.n, self.W, self.H, self.step, tfname, cam_id = 48, 128, 128, 0.5, "bonsai", 0
        elif kind == "dos":
            self.n, self.W, self.H, self.step, tfname, cam_id = 48, 112, 112, 0.5, "bonsai", 0
        elif kind == "vct":
            self.n, self.W, self.H, self.step, tfname, cam_id = 48, 112, 112, 0.5, "bonsai", 0
        elif kind == "gt":
            self.n, self.W, self.H, self.step, tfname, cam_id = 32, 64, 64, 0.5, "bonsai", 0
        self.vox = synth.volume_gauss(self.n)
        self.tfname = tfname
        self.tf = bind.TF(*synth.TFS[tfname])
        self.eye, self.center, self.up = synth.camera_state(cam_id, self.n)
        self.cam = bind.camera(self.eye, self.center, self.up, self.W, self.H)
        self.diag = float(np.sqrt(3.0) * self.n)
        fwd = synth.camera_forward(self.eye, self.center)
        lp = synth.light_position(self.n)
        if kind == "rc1pass":
            self.light = capi.default_lighting(light_pos=lp)
        elif kind == "ebs":
            self.light = capi.default_lighting(light_pos=lp, forward=fwd)
            self.prm = capi.default_ebs_params(self.diag, self.step)
        elif kind == "dos":
            self.light = capi.default_lighting(light_pos=lp, forward=fwd, up=(0.0, 1.0, 0.0), right=(1.0, 0.0, 0.0))
            self.prm = capi.default_dos_params(self.step, spot_angle_deg=20.0)
            self.pyramid_res = (32, 32, 32)
            self.occ_spec, self.sdw_spec = (20.0, 1, 0.35), (0.5, 0, 1.0)
        elif kind == "vct":
            self.light = capi.default_lighting(light_pos=lp)
        elif kind == "gt":
            self.nocc = self.nsdw = 8
            self.light = capi.default_lighting(light_pos=lp, forward=tuple(-f for f in fwd), up=(0.0, 1.0, 0.0), right=(1.0, 0.0, 0.0))
            self.prm = capi.default_gt_params(self.diag, self.nocc, self.nsdw, self.step)
            self.occ_rays, self.sdw_rays = capi.host_gt_ray_tables(self.nocc, 90.0, self.nsdw, 10.0)

    def orc_light(self):
        return bind.copy_struct(self.light, bind.OrcLighting)

    def references(self, rg, vct_max_stddev=None):
        """(reference shader image, oracle image) on the CPU."""
        L = self.orc_light()
        if self.kind == "rc1pass":
            return (rg.run_rc1pass(self.vox, self.tf, self.cam, L, self.W, self.H, self.step),
                    bind.rc1pass(self.vox, self.tf, self.cam, self.W, self.H, self.step))
        if self.kind == "ebs":
            sat = bind.sat_build(self.vox, self.tf.ext_lut(1))
            P = bind.copy_struct(self.prm, bind.OrcEbsParams)
            return rg.run_ebs(self.vox, self.tf, sat, self.cam, L, P, self.W, self.H), bind.ebs(self.vox, self.tf, sat, self.cam, L, P, self.W, self.H)
        if self.kind == "dos":
            po = bind.cone_params(self.occ_spec[0], self.occ_spec[1], 0.5 * self.diag, self.occ_spec[2])
            ps = bind.cone_params(self.sdw_spec[0], self.sdw_spec[1], 0.75 * self.diag, self.sdw_spec[2])
            so, oo = bind.cone_sampler(po, 1.0)
            ss, os_ = bind.cone_sampler(ps, 1.0)
            occ, sdw = bind.dos_cone(so, oo, po), bind.dos_cone(ss, os_, ps)
            pyr, dims = bind.extcoef_build(self.vox, self.tf, 1.0, self.pyramid_res)
            P = bind.copy_struct(self.prm, bind.OrcDosParams)
            return (rg.run_dos(self.vox, self.tf, pyr, dims, self.cam, L, occ, sdw, P, self.W, self.H),
                    bind.dos(self.vox, self.tf, pyr, dims, self.cam, L, occ, sdw, P, self.W, self.H))
        if self.kind == "vct":
            opc = capi.host_opacity_by_density(synth.TFS[self.tfname], 1)
            levels, dims, ms = bind.vct_supervoxels(self.vox)
            lut = bind.vct_preintegration(opc, 255, ms)
            prm = capi.default_vct_params(255.0, np.float32(ms) if vct_max_stddev is None else vct_max_stddev, self.step)
            P = bind.copy_struct(prm, bind.OrcVctParams)
            return (rg.run_vct(self.vox, self.tf, levels, lut, self.cam, L, P, self.W, self.H),
                    bind.vct(self.vox, self.tf, levels, dims, lut, self.cam, L, P, self.W, self.H))
        if self.kind == "gt":
            P = bind.copy_struct(self.prm, bind.OrcGtParams)
            out, dispatches, stalled = rg.run_gt(self.vox, self.tf, self.cam, L, P, self.occ_rays, self.sdw_rays, self.W, self.H)
            assert dispatches > 10 and stalled <= 0.02 * self.W * self.H
            return out, bind.gt(self.vox, self.tf, self.cam, L, P, self.occ_rays, self.sdw_rays, self.W, self.H)
        raise ValueError(self.kind)


KINDS = ["rc1pass", "ebs", "dos", "vct", "gt"]


@pytest.mark.parametrize("kind", KINDS)
def test_reference_shader_equals_oracle_on_the_gpu_parity_inputs(rg, kind):
    shader, oracle = Case(kind).references(rg)
    assert np.array_equal(np.isfinite(shader), np.isfinite(oracle))
    fin = np.isfinite(oracle)
    assert np.array_equal(shader[fin], oracle[fin]), float(np.abs(shader[fin] - oracle[fin]).max())
    assert (oracle[..., 3] > 0).sum() > 100


def _render_case(ctx, c):
    """The CUDA frame of a Case through the C ABI; also returns the GPU's maximum deviation for VCT (None otherwise)."""
    kind = c.kind
    cam = capi.make_camera(c.eye, c.center, c.up, c.W, c.H)
    ctx.volume_upload(c.vox)
    ctx.tf_upload(c.tf.floats_rgbt(), c.tf.floats_rgba())
    ms = None
    if kind == "rc1pass":
        ctx.frame_resize(c.W, c.H)
        ctx.rc1pass_render(cam, c.step, count_samples=True)
    elif kind == "ebs":
        ctx.sat_build(c.tf.ext_lut(1))
        ctx.frame_resize(c.W, c.H)
        ctx.ebs_render(cam, c.light, c.prm)
    elif kind == "dos":
        ho, _, _ = capi.host_cone_sampler(c.occ_spec[0], c.occ_spec[1], 0.5 * c.diag, c.occ_spec[2])
        hs, _, _ = capi.host_cone_sampler(c.sdw_spec[0], c.sdw_spec[1], 0.75 * c.diag, c.sdw_spec[2])
        ctx.extcoef_build(1.0, c.pyramid_res)
        ctx.dos_set_cones(ho, hs)
        ctx.frame_resize(c.W, c.H)
        ctx.dos_render(cam, c.light, c.prm)
    elif kind == "vct":
        ctx.vct_build(capi.host_opacity_by_density(synth.TFS[c.tfname], 1))
        ctx.frame_resize(c.W, c.H)
        _, _, ms = ctx.vct_info()
        ctx.vct_render(cam, c.light, capi.default_vct_params(255.0, ms, c.step))
    elif kind == "gt":
        ctx.frame_resize(c.W, c.H)
        ctx.gt_set_rays(c.occ_rays, c.sdw_rays)
        ctx.gt_render(cam, c.light, c.prm)
    return ctx.frame_read().copy(), ms


@pytest.mark.gpu
@pytest.mark.parametrize("kind", KINDS)
def test_cuda_kernel_matches_reference_shader(ctx, rg, kind):
    c = Case(kind)
    img, ms = _render_case(ctx, c)
    shader, _ = c.references(rg, vct_max_stddev=ms)
    assert_image_parity(img, shader, what=f"{kind}: CUDA vs the reference's own shader")


@pytest.mark.gpu
def test_object_space_march_phong_branch_matches_oracle_and_reference_shader(ctx, rg):
    """k_obj_march<PHONG> (obj_ray_marching.comp:236-257): a DOS light cache is built on the GPU and read back, then the
    march over exactly that cache is compared with the oracle and with the reference's shader."""
    n, W, H, step = 36, 88, 80, 0.5
    vox = synth.volume_gauss(n)
    tf = bind.TF(*synth.TF_BONSAI)
    eye, center, up = synth.camera_state(0, n)
    diag = float(np.sqrt(3.0) * n)
    f = -(np.asarray(center, np.float32) - np.asarray(eye, np.float32)); f /= np.sqrt(np.sum(f * f, dtype=np.float32))
    r = np.cross(np.asarray(up, np.float32), f).astype(np.float32); r /= np.sqrt(np.sum(r * r, dtype=np.float32))
    eye_up = np.cross(f, r).astype(np.float32); eye_up /= np.sqrt(np.sum(eye_up * eye_up, dtype=np.float32))
    ho, _, _ = capi.host_cone_sampler(20.0, 1, 0.5 * diag, 0.35)
    hs, _, _ = capi.host_cone_sampler(0.5, 0, 0.75 * diag, 1.0)
    light = capi.default_lighting(light_pos=synth.light_position(n), forward=synth.camera_forward(eye, center), up=(0.0, 1.0, 0.0), right=(1.0, 0.0, 0.0))
    prm = capi.default_dos_params(step, spot_angle_deg=20.0)
    prm.apply_shadow = 1
    ctx.volume_upload(vox)
    ctx.tf_upload(tf.floats_rgbt(), tf.floats_rgba())
    ctx.extcoef_build(1.0, (32, 32, 32))
    ctx.dos_set_cones(ho, hs)
    ctx.frame_resize(W, H)
    ctx.dos_light_cache_build(eye, tuple(float(v) for v in eye_up), light, prm, (12, 12, 12))
    cache = ctx.light_cache_read()
    cam = capi.make_camera(eye, center, up, W, H)
    light.apply_phong = 1
    with pytest.raises(capi.VrbError, match="gradient texture"):
        ctx.obj_march_render(cam, light, step, 1, 1)
    ctx.gradient_build(capi.GRADIENT_SOBEL_FELDMAN)
    ctx.obj_march_render(cam, light, step, 1, 1, count_samples=True)
    img = ctx.frame_read().copy()
    nsamp = ctx.last_sample_count
    light.apply_phong = 0
    ctx.obj_march_render(cam, light, step, 1, 1)
    plain = ctx.frame_read().copy()
    light.apply_phong = 1
    L = bind.copy_struct(light, bind.OrcLighting)
    ocam = bind.camera(eye, center, up, W, H)
    grad = bind.gradient_build(vox, 1)
    bind.set_gradient(grad)
    try:
        ref, ns = bind.obj_march_lit(vox, tf, ocam, L, 1, 1, step, cache, W, H, count=True)
    finally:
        bind.set_gradient(None)
    shader = rg.run_obj(vox, tf, ocam, L, 1, 1, step, cache, W, H, grad)
    assert np.array_equal(shader, ref)
    assert np.abs(ref[..., :3] - plain[..., :3]).max() > 0.02, "the Phong branch must change the image for the test to mean anything"
    assert_image_parity(img, shader, what="object-space march, Phong branch")
    assert abs(nsamp - int(ns.sum())) <= max(2, int(ns.sum()) // 100000)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", KINDS)
def test_cuda_kernel_matches_golden_reference_shader_frame(ctx, kind):
    """The same comparison against the committed golden frames (tests/golden/reference_outputs.npz): needs no oracle/_ref."""
    import os
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_outputs.npz"))
    c = Case(kind)
    img, _ = _render_case(ctx, c)
    assert_image_parity(img, gold[f"frame_{kind}"].astype(np.float32), what=f"{kind}: CUDA vs the golden frame of the reference's shader")


@pytest.mark.gpu
def test_gt_bounding_box_placeholder_matches_oracle_and_reference_shader(ctx, rg):
    """k_gt_cube (RedrawCube of rc1pcrtgt, vol_intersection.comp) against the oracle and the reference's shader; the ray
    set-up is the marchers', bit-identical by construction, so the face picked per pixel must be the same."""
    for cam_id, (W, H), shape, scale in ((0, (96, 80), (20, 20, 20), (1.0, 1.0, 1.0)), (3, (70, 90), (12, 20, 30), (1.0, 0.5, 2.0))):
        vox = np.ascontiguousarray(synth.volume_noise(max(shape))[:shape[0], :shape[1], :shape[2]])
        eye, center, up = synth.camera_state(cam_id, max(shape))
        ctx.volume_upload(vox, scale)
        ctx.frame_resize(W, H)
        ctx.gt_cube_render(capi.make_camera(eye, center, up, W, H))
        img = ctx.frame_read().copy()
        ocam = bind.camera(eye, center, up, W, H)
        ref = bind.gt_cube(shape, ocam, W, H, scale)
        assert np.array_equal(rg.run_gt_cube(shape, ocam, W, H, scale), ref)
        assert (ref[..., 3] > 0).sum() > 100
        assert (np.abs(img - ref).max(-1) > 0).mean() <= 0.001, int((np.abs(img - ref).max(-1) > 0).sum())


@pytest.mark.gpu
def test_same_size_extinction_pyramid_with_voxel_scales_matches_oracle(ctx):
    """vrb_extcoef_build(.., 0, 0, 0) = GenerateExtinctionCoefficientVolumeSameSize: the base level's taps are placed with the
    volume's own voxel size (gen_extcoefvol_samesize.comp:45).  Same tolerance as tests/test_dos.py's pyramid test (every level
    is stored fp16 and filtered from the fp16 level above)."""
    for shape, scale in (((20, 24, 28), (1.0, 1.0, 1.0)), ((18, 22, 26), (1.7525913, 2.785067, 0.45124617))):
        vox = np.ascontiguousarray(synth.volume_noise(max(shape))[:shape[0], :shape[1], :shape[2]])
        tf = bind.TF(*synth.TF_BONSAI)
        ctx.volume_upload(vox, scale)
        ctx.tf_upload(tf.floats_rgbt(), tf.floats_rgba())
        ctx.extcoef_build(1.0, None)
        levels = ctx.extcoef_levels()
        pyr, dims = bind.extcoef_build(vox, tf, 1.0, None, scale)
        assert len(levels) == len(dims)
        off = 0
        for l, lev in enumerate(levels):
            w, h, d = (int(v) for v in dims[l])
            want = pyr[off:off + w * h * d].reshape(d, h, w)
            off += w * h * d
            assert lev.shape == want.shape
            tol = (2 + l) * 2.0 ** -10
            assert np.all(np.abs(lev - want) <= tol * np.maximum(np.abs(want), 2.0 ** -10)), (l, float(np.abs(lev - want).max()))
            if lev.size >= 64:                       # a one-texel top level is either 0 % or 100 % equal
                assert np.mean(lev == want) > 0.9
