"""Golden vectors produced by the reference itself (tests/golden/reference_outputs.npz, written by
tests/golden/make_golden.py where /root/reference exists): frames of the reference's own GLSL shaders run on the CPU and
outputs of its own CPU code.  These tests need neither /root/reference nor oracle/_ref.
CPU: the oracle and the host mirror must reproduce every vector exactly.  The GPU comparison of the CUDA kernels with the
golden frames lives in tests/test_zz_gpu_vs_reference_shader.py (that file sorts last, so `-x` runs everything else first)."""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest

from cpp_volume_rendering_b200 import capi, synth
from oracle import bind

HERE = os.path.dirname(os.path.abspath(__file__))
KINDS = ["rc1pass", "ebs", "dos", "vct", "gt"]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(HERE, "golden", "reference_outputs.npz"))


def _small_volume(dt=np.uint8):
    return np.ascontiguousarray(synth.volume_noise(16, dt)[:10, :12, :14])


def golden_scene():
    """The small scene behind the pre-pass / light-cache / filter vectors (also used by tests/golden/make_golden.py)."""
    n, W, H = 20, 40, 32
    vox = synth.volume_gauss(n)
    tf = bind.TF(*synth.TF_BONSAI)
    eye, center, up = synth.camera_state(0, n)
    diag = float(np.sqrt(3.0) * n)
    light = bind.copy_struct(capi.default_lighting(light_pos=synth.light_position(n), forward=synth.camera_forward(eye, center),
                                                   up=(0.0, 1.0, 0.0), right=(1.0, 0.0, 0.0)), bind.OrcLighting)
    po, ps = bind.cone_params(20.0, 1, 0.5 * diag, 0.35), bind.cone_params(0.5, 0, 0.75 * diag, 1.0)
    occ, sdw = bind.dos_cone(*bind.cone_sampler(po, 1.0), po), bind.dos_cone(*bind.cone_sampler(ps, 1.0), ps)
    dos_prm = bind.copy_struct(capi.default_dos_params(0.5, spot_angle_deg=20.0), bind.OrcDosParams)
    dos_prm.apply_shadow = 1
    pyramid_res = (8, 8, 8)
    pyr, pyr_dims = bind.extcoef_build(vox, tf, 1.0, pyramid_res)
    levels, vdims, ms = bind.vct_supervoxels(vox)
    opc = np.array([tf.get_opc(i, 255.0) for i in range(256)], np.float32)
    z, y, x = np.mgrid[0:n, 0:n, 0:n].astype(np.float64)
    r = np.sqrt((x - n / 2 + 0.5) ** 2 + (y - n / 2 + 0.5) ** 2 + (z - n / 2 + 0.5) ** 2)
    iso_prm = bind.copy_struct(capi.default_iso_params(), bind.OrcIsoParams)
    iso_prm.isovalue = 0.4
    rng = np.random.default_rng(5)
    src = rng.random((26, 34, 4)).astype(np.float32)
    src[..., :3] *= src[..., 3:4]
    return dict(n=n, W=W, H=H, vox=vox, tf=tf, eye=eye, center=center, up=up, cam=bind.camera(eye, center, up, W, H), light=light, occ=occ, sdw=sdw,
                dos_prm=dos_prm, pyramid_res=pyramid_res, pyr=pyr, pyr_dims=pyr_dims, sat=bind.sat_build(vox, tf.ext_lut(1)),
                ebs_prm=bind.copy_struct(capi.default_ebs_params(diag), bind.OrcEbsParams), vct_levels=levels, vct_dims=vdims,
                vct_lut=bind.vct_preintegration(opc, 255, ms), vct_prm=bind.copy_struct(capi.default_vct_params(255.0, ms, 0.5), bind.OrcVctParams),
                cache_res=(6, 5, 7), iso_vox=np.clip(255.0 * (1.0 - r / (0.5 * n)), 0, 255).astype(np.uint8), iso_prm=iso_prm,
                filter_src=src.astype(np.float16).astype(np.float32))


def _f16(a):
    with np.errstate(over="ignore"):
        return np.asarray(a, np.float32).astype(np.float16).astype(np.float32)


# ---------------------------------------------------------------------------------------------------------------- CPU
def test_golden_inputs_are_reproduced(gold):
    from test_zz_gpu_vs_reference_shader import Case
    assert _sha(_small_volume()) == str(gold["small_volume_sha1"])
    for kind in KINDS:
        assert _sha(Case(kind).vox) == str(gold[f"frame_{kind}_input_sha1"]), kind


@pytest.mark.parametrize("kind", KINDS)
def test_oracle_frames_equal_the_golden_reference_shader_frames(built, gold, kind):
    from test_zz_gpu_vs_reference_shader import Case
    c = Case(kind)
    L = c.orc_light()
    want = gold[f"frame_{kind}"].astype(np.float32)
    if kind == "rc1pass":
        got = bind.rc1pass(c.vox, c.tf, c.cam, c.W, c.H, c.step)
    elif kind == "ebs":
        got = bind.ebs(c.vox, c.tf, bind.sat_build(c.vox, c.tf.ext_lut(1)), c.cam, L, bind.copy_struct(c.prm, bind.OrcEbsParams), c.W, c.H)
    elif kind == "dos":
        po = bind.cone_params(c.occ_spec[0], c.occ_spec[1], 0.5 * c.diag, c.occ_spec[2])
        ps = bind.cone_params(c.sdw_spec[0], c.sdw_spec[1], 0.75 * c.diag, c.sdw_spec[2])
        occ, sdw = bind.dos_cone(*bind.cone_sampler(po, 1.0), po), bind.dos_cone(*bind.cone_sampler(ps, 1.0), ps)
        pyr, dims = bind.extcoef_build(c.vox, c.tf, 1.0, c.pyramid_res)
        got = bind.dos(c.vox, c.tf, pyr, dims, c.cam, L, occ, sdw, bind.copy_struct(c.prm, bind.OrcDosParams), c.W, c.H)
    elif kind == "vct":
        opc = capi.host_opacity_by_density(synth.TFS[c.tfname], 1)
        levels, dims, ms = bind.vct_supervoxels(c.vox)
        lut = bind.vct_preintegration(opc, 255, ms)
        prm = bind.copy_struct(capi.default_vct_params(255.0, np.float32(ms), c.step), bind.OrcVctParams)
        got = bind.vct(c.vox, c.tf, levels, dims, lut, c.cam, L, prm, c.W, c.H)
    else:
        got = bind.gt(c.vox, c.tf, c.cam, L, bind.copy_struct(c.prm, bind.OrcGtParams), c.occ_rays, c.sdw_rays, c.W, c.H)
    assert np.array_equal(got, want, equal_nan=True), float(np.nanmax(np.abs(got - want)))


def test_oracle_cpu_pieces_equal_the_golden_reference_outputs(built, gold):
    vox = _small_volume()
    tf = bind.TF(*synth.TF_BONSAI)
    # SAT: SummedAreaTable3D<double> as the EBS renderer fills and builds it
    assert np.array_equal(bind.sat_build(vox, tf.ext_lut(1)), gold["sat_bonsai"])
    # transfer function textures (GL_FLOAT client arrays)
    assert np.array_equal(tf.floats_rgbt(), gold["tf_bonsai_rgbt"]) and np.array_equal(tf.floats_rgba(), gold["tf_bonsai_rgba"])
    # VCT pre-passes, after the RG16F / R16F rounding of the upload
    levels, dims, ms = bind.vct_supervoxels(vox)
    assert np.array_equal(dims, gold["vct_dims"]) and ms == float(gold["vct_max_stddev"])
    flat = np.concatenate([l.ravel() for l in levels])
    assert np.array_equal(flat, _f16(gold["vct_levels_rg"]))
    opc = np.array([tf.get_opc(i, 255.0) for i in range(256)], np.float32)
    assert np.array_equal(bind.vct_preintegration(opc, 255, ms), _f16(gold["vct_lut"]))
    # cone section tables, ray axes, adjacent weights
    diag = float(np.sqrt(3.0) * 64)
    for name, spec, cov in (("occ", (20.0, 1, 0.35), 0.5 * diag), ("sdw", (0.5, 0, 1.0), 0.75 * diag)):
        sec, o = bind.cone_sampler(bind.cone_params(spec[0], spec[1], cov, spec[2]), 1.0)
        assert np.array_equal(sec, gold[f"cone_{name}_sections"])
        assert list(o.counts) == gold[f"cone_{name}_counts"].tolist()
        assert np.array_equal(np.array([[o.ray_axes[i][j] for j in range(3)] for i in range(10)], np.float32), gold[f"cone_{name}_axes"])
        assert np.array_equal(np.array([o.ray3_adj_weight, o.ray7_adj_weight], np.float32), gold[f"cone_{name}_adj"])
    # gradients, after the RGB16F rounding
    for mode in (1, 2):
        assert np.array_equal(bind.gradient_build(vox, mode), _f16(gold[f"gradient_mode{mode}"]))


def test_host_mirror_equals_the_golden_reference_outputs(built, gold):
    h = capi.load_host()
    # CIEDE2000 of the evaluation harness
    h.vrbh_cie2000.restype = C.c_double
    h.vrbh_cie2000.argtypes = [C.c_void_p, C.c_void_p]
    for p, want in zip(gold["cie2000_pairs"], gold["cie2000_values"]):
        a, b = np.ascontiguousarray(p[0]), np.ascontiguousarray(p[1])
        assert h.vrbh_cie2000(_p(a), _p(b)) == want
    # DDS: the stored streams decode to what the reference's decoder produced; the encoder still writes the same streams
    for version in (1, 2):
        stream = np.ascontiguousarray(gold[f"dds_v{version}_stream"])
        want = gold[f"dds_v{version}_decoded_by_reference"]
        out = np.empty(want.size + 16, np.uint8)
        n = h.vrbh_dds_decode(_p(stream), stream.size, _p(out), out.size)
        assert n == want.size and np.array_equal(out[:n], want)
        again = np.empty(stream.size, np.uint8)
        assert h.vrbh_dds_encode(_p(np.ascontiguousarray(want)), want.size, 2, 28, version, _p(again), again.size) == stream.size
        assert np.array_equal(again, stream)


def test_oracle_prepasses_light_caches_and_filters_equal_the_golden_reference_shader_outputs(built, gold):
    sc = golden_scene()
    assert _sha(sc["vox"]) == str(gold["scene_volume_sha1"])
    off = 0
    for i, (w, h, d) in enumerate((int(a), int(b), int(c)) for a, b, c in sc["pyr_dims"]):
        assert np.array_equal(sc["pyr"][off:off + w * h * d].reshape(d, h, w), gold[f"pyramid_level{i}"].astype(np.float32)), i
        off += w * h * d
    assert np.array_equal(bind.gradient_build(sc["vox"], 3), gold["sobel_mode3"].astype(np.float32))
    from oracle.refglsl import FILTER_KERNELS, camera_vectors          # pure-python helpers; the shader library is not loaded
    _, up_v, _ = camera_vectors(sc["eye"], sc["center"], sc["up"])
    dos_cache = bind.dos_light_cache(sc["vox"].shape, sc["pyr"], sc["pyr_dims"], sc["eye"], tuple(float(v) for v in up_v), sc["light"], sc["occ"], sc["sdw"],
                                     sc["dos_prm"], sc["cache_res"])
    assert np.array_equal(dos_cache, gold["light_cache_dos"].astype(np.float32))
    assert np.array_equal(bind.ebs_light_cache(sc["vox"].shape, sc["sat"], sc["light"], sc["ebs_prm"], sc["cache_res"]), gold["light_cache_ebs"].astype(np.float32),
                          equal_nan=True)
    assert np.array_equal(bind.vct_light_cache(sc["vox"].shape, sc["vct_levels"], sc["vct_dims"], sc["vct_lut"], sc["light"], sc["vct_prm"], sc["cache_res"]),
                          gold["light_cache_vct"].astype(np.float32))
    assert np.array_equal(bind.obj_march(sc["vox"], sc["tf"], sc["cam"], sc["light"].ka, sc["light"].kd, 1, 1, 0.5, dos_cache, sc["W"], sc["H"]),
                          gold["frame_obj"].astype(np.float32))
    assert np.array_equal(bind.iso(sc["iso_vox"], sc["cam"], sc["light"], sc["iso_prm"], sc["W"], sc["H"]), gold["frame_iso"].astype(np.float32))
    src = sc["filter_src"]
    assert np.array_equal(bind.frame_filter(src, src.shape[1] // 2, src.shape[0] // 2, 1), gold["filter_multisample"].astype(np.float32))
    for k, name in enumerate(FILTER_KERNELS):
        assert np.array_equal(bind.frame_filter(src, 17, 13, 2, k), gold[f"filter_down_{name}"].astype(np.float32), equal_nan=True), name
        assert np.array_equal(bind.frame_filter(src, 50, 41, 3, k), gold[f"filter_up_{name}"].astype(np.float32), equal_nan=True), name
