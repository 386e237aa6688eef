#!/usr/bin/env python
"""tests/golden/make_golden.py -- regenerates tests/golden/reference_outputs.npz.

Golden vectors produced by the REFERENCE ITSELF, run in a container that has /root/reference:
  * frames of the reference's own GLSL compute shaders executed on the CPU (oracle/_ref/librefglsl.so, oracle/glsl_cpu)
    for the inputs of tests/test_zz_gpu_vs_reference_shader.py::Case (one per renderer) -- rgba16f values, stored as fp16;
  * outputs of the reference's own CPU code compiled in place (oracle/_ref/libref.so): SummedAreaTable3D, TransferFunction1D
    textures, ConeGaussianSampler section tables, gradient generators, VCT pre-passes, CIEDE2000, the DDS decoder.
They let tests/test_golden.py pin the oracle (CPU) and the CUDA kernels (GPU) where neither /root/reference nor
oracle/_ref exists.  Inputs are regenerated from seeds by cpp_volume_rendering_b200/synth.py; their SHA-1 is stored too,
so a drift of the input generators shows up as such.

usage: python tests/golden/make_golden.py        (from the repository root)"""
import ctypes as C
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from cpp_volume_rendering_b200 import capi, synth          # noqa: E402
from oracle import bind, refglsl                            # noqa: E402


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def small_volume(dt=np.uint8):
    return np.ascontiguousarray(synth.volume_noise(16, dt)[:10, :12, :14])


def main():
    import __graft_entry__ as g
    g.build()
    from test_zz_gpu_vs_reference_shader import Case, KINDS
    assert refglsl.lib() is not None and bind.ref() is not None, "needs /root/reference (oracle/_ref)"
    out = {}
    # ---- reference shader frames
    for kind in KINDS:
        c = Case(kind)
        shader, _ = c.references(refglsl)
        with np.errstate(over="ignore"):
            h = shader.astype(np.float16)
        assert np.array_equal(h.astype(np.float32), shader, equal_nan=True)      # rgba16f: lossless
        out[f"frame_{kind}"] = h
        out[f"frame_{kind}_input_sha1"] = np.array(sha(c.vox))
    r = bind.ref()
    # ---- SummedAreaTable3D<double> through GenerateExtinctionSAT3DTex's fill (ebsrenderer.cpp:624-723)
    vox = small_volume()
    tf = bind.TF(*synth.TF_BONSAI)
    lut = tf.ext_lut(1)
    d, h, w = vox.shape
    sat = np.empty((d + 2, h + 2, w + 2), np.float32)
    r.ref_sat3d_from_volume(_p(vox), w, h, d, 1, _p(lut), _p(sat))
    out["small_volume_sha1"] = np.array(sha(vox))
    out["sat_bonsai"] = sat
    # ---- TransferFunction1D textures (GL_FLOAT client arrays of GenerateTexture_1D_RGBt / _RGBA)
    rgb, a = synth.TF_BONSAI
    rgb = np.ascontiguousarray(rgb, np.float64); a = np.ascontiguousarray(a, np.float64)
    r.ref_tf_create.restype = C.c_void_p
    rtf = r.ref_tf_create(_p(rgb), len(rgb), _p(a), len(a), 255, 0)
    rgbt = np.zeros((256, 4), np.float32); rgba = np.zeros((256, 4), np.float32)
    assert r.ref_tf_texture_rgbt(C.c_void_p(rtf), _p(rgbt), rgbt.size) == rgbt.size and r.ref_tf_texture_rgba(C.c_void_p(rtf), _p(rgba), rgba.size) == rgba.size
    out["tf_bonsai_rgbt"] = rgbt
    out["tf_bonsai_rgba"] = rgba
    # ---- VCT pre-passes
    r.ref_vct_preprocess.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_ulonglong, C.c_void_p, C.c_int,
                                     C.POINTER(C.c_double), C.c_void_p, C.c_ulonglong, C.c_void_p]
    buf = np.zeros(vox.size * 4 + 64, np.float32); dims = np.zeros((16, 3), np.int32); ms = C.c_double(0.0)
    lutb = np.zeros(256 * 64, np.float32); lut_wh = np.zeros(2, np.int32)
    n = r.ref_vct_preprocess(_p(vox), w, h, d, 1, C.c_void_p(rtf), _p(buf), buf.size, _p(dims), 16, C.byref(ms), _p(lutb), lutb.size, _p(lut_wh))
    r.ref_tf_destroy(C.c_void_p(rtf))
    assert n > 0
    total = int((dims[:n, 0].astype(np.int64) * dims[:n, 1] * dims[:n, 2] * 2).sum())
    out["vct_dims"] = dims[:n].copy()
    out["vct_levels_rg"] = buf[:total].copy()
    out["vct_max_stddev"] = np.array(ms.value)
    out["vct_lut"] = lutb[:int(lut_wh[0]) * int(lut_wh[1])].reshape(int(lut_wh[1]), int(lut_wh[0])).copy()
    # ---- ConeGaussianSampler (default occlusion and shadow cones of dosrcrenderer.cpp:47-59 for a 64^3 volume)
    diag = float(np.sqrt(3.0) * 64)
    for name, spec, cov in (("occ", (20.0, 1, 0.35), 0.5 * diag), ("sdw", (0.5, 0, 1.0), 0.75 * diag)):
        sec, o = bind.cone_sampler(bind.cone_params(spec[0], spec[1], cov, spec[2]), 1.0, use_ref=True)
        out[f"cone_{name}_sections"] = sec
        out[f"cone_{name}_counts"] = np.array(list(o.counts), np.int32)
        out[f"cone_{name}_axes"] = np.array([[o.ray_axes[i][j] for j in range(3)] for i in range(10)], np.float32)
        out[f"cone_{name}_adj"] = np.array([o.ray3_adj_weight, o.ray7_adj_weight], np.float32)
    # ---- gradient generators (utils.cpp:146-350), before the RGB16F rounding
    for mode in (1, 2):
        gr = np.empty(vox.shape + (3,), np.float32)
        assert r.ref_gradient_texture(_p(vox), w, h, d, 1, mode, _p(gr)) == 3
        out[f"gradient_mode{mode}"] = gr
    # ---- CIEDE2000 (colorutils.cpp:221-311)
    rng = np.random.default_rng(77)
    pairs = rng.integers(0, 256, (32, 2, 3)).astype(np.float64)
    r.ref_cie2000.restype = C.c_double
    r.ref_cie2000.argtypes = [C.c_void_p, C.c_void_p]
    out["cie2000_pairs"] = pairs
    out["cie2000_values"] = np.array([r.ref_cie2000(_p(np.ascontiguousarray(p[0])), _p(np.ascontiguousarray(p[1]))) for p in pairs])
    # ---- DDS v3d / v3e: streams written by the product's encoder, decoded by the REFERENCE's DDSV3
    host = capi.load_host()
    r.ref_dds_read.restype = C.c_longlong
    r.ref_dds_read.argtypes = [C.c_char_p, C.c_void_p, C.c_ulonglong]
    payload = np.concatenate([small_volume(np.uint16).astype("<u2").view(np.uint8).ravel(), rng.integers(0, 256, 777).astype(np.uint8)])
    import tempfile
    for version in (1, 2):
        need = host.vrbh_dds_encode(_p(payload), payload.size, 2, 28, version, None, 0)
        img = np.empty(need, np.uint8)
        host.vrbh_dds_encode(_p(payload), payload.size, 2, 28, version, _p(img), need)
        with tempfile.NamedTemporaryFile(suffix=".dds", delete=False) as f:
            f.write(img.tobytes())
        dec = np.empty(payload.size + 16, np.uint8)
        nd = r.ref_dds_read(f.name.encode(), _p(dec), dec.size)
        os.unlink(f.name)
        assert nd == payload.size
        out[f"dds_v{version}_stream"] = img
        out[f"dds_v{version}_decoded_by_reference"] = dec[:nd].copy()
    # ---- pre-pass / light-cache / filter shaders of the reference on a small scene (see golden_scene() in tests/test_golden.py)
    from test_golden import golden_scene
    sc = golden_scene()
    out["scene_volume_sha1"] = np.array(sha(sc["vox"]))
    for i, lev in enumerate(refglsl.run_extcoef_pyramid(sc["vox"], sc["tf"], 1.0, sc["pyramid_res"])):
        out[f"pyramid_level{i}"] = lev.astype(np.float16)
    out["sobel_mode3"] = refglsl.run_sobel(sc["vox"]).astype(np.float16)
    dos_cache = refglsl.run_dos_light_cache(sc["vox"], sc["tf"], sc["pyr"], sc["pyr_dims"], sc["eye"], sc["center"], sc["up"], sc["light"],
                                            sc["occ"], sc["sdw"], sc["dos_prm"], sc["cache_res"])
    out["light_cache_dos"] = dos_cache.astype(np.float16)
    out["light_cache_ebs"] = refglsl.run_ebs_light_cache(sc["vox"], sc["tf"], sc["sat"], sc["eye"], sc["light"], sc["ebs_prm"], sc["cache_res"]).astype(np.float16)
    out["light_cache_vct"] = refglsl.run_vct_light_cache(sc["vox"], sc["tf"], sc["vct_levels"], sc["vct_lut"], sc["light"], sc["vct_prm"], sc["cache_res"]).astype(np.float16)
    out["frame_obj"] = refglsl.run_obj(sc["vox"], sc["tf"], sc["cam"], sc["light"], 1, 1, 0.5, dos_cache, sc["W"], sc["H"]).astype(np.float16)
    out["frame_iso"] = refglsl.run_iso(sc["iso_vox"], sc["cam"], sc["light"], sc["iso_prm"], sc["W"], sc["H"]).astype(np.float16)
    src = sc["filter_src"]
    out["filter_multisample"] = refglsl.run_frame_filter(src, src.shape[1] // 2, src.shape[0] // 2, 1).astype(np.float16)
    for k, name in enumerate(refglsl.FILTER_KERNELS):
        with np.errstate(over="ignore"):
            out[f"filter_down_{name}"] = refglsl.run_frame_filter(src, 17, 13, 2, k).astype(np.float16)
            out[f"filter_up_{name}"] = refglsl.run_frame_filter(src, 50, 41, 3, k).astype(np.float16)
    path = os.path.join(HERE, "reference_outputs.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
