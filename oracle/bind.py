"""oracle/bind.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes bindings of oracle/liboracle.so (CPU restatement) and oracle/_ref/libref.so (the reference's own CPU sources
compiled in place).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module; nothing under cpp_volume_rendering_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_orc = None
_ref = None


class OrcCamera(C.Structure):
    _fields_ = [("eye", C.c_float * 3), ("lookat", C.c_float * 16), ("tan_fovy", C.c_float), ("aspect", C.c_float)]


class OrcLighting(C.Structure):
    _fields_ = [("ka", C.c_float), ("kd", C.c_float), ("ks", C.c_float), ("shininess", C.c_float),
                ("ispecular", C.c_float * 3), ("light_pos", C.c_float * 3), ("light_forward", C.c_float * 3),
                ("light_up", C.c_float * 3), ("light_right", C.c_float * 3), ("spot_angle_deg", C.c_float),
                ("apply_phong", C.c_int)]


class OrcEbsParams(C.Structure):
    _fields_ = [("step_size", C.c_float), ("apply_occlusion", C.c_int), ("apply_shadow", C.c_int),
                ("amb_occ_shells", C.c_int), ("amb_occ_radius", C.c_float), ("sdw_cone_angle_rad", C.c_float),
                ("sdw_sample_interval", C.c_float), ("sdw_initial_step", C.c_float), ("sdw_ui_weight", C.c_float),
                ("sdw_cone_max_distance", C.c_float), ("type_of_shadow", C.c_int), ("count_samples", C.c_int)]


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    if force or not os.path.exists(so):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    if os.path.isdir("/root/reference") and (force or not os.path.exists(os.path.join(_HERE, "_ref", "libref.so"))):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref"] + (["-B"] if force else []))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def orc():
    global _orc
    if _orc is None:
        build()
        _orc = C.CDLL(os.path.join(_HERE, "liboracle.so"))
        _orc.orc_tf_get_extn.restype = C.c_float
        _orc.orc_tf_get_extn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double]
        _orc.orc_tf_get_opcn.restype = C.c_float
        _orc.orc_tf_get_opcn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double]
        _orc.orc_tf_get_opc.restype = C.c_float
        _orc.orc_tf_get_opc.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double]
        _orc.orc_tf_get.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_void_p]
        _orc.orc_f32_to_f16_bits.restype = C.c_uint16
        _orc.orc_f32_to_f16_bits.argtypes = [C.c_float]
        _orc.orc_f16_bits_to_f32.restype = C.c_float
        _orc.orc_f16_bits_to_f32.argtypes = [C.c_uint16]
        _orc.orc_round_f16_array.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        _orc.orc_volume_to_r16f.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
        _orc.orc_rc1pass_render.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                            C.POINTER(OrcCamera), C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        _orc.orc_sat_build.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _orc.orc_sat_build_u64.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        _orc.orc_ebs_render.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                        C.POINTER(OrcCamera), C.POINTER(OrcLighting), C.POINTER(OrcEbsParams),
                                        C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    return _orc


def ref():
    """The reference's own CPU code (oracle/_ref/libref.so); None when it was never built and cannot be built here."""
    global _ref
    if _ref is None:
        so = os.path.join(_HERE, "_ref", "libref.so")
        if not os.path.exists(so):
            if not os.path.isdir("/root/reference"):
                return None
            build()
        _ref = C.CDLL(so)
        _ref.ref_tf_create.restype = C.c_void_p
        _ref.ref_tf_create.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
        _ref.ref_tf_destroy.argtypes = [C.c_void_p]
        _ref.ref_tf_get.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_void_p]
        for n in ("ref_tf_get_extn", "ref_tf_get_opcn"):
            getattr(_ref, n).restype = C.c_float
            getattr(_ref, n).argtypes = [C.c_void_p, C.c_double]
        _ref.ref_tf_get_opc.restype = C.c_float
        _ref.ref_tf_get_opc.argtypes = [C.c_void_p, C.c_double, C.c_double]
        _ref.ref_tf_texture_rgbt.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        _ref.ref_tf_texture_rgba.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        _ref.ref_sat3d_double.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        _ref.ref_sat3d_u64.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        _ref.ref_sat3d_from_volume.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        _ref.ref_volume_normalized_sample.restype = C.c_double
        _ref.ref_volume_normalized_sample.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    return _ref


# ---- convenience wrappers ------------------------------------------------------------------------------------------
class TF:
    """Oracle transfer function built from .tf1d control points."""

    def __init__(self, rgb_pts, a_pts, max_density=255, ext_type=0):
        self.rgb = np.ascontiguousarray(rgb_pts, np.float64)
        self.a = np.ascontiguousarray(a_pts, np.float64)
        self.max_density = int(max_density)
        self.ext_type = int(ext_type)
        self.n = self.max_density + 1
        self.table = np.zeros((self.n, 4), np.float64)
        orc().orc_tf_build(_p(self.rgb), len(self.rgb), _p(self.a), len(self.a), self.max_density, _p(self.table))

    def _tex(self, fn):
        out = np.zeros((self.n, 4), np.float32)
        getattr(orc(), fn)(_p(self.table), self.max_density, self.ext_type, _p(out))
        return out

    def texture_rgbt(self):   # fp16-rounded texels
        return self._tex("orc_tf_texture_rgbt")

    def texture_rgba(self):
        return self._tex("orc_tf_texture_rgba")

    def floats_rgbt(self):    # GL_FLOAT client array (what the C ABI takes)
        return self._tex("orc_tf_floats_rgbt")

    def floats_rgba(self):
        return self._tex("orc_tf_floats_rgba")

    def get_extn(self, x):
        return orc().orc_tf_get_extn(_p(self.table), self.max_density, self.ext_type, float(x))

    def get_opcn(self, x):
        return orc().orc_tf_get_opcn(_p(self.table), self.max_density, self.ext_type, float(x))

    def get_opc(self, v, mx):
        return orc().orc_tf_get_opc(_p(self.table), self.max_density, self.ext_type, float(v), float(mx))

    def ext_lut(self, bpv):
        """GetExtN(v / max) for every voxel value (ebsrenderer.cpp:655)."""
        n = 256 if bpv == 1 else 65536
        mx = 255.0 if bpv == 1 else 65535.0
        return np.array([self.get_extn(v / mx) for v in range(n)], np.float32)


def camera(eye, center, up, width, height, fovy_deg=45.0):
    cam = OrcCamera()
    e = np.asarray(eye, np.float32); c = np.asarray(center, np.float32); u = np.asarray(up, np.float32)
    m = np.zeros(16, np.float32)
    orc().orc_look_at(_p(e), _p(c), _p(u), _p(m))
    cam.eye[:] = e.tolist()
    cam.lookat[:] = m.tolist()
    cam.tan_fovy = np.float32(np.tan(np.float64(np.float32(fovy_deg)) * (np.pi / 180.0) / 2.0))
    cam.aspect = np.float32(np.float32(width) / np.float32(height))
    return cam


def volume_r16f(vox):
    vox = np.ascontiguousarray(vox)
    out = np.empty(vox.shape, np.float32)
    orc().orc_volume_to_r16f(_p(vox), vox.size, vox.dtype.itemsize, _p(out))
    return out


def rc1pass(vox, tf, cam, W, H, step=0.5, scale=(1.0, 1.0, 1.0), count=False):
    tex = volume_r16f(vox)
    d, h, w = vox.shape
    G = np.array([w * scale[0], h * scale[1], d * scale[2]], np.float32)
    rgbt = tf.texture_rgbt()
    out = np.zeros((H, W, 4), np.float32)
    ns = np.zeros((H, W), np.uint32) if count else None
    orc().orc_rc1pass_render(_p(tex), w, h, d, _p(G), _p(rgbt), tf.n, C.byref(cam), C.c_float(step), W, H, _p(out),
                             _p(ns) if count else None)
    return (out, ns) if count else out


def rc1pass_lit(vox, tf, cam, light, W, H, step=0.5, scale=(1.0, 1.0, 1.0), count=False):
    """rc1pass with the ShadeBlinnPhong branch (light.apply_phong = 1 needs set_gradient first)."""
    tex = volume_r16f(vox)
    d, h, w = vox.shape
    G = np.array([w * scale[0], h * scale[1], d * scale[2]], np.float32)
    rgbt = tf.texture_rgbt()
    out = np.zeros((H, W, 4), np.float32)
    ns = np.zeros((H, W), np.uint32) if count else None
    rc = orc().orc_rc1pass_render_lit(_p(tex), w, h, d, _p(G), _p(rgbt), tf.n, C.byref(cam), C.c_float(step), W, H, _p(out),
                                      _p(ns) if count else None, C.byref(light))
    assert rc == 0, rc
    return (out, ns) if count else out


class OrcIsoParams(C.Structure):
    _fields_ = [("isovalue", C.c_float), ("step_size_small", C.c_float), ("step_size_large", C.c_float), ("step_size_range", C.c_float),
                ("color", C.c_float * 4), ("count_samples", C.c_int)]


def iso(vox, cam, light, params, W, H, scale=(1.0, 1.0, 1.0), count=False):
    """rc1pisoadapt (adaptive-step isosurface ray caster); light.apply_phong = 1 needs set_gradient first."""
    tex = volume_r16f(vox)
    d, h, w = vox.shape
    G = np.array([w * scale[0], h * scale[1], d * scale[2]], np.float32)
    out = np.zeros((H, W, 4), np.float32)
    ns = np.zeros((H, W), np.uint32) if count else None
    rc = orc().orc_iso_render(_p(tex), w, h, d, _p(G), C.byref(cam), C.byref(light), C.byref(params), W, H, _p(out), _p(ns) if count else None)
    assert rc == 0, rc
    return (out, ns) if count else out


GRADIENT_SOBEL_FELDMAN, GRADIENT_FINITE_DIFFERENCES, GRADIENT_COMPUTE_SHADER_SOBEL = 1, 2, 3
_bound_gradient = None


def gradient_build(vox, mode, use_ref=False):
    """RGB16F gradient texels (d,h,w,3 float32, fp16-rounded) of DataManager::GenerateStructuredGradientTexture.
    use_ref: the reference's own libs/volvis_utils/utils.cpp (modes 1 and 2 only), rounded to fp16 here."""
    vox = np.ascontiguousarray(vox)
    d, h, w = vox.shape
    out = np.empty((d, h, w, 3), np.float32)
    if use_ref:
        r = ref()
        assert mode in (1, 2)
        ch = r.ref_gradient_texture(_p(vox), w, h, d, vox.dtype.itemsize, mode, _p(out))
        assert ch == 3, ch
        with np.errstate(over="ignore"):
            return out.astype(np.float16).astype(np.float32)
    rc = orc().orc_gradient_build(_p(vox), w, h, d, vox.dtype.itemsize, mode, _p(out))
    assert rc == 0, rc
    return out


def set_gradient(grad):
    """Bind (or with None unbind) TexVolumeGradient for the oracle's renderers."""
    global _bound_gradient
    if grad is None:
        orc().orc_set_gradient(None, 0, 0, 0)
        _bound_gradient = None
        return
    g = np.ascontiguousarray(grad, np.float32)
    _bound_gradient = g                                   # keep the array alive
    orc().orc_set_gradient(_p(g), g.shape[2], g.shape[1], g.shape[0])


def frame_filter(src, out_w, out_h, pass_id, kernel=1):
    """Pixel multi-scaling pass of RenderFrameToScreen: pass 1 multisample, 2 down-scale, 3 up-scale; kernel =
    vis::IMAGE_FILTER_KERNEL.  src (H, W, 4) float32 of fp16 values; returns (out_h, out_w, 4).  src is not modified."""
    s = np.ascontiguousarray(src, np.float32).copy()
    out = np.zeros((out_h, out_w, 4), np.float32)
    rc = orc().orc_frame_filter(_p(s), s.shape[1], s.shape[0], _p(out), out_w, out_h, int(pass_id), int(kernel))
    assert rc == 0, rc
    return out


def sat_build(vox, ext_lut, want_f64=False):
    vox = np.ascontiguousarray(vox)
    d, h, w = vox.shape
    lut = np.ascontiguousarray(ext_lut, np.float32)
    out = np.empty((d + 2, h + 2, w + 2), np.float32)
    o64 = np.empty((d + 2, h + 2, w + 2), np.float64) if want_f64 else None
    orc().orc_sat_build(_p(vox), w, h, d, vox.dtype.itemsize, _p(lut), _p(out), _p(o64) if want_f64 else None)
    return (out, o64) if want_f64 else out


def sat_build_u64(vox, lut_u32):
    vox = np.ascontiguousarray(vox)
    d, h, w = vox.shape
    lut = np.ascontiguousarray(lut_u32, np.uint32)
    out = np.empty((d, h, w), np.uint64)
    orc().orc_sat_build_u64(_p(vox), w, h, d, vox.dtype.itemsize, _p(lut), _p(out))
    return out


def ebs(vox, tf, sat, cam, light, params, W, H, scale=(1.0, 1.0, 1.0), count=False):
    tex = volume_r16f(vox)
    d, h, w = vox.shape
    sc = np.array(scale, np.float32)
    rgbt = tf.texture_rgbt()
    sat = np.ascontiguousarray(sat, np.float32)
    out = np.zeros((H, W, 4), np.float32)
    ns = np.zeros((H, W), np.uint32) if count else None
    orc().orc_ebs_render(_p(tex), w, h, d, _p(sc), _p(sat), _p(rgbt), tf.n, C.byref(cam), C.byref(light), C.byref(params),
                         W, H, _p(out), _p(ns) if count else None)
    return (out, ns) if count else out


class ConeSamplerParams(C.Structure):
    _fields_ = [("cone_half_angle", C.c_float), ("initial_step", C.c_float), ("max_packing", C.c_int),
                ("covered_distance", C.c_float), ("d_sigma", C.c_float), ("r_sigma", C.c_float), ("ui_weight", C.c_float)]


class ConeSamplerOut(C.Structure):
    _fields_ = [("n_sections", C.c_int), ("counts", C.c_int * 3), ("ray_axes", (C.c_float * 3) * 10),
                ("ray3_adj_weight", C.c_float), ("ray7_adj_weight", C.c_float)]


class OrcDosCone(C.Structure):
    _fields_ = [("sections", C.c_void_p), ("n_sections", C.c_int), ("counts", C.c_int * 3), ("initial_step", C.c_float),
                ("ray7_adj_weight", C.c_float), ("ui_weight", C.c_float), ("axes", (C.c_float * 3) * 10)]


class OrcDosParams(C.Structure):
    _fields_ = [("step_size", C.c_float), ("apply_occlusion", C.c_int), ("apply_shadow", C.c_int), ("type_of_shadow", C.c_int),
                ("spot_cos", C.c_float), ("count_samples", C.c_int)]


def cone_params(half_angle, max_packing, covered_distance, ui_weight, initial_step=3.0, d_sigma=1.25, r_sigma=2.0):
    """ConeGaussianSampler set-up as RC1PConeTracingDirOcclusionShading does it (dosrcrenderer.cpp:42-59,111-113);
    the setters clamp like the reference's (conegaussiansampler.cpp:57-60,86-89,147-150)."""
    return ConeSamplerParams(min(max(half_angle, 0.5), 89.5), max(initial_step, 0.0), max_packing,
                             max(covered_distance, 10.0), min(max(d_sigma, 1.0), 3.0), min(max(r_sigma, 0.5), 3.0), ui_weight)


def cone_sampler(params, min_sigma=1.0, use_ref=False):
    """Returns (sections float32 [n,4] (GL_FLOAT client array), ConeSamplerOut)."""
    lib = ref() if use_ref else orc()
    fn = lib.ref_cone_sampler_compute if use_ref else lib.orc_cone_sampler_compute
    fn.argtypes = [C.POINTER(ConeSamplerParams), C.c_double, C.c_void_p, C.c_int, C.POINTER(ConeSamplerOut)]
    buf = np.zeros((4096, 4), np.float32)
    out = ConeSamplerOut()
    rc = fn(C.byref(params), float(min_sigma), _p(buf), 4096, C.byref(out))
    assert rc == 0, rc
    return buf[:out.n_sections].copy(), out


def dos_cone(sections, out, params):
    """Uniform block of one sampler for orc_dos_render; sections are rounded to RGBA16F (texelFetch sees fp16 texels)."""
    with np.errstate(over="ignore"):
        sec16 = np.ascontiguousarray(sections.astype(np.float16).astype(np.float32))
    c = OrcDosCone()
    c._keep = sec16
    c.sections = sec16.ctypes.data
    c.n_sections = out.n_sections
    c.counts[:] = list(out.counts)
    c.initial_step = params.initial_step
    c.ray7_adj_weight = out.ray7_adj_weight
    c.ui_weight = params.ui_weight
    for i in range(10):
        for j in range(3):
            c.axes[i][j] = out.ray_axes[i][j]
    return c


def extcoef_build(vox, tf, sigma0=1.0, res=(128, 128, 128), scale=(1.0, 1.0, 1.0)):
    """Extinction-coefficient pyramid: (concatenated fp16-rounded levels, dims [n_levels,3]).  res=None: the same-size build."""
    o = orc()
    o.orc_extcoef_build_ex.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float,
                                       C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
    tex = volume_r16f(vox)
    d, h, w = vox.shape
    G = np.array([w * scale[0], h * scale[1], d * scale[2]], np.float32)
    rgba = tf.texture_rgba()
    same_size = res is None                      # GenerateExtinctionCoefficientVolumeSameSize: base resolution = the volume's
    rw, rh, rd = (w, h, d) if same_size else res
    nlev = o.orc_extcoef_levels(rw, rh, rd)
    dims = np.zeros((nlev, 3), np.int32)
    cap = int(rw * rh * rd * 1.2) + 64
    buf = np.zeros(cap, np.float32)
    sc = np.array(scale, np.float32)
    n = o.orc_extcoef_build_ex(_p(tex), w, h, d, _p(G), _p(sc), _p(rgba), tf.n, C.c_float(sigma0), *((0, 0, 0) if same_size else (rw, rh, rd)),
                               _p(buf), cap, _p(dims))
    assert n == nlev, n
    total = int((dims[:, 0].astype(np.int64) * dims[:, 1] * dims[:, 2]).sum())
    return buf[:total].copy(), dims


def dos(vox, tf, pyramid, dims, cam, light, occ, sdw, params, W, H, scale=(1.0, 1.0, 1.0), count=False):
    o = orc()
    o.orc_dos_render.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                 C.POINTER(OrcCamera), C.POINTER(OrcLighting), C.POINTER(OrcDosCone), C.POINTER(OrcDosCone),
                                 C.POINTER(OrcDosParams), C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    tex = volume_r16f(vox)
    d, h, w = vox.shape
    sc = np.array(scale, np.float32)
    rgbt = tf.texture_rgbt()
    pyramid = np.ascontiguousarray(pyramid, np.float32)
    dims = np.ascontiguousarray(dims, np.int32)
    out = np.zeros((H, W, 4), np.float32)
    ns = np.zeros((H, W), np.uint32) if count else None
    o.orc_dos_render(_p(tex), w, h, d, _p(sc), _p(pyramid), _p(dims), len(dims), _p(rgbt), tf.n, C.byref(cam), C.byref(light),
                     C.byref(occ), C.byref(sdw), C.byref(params), W, H, _p(out), _p(ns) if count else None)
    return (out, ns) if count else out


def dos_light_cache(vox_shape, pyramid, dims, eye, eye_up, light, occ, sdw, params, res=(32, 32, 32), scale=(1.0, 1.0, 1.0)):
    """K6 rc1pdosct/lightcachecomputation.comp: [rd, rh, rw, 2] fp16-rounded (Iocc, Ishadow)."""
    o = orc()
    o.orc_dos_light_cache.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                      C.POINTER(OrcLighting), C.POINTER(OrcDosCone), C.POINTER(OrcDosCone), C.POINTER(OrcDosParams),
                                      C.c_int, C.c_int, C.c_int, C.c_void_p]
    d, h, w = vox_shape
    sc = np.array(scale, np.float32)
    pyramid = np.ascontiguousarray(pyramid, np.float32)
    dims = np.ascontiguousarray(dims, np.int32)
    e = np.array(eye, np.float32); u = np.array(eye_up, np.float32)
    out = np.zeros((res[2], res[1], res[0], 2), np.float32)
    o.orc_dos_light_cache(w, h, d, _p(sc), _p(pyramid), _p(dims), len(dims), _p(e), _p(u), C.byref(light), C.byref(occ), C.byref(sdw),
                          C.byref(params), int(res[0]), int(res[1]), int(res[2]), _p(out))
    return out


def obj_march(vox, tf, cam, ka, kd, apply_occlusion, apply_shadow, step, cache, W, H, scale=(1.0, 1.0, 1.0), count=False):
    """K7 _common_shaders/obj_ray_marching.comp (:210-333) over a light cache [rd, rh, rw, 2]."""
    o = orc()
    o.orc_obj_march.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(OrcCamera), C.c_float,
                                C.c_float, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.c_void_p, C.c_void_p]
    tex = volume_r16f(vox)
    d, h, w = vox.shape
    sc = np.array(scale, np.float32)
    rgbt = tf.texture_rgbt()
    cache = np.ascontiguousarray(cache, np.float32)
    rd, rh, rw, _ = cache.shape
    out = np.zeros((H, W, 4), np.float32)
    ns = np.zeros((H, W), np.uint32) if count else None
    o.orc_obj_march(_p(tex), w, h, d, _p(sc), _p(rgbt), tf.n, C.byref(cam), float(ka), float(kd), int(apply_occlusion), int(apply_shadow),
                    float(step), _p(cache), rw, rh, rd, W, H, _p(out), _p(ns) if count else None)
    return (out, ns) if count else out


def obj_march_lit(vox, tf, cam, light, apply_occlusion, apply_shadow, step, cache, W, H, scale=(1.0, 1.0, 1.0), count=False):
    """obj_march with the gradient Blinn-Phong branch (light.apply_phong = 1 needs set_gradient first)."""
    o = orc()
    o.orc_obj_march_lit.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(OrcCamera), C.POINTER(OrcLighting),
                                    C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    tex = volume_r16f(vox)
    d, h, w = vox.shape
    sc = np.array(scale, np.float32)
    rgbt = tf.texture_rgbt()
    cache = np.ascontiguousarray(cache, np.float32)
    rd, rh, rw, _ = cache.shape
    out = np.zeros((H, W, 4), np.float32)
    ns = np.zeros((H, W), np.uint32) if count else None
    rc = o.orc_obj_march_lit(_p(tex), w, h, d, _p(sc), _p(rgbt), tf.n, C.byref(cam), C.byref(light), int(apply_occlusion), int(apply_shadow),
                             float(step), _p(cache), rw, rh, rd, W, H, _p(out), _p(ns) if count else None)
    assert rc == 0, rc
    return (out, ns) if count else out


class OrcGtParams(C.Structure):
    _fields_ = [("step_size", C.c_float), ("light_ray_initial_gap", C.c_float), ("light_ray_step_size", C.c_float),
                ("apply_occlusion", C.c_int), ("occ_num_rays", C.c_int), ("occ_cone_distance", C.c_float),
                ("apply_shadow", C.c_int), ("sdw_num_rays", C.c_int), ("sdw_cone_distance", C.c_float),
                ("shadow_type", C.c_int), ("count_samples", C.c_int)]


class OrcVctParams(C.Structure):
    _fields_ = [("step_size", C.c_float), ("apply_occlusion", C.c_int), ("apply_shadow", C.c_int),
                ("tan_cone_apex_angle", C.c_float), ("cone_step_size", C.c_float), ("cone_step_increase_rate", C.c_float),
                ("cone_initial_step", C.c_float), ("opacity_correction_factor", C.c_float), ("apply_opacity_correction", C.c_int),
                ("cone_number_of_samples", C.c_int), ("volume_max_density", C.c_float), ("volume_max_stddev", C.c_float),
                ("count_samples", C.c_int)]


def gt(vox, tf, cam, light, params, occ_rays, sdw_rays, W, H, scale=(1.0, 1.0, 1.0), count=False):
    """occ_rays / sdw_rays: float tables BEFORE the RGB16F rounding (rounded here like the texture upload)."""
    o = orc()
    o.orc_gt_render.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(OrcCamera),
                                C.POINTER(OrcLighting), C.POINTER(OrcGtParams), C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                C.c_void_p, C.c_void_p, C.c_void_p]
    tex = volume_r16f(vox)
    d, h, w = vox.shape
    G = np.array([w * scale[0], h * scale[1], d * scale[2]], np.float32)
    rgbt = tf.texture_rgbt()
    r16 = lambda a: np.ascontiguousarray(np.asarray(a, np.float32).reshape(-1, 3).astype(np.float16).astype(np.float32))
    oc, sd = r16(occ_rays) if len(occ_rays) else np.zeros((1, 3), np.float32), r16(sdw_rays) if len(sdw_rays) else np.zeros((1, 3), np.float32)
    out = np.zeros((H, W, 4), np.float32)
    ns = np.zeros((H, W), np.uint32)
    nsec = C.c_uint64(0)
    o.orc_gt_render(_p(tex), w, h, d, _p(G), _p(rgbt), tf.n, C.byref(cam), C.byref(light), C.byref(params), _p(oc), _p(sd), W, H,
                    _p(out), _p(ns), C.byref(nsec))
    return (out, ns, nsec.value) if count else out


def gt_cube(vox_shape, cam, W, H, scale=(1.0, 1.0, 1.0)):
    """RedrawCube of rc1pcrtgt (vol_intersection.comp): the bounding-box placeholder frame."""
    o = orc()
    o.orc_gt_cube_render.argtypes = [C.c_void_p, C.POINTER(OrcCamera), C.c_int, C.c_int, C.c_void_p]
    d, h, w = vox_shape
    G = np.array([w * scale[0], h * scale[1], d * scale[2]], np.float32)
    out = np.zeros((H, W, 4), np.float32)
    assert o.orc_gt_cube_render(_p(G), C.byref(cam), W, H, _p(out)) == 0
    return out


def vct_supervoxels(vox):
    """(levels [(d,h,w,2) fp16-rounded float32 ...], max_stddev double)."""
    o = orc()
    o.orc_vct_supervoxels.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_double)]
    vox = np.ascontiguousarray(vox)
    d, h, w = vox.shape
    nlev = o.orc_vct_levels(w, h, d)
    dims = np.zeros((nlev, 3), np.int32)
    cap = int(vox.size * 2 * 1.2) + 64
    buf = np.zeros(cap, np.float32)
    ms = C.c_double(0.0)
    n = o.orc_vct_supervoxels(_p(vox), w, h, d, vox.dtype.itemsize, _p(buf), cap, _p(dims), C.byref(ms))
    assert n == nlev, n
    levels, off = [], 0
    for l in range(nlev):
        lw, lh, ld = (int(v) for v in dims[l])
        cnt = lw * lh * ld * 2
        levels.append(buf[off:off + cnt].reshape(ld, lh, lw, 2).copy())
        off += cnt
    return levels, dims, ms.value


def vct_preintegration(opc_by_density, dens_val, max_stddev, rows=None):
    o = orc()
    o.orc_vct_preintegration.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_int, C.c_void_p]
    opc = np.ascontiguousarray(opc_by_density, np.float32)
    w = int(np.ceil(float(dens_val))); h = int(np.ceil(max_stddev))
    r0, r1 = (0, h) if rows is None else rows
    out = np.zeros((r1 - r0, w), np.float32)
    o.orc_vct_preintegration(_p(opc), int(dens_val), float(max_stddev), r0, r1, _p(out))
    return out


def vct(vox, tf, levels, dims, lut, cam, light, params, W, H, scale=(1.0, 1.0, 1.0), count=False):
    o = orc()
    o.orc_vct_render.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                 C.c_int, C.c_void_p, C.c_int, C.POINTER(OrcCamera), C.POINTER(OrcLighting), C.POINTER(OrcVctParams),
                                 C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    tex = volume_r16f(vox)
    d, h, w = vox.shape
    sc = np.array(scale, np.float32)
    rgbt = tf.texture_rgbt()
    flat = np.ascontiguousarray(np.concatenate([l.ravel() for l in levels]).astype(np.float32))
    dims = np.ascontiguousarray(dims, np.int32)
    lut = np.ascontiguousarray(lut, np.float32)
    out = np.zeros((H, W, 4), np.float32)
    ns = np.zeros((H, W), np.uint32) if count else None
    o.orc_vct_render(_p(tex), w, h, d, _p(sc), _p(flat), _p(dims), len(dims), _p(lut), lut.shape[1], lut.shape[0], _p(rgbt), tf.n,
                     C.byref(cam), C.byref(light), C.byref(params), W, H, _p(out), _p(ns) if count else None)
    return (out, ns) if count else out


def ebs_light_cache(vox_shape, sat, light, params, res=(32, 32, 32), scale=(1.0, 1.0, 1.0)):
    """K9 rc1pextbsd/lightcachecomputation.comp: [rd, rh, rw, 2] fp16-rounded (Iao, Ids)."""
    o = orc()
    o.orc_ebs_light_cache.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    d, h, w = vox_shape
    sc = np.array(scale, np.float32)
    sat = np.ascontiguousarray(sat, np.float32)
    out = np.zeros((res[2], res[1], res[0], 2), np.float32)
    o.orc_ebs_light_cache(w, h, d, _p(sc), _p(sat), C.byref(light), C.byref(params), int(res[0]), int(res[1]), int(res[2]), _p(out))
    return out


def vct_light_cache(vox_shape, levels, dims, lut, light, params, res=(32, 32, 32), scale=(1.0, 1.0, 1.0)):
    """K13 rc1pvctsg/lightcachecomputation.comp: [rd, rh, rw, 2] fp16-rounded (1, Ivd)."""
    o = orc()
    o.orc_vct_light_cache.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                      C.POINTER(OrcLighting), C.POINTER(OrcVctParams), C.c_int, C.c_int, C.c_int, C.c_void_p]
    d, h, w = vox_shape
    sc = np.array(scale, np.float32)
    flat = np.ascontiguousarray(np.concatenate([l.ravel() for l in levels]).astype(np.float32))
    dims = np.ascontiguousarray(dims, np.int32)
    lut = np.ascontiguousarray(lut, np.float32)
    out = np.zeros((res[2], res[1], res[0], 2), np.float32)
    o.orc_vct_light_cache(w, h, d, _p(sc), _p(flat), _p(dims), len(dims), _p(lut), lut.shape[1], lut.shape[0], C.byref(light), C.byref(params),
                          int(res[0]), int(res[1]), int(res[2]), _p(out))
    return out


def copy_struct(src, dst_type):
    """Copy a ctypes struct of identical layout (product <-> oracle POD blocks)."""
    dst = dst_type()
    assert C.sizeof(src) == C.sizeof(dst)
    C.memmove(C.byref(dst), C.byref(src), C.sizeof(dst))
    return dst
