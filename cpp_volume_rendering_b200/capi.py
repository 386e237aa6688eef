"""ctypes bindings of include/vrb200.h (libvrb200.so) and of the vrbh_* driver surface of libvrbhost.so.

No compute happens in Python and nothing here falls back to a CPU path: a missing library raises VrbError.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


class VrbError(RuntimeError):
    pass


def lib_path():
    return os.path.join(_HERE, "libvrb200.so")


def host_lib_path():
    return os.path.join(_HERE, "libvrbhost.so")


# ---- POD structs of include/vrb200.h ---------------------------------------------------------------------------
class Camera(C.Structure):
    _fields_ = [("eye", C.c_float * 3), ("lookat", C.c_float * 16), ("tan_fovy", C.c_float), ("aspect", C.c_float)]


class Lighting(C.Structure):
    _fields_ = [("ka", C.c_float), ("kd", C.c_float), ("ks", C.c_float), ("shininess", C.c_float),
                ("ispecular", C.c_float * 3), ("light_pos", C.c_float * 3), ("light_forward", C.c_float * 3),
                ("light_up", C.c_float * 3), ("light_right", C.c_float * 3), ("spot_angle_deg", C.c_float),
                ("apply_phong", C.c_int)]


FILTER_PASS_MULTISAMPLE, FILTER_PASS_DOWNSCALE, FILTER_PASS_UPSCALE = 1, 2, 3
KERNEL_BOX, KERNEL_HAT, KERNEL_CATMULL_ROM, KERNEL_MITCHELL_NETRAVALI, KERNEL_CARDINAL_BSPLINE_3, KERNEL_CARDINAL_OMOMS3 = range(6)
GRADIENT_NONE, GRADIENT_SOBEL_FELDMAN, GRADIENT_FINITE_DIFFERENCES, GRADIENT_COMPUTE_SHADER_SOBEL = 0, 1, 2, 3


class IsoParams(C.Structure):
    _fields_ = [("isovalue", C.c_float), ("step_size_small", C.c_float), ("step_size_large", C.c_float), ("step_size_range", C.c_float),
                ("color", C.c_float * 4), ("count_samples", C.c_int)]


class Partition(C.Structure):
    _fields_ = [("rank", C.c_int), ("nranks", C.c_int), ("tile_w", C.c_int), ("tile_h", C.c_int)]


class Rc1passParams(C.Structure):
    _fields_ = [("step_size", C.c_float), ("count_samples", C.c_int), ("skip_empty", C.c_int)]


class EbsParams(C.Structure):
    _fields_ = [("step_size", C.c_float), ("apply_occlusion", C.c_int), ("apply_shadow", C.c_int),
                ("amb_occ_shells", C.c_int), ("amb_occ_radius", C.c_float), ("sdw_cone_angle_rad", C.c_float),
                ("sdw_sample_interval", C.c_float), ("sdw_initial_step", C.c_float), ("sdw_ui_weight", C.c_float),
                ("sdw_cone_max_distance", C.c_float), ("type_of_shadow", C.c_int), ("count_samples", C.c_int)]


BRICK_SEGMENT, BRICK_ALPHA, BRICK_EXACT = 0, 1, 2


class Brick(C.Structure):
    _fields_ = [("global_dims", C.c_int * 3), ("origin", C.c_int * 3), ("owned", C.c_int * 3),
                ("ghost_lo", C.c_int * 3), ("ghost_hi", C.c_int * 3)]


class ConeSampler(C.Structure):
    _fields_ = [("sections", C.c_void_p), ("n_sections", C.c_int), ("integration_samples", C.c_int * 3),
                ("initial_step", C.c_float), ("ray7_adj_weight", C.c_float), ("ui_weight", C.c_float),
                ("ray_axes", (C.c_float * 3) * 10)]


class DosParams(C.Structure):
    _fields_ = [("step_size", C.c_float), ("apply_occlusion", C.c_int), ("apply_shadow", C.c_int),
                ("type_of_shadow", C.c_int), ("spot_cos", C.c_float), ("count_samples", C.c_int)]


class ObjParams(C.Structure):
    _fields_ = [("step_size", C.c_float), ("apply_occlusion", C.c_int), ("apply_shadow", C.c_int), ("count_samples", C.c_int)]


class GtParams(C.Structure):
    _fields_ = [("step_size", C.c_float), ("light_ray_initial_gap", C.c_float), ("light_ray_step_size", C.c_float),
                ("apply_occlusion", C.c_int), ("occ_num_rays", C.c_int), ("occ_cone_distance", C.c_float),
                ("apply_shadow", C.c_int), ("sdw_num_rays", C.c_int), ("sdw_cone_distance", C.c_float),
                ("shadow_type", C.c_int), ("count_samples", C.c_int)]


class VctParams(C.Structure):
    _fields_ = [("step_size", C.c_float), ("apply_occlusion", C.c_int), ("apply_shadow", C.c_int),
                ("tan_cone_apex_angle", C.c_float), ("cone_step_size", C.c_float), ("cone_step_increase_rate", C.c_float),
                ("cone_initial_step", C.c_float), ("opacity_correction_factor", C.c_float), ("apply_opacity_correction", C.c_int),
                ("cone_number_of_samples", C.c_int), ("volume_max_density", C.c_float), ("volume_max_stddev", C.c_float),
                ("count_samples", C.c_int)]


_lib = None
_host = None

# name -> (restype, argtypes); every symbol include/vrb200.h declares
C_ABI = {
    "vrb_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "vrb_ctx_destroy": (C.c_int, [C.c_void_p]),
    "vrb_ctx_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vrb_ctx_synchronize": (C.c_int, [C.c_void_p]),
    "vrb_ctx_set_partition": (C.c_int, [C.c_void_p, C.POINTER(Partition)]),
    "vrb_ctx_set_filter": (C.c_int, [C.c_void_p, C.c_int]),
    "vrb_ctx_get_filter": (C.c_int, [C.c_void_p]),
    "vrb_last_error": (C.c_char_p, []),
    "vrb_version": (C.c_char_p, []),
    "vrb_launch_count": (C.c_uint64, [C.c_void_p]),
    "vrb_last_sample_count": (C.c_uint64, [C.c_void_p]),
    "vrb_last_aux_count": (C.c_uint64, [C.c_void_p]),
    "vrb_last_prepass_ms": (C.c_float, [C.c_void_p]),
    "vrb_sat_build_slab": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "vrb_sat_slab_plane": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "vrb_sat_finish_slab": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vrb_sat_device_ptr": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int)]),
    "vrb_sat_commit": (C.c_int, [C.c_void_p]),
    "vrb_step_multiples_exact_f": (C.c_int, [C.c_float, C.c_float]),
    "vrb_ctx_set_kernel_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "vrb_last_kernel_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_char_p)]),
    "vrb_sat_layout": (C.c_int, [C.c_void_p]),
    "vrb_measure_l1_bandwidth": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "vrb_measure_hbm_bandwidth": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "vrb_measure_gather_rate": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "vrb_measure_tex3d_rate": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "vrb_measure_ldg16_rate": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "vrb_volume_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "vrb_volume_upload_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "vrb_tf_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "vrb_frame_resize": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "vrb_frame_resize_multiscaling": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]),
    "vrb_frame_filter": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "vrb_filtered_frame_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "vrb_filtered_frame_read_rgba32f": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vrb_frame_clear": (C.c_int, [C.c_void_p]),
    "vrb_frame_read_rgba32f": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vrb_frame_read_rgba32f_async": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vrb_frame_read_wait": (C.c_int, [C.c_void_p, C.c_int]),
    "vrb_frame_set_target": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vrb_frame_extra": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "vrb_frame_device_ptr": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "vrb_rc1pass_render": (C.c_int, [C.c_void_p, C.POINTER(Camera), C.POINTER(Rc1passParams)]),
    "vrb_rc1pass_render_lit": (C.c_int, [C.c_void_p, C.POINTER(Camera), C.POINTER(Rc1passParams), C.POINTER(Lighting)]),
    "vrb_iso_render": (C.c_int, [C.c_void_p, C.POINTER(Camera), C.POINTER(Lighting), C.POINTER(IsoParams)]),
    "vrb_gradient_build": (C.c_int, [C.c_void_p, C.c_int]),
    "vrb_gradient_mode": (C.c_int, [C.c_void_p]),
    "vrb_gradient_read": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vrb_rc1pass_render_brick": (C.c_int, [C.c_void_p, C.POINTER(Camera), C.POINTER(Rc1passParams), C.POINTER(Brick)]),
    "vrb_partial_device_ptr": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "vrb_rc1pass_brick_alpha": (C.c_int, [C.c_void_p, C.POINTER(Camera), C.POINTER(Rc1passParams), C.POINTER(Brick)]),
    "vrb_brick_alpha_device_ptr": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "vrb_rc1pass_render_brick_exact": (C.c_int, [C.c_void_p, C.POINTER(Camera), C.POINTER(Rc1passParams), C.POINTER(Brick), C.POINTER(C.c_void_p), C.c_int]),
    "vrb_composite_sum": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int]),
    "vrb_composite_ordered": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int]),
    "vrb_ipc_export": (C.c_int, [C.c_void_p, C.c_void_p, C.c_char_p]),
    "vrb_ipc_import": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p)]),
    "vrb_ipc_close": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vrb_sat_set_order": (C.c_int, [C.c_void_p, C.c_int]),
    "vrb_sat_get_order": (C.c_int, [C.c_void_p]),
    "vrb_sat_build": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "vrb_sat_build_u64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "vrb_sat_read": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vrb_ebs_render": (C.c_int, [C.c_void_p, C.POINTER(Camera), C.POINTER(Lighting), C.POINTER(EbsParams)]),
    "vrb_extcoef_build": (C.c_int, [C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_int]),
    "vrb_extcoef_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.c_void_p, C.c_int]),
    "vrb_extcoef_read_level": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "vrb_dos_set_cones": (C.c_int, [C.c_void_p, C.POINTER(ConeSampler), C.POINTER(ConeSampler)]),
    "vrb_dos_light_cache_build": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "vrb_ebs_light_cache_build": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "vrb_vct_light_cache_build": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "vrb_light_cache_read": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "vrb_obj_march_render": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vrb_dos_render": (C.c_int, [C.c_void_p, C.POINTER(Camera), C.POINTER(Lighting), C.POINTER(DosParams)]),
    "vrb_gt_set_rays": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "vrb_gt_render": (C.c_int, [C.c_void_p, C.POINTER(Camera), C.POINTER(Lighting), C.POINTER(GtParams)]),
    "vrb_gt_cube_render": (C.c_int, [C.c_void_p, C.POINTER(Camera)]),
    "vrb_vct_build": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "vrb_vct_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_float)]),
    "vrb_vct_read": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "vrb_vct_render": (C.c_int, [C.c_void_p, C.POINTER(Camera), C.POINTER(Lighting), C.POINTER(VctParams)]),
    "vrb_vct_render_brick": (C.c_int, [C.c_void_p, C.POINTER(Camera), C.POINTER(Lighting), C.POINTER(VctParams), C.POINTER(Brick), C.c_int,
                                        C.POINTER(C.c_void_p), C.c_int]),
    "vrb_sv_build_brick": (C.c_int, [C.c_void_p, C.POINTER(Brick), C.c_int, C.POINTER(C.c_double)]),
    "vrb_sv_top_means_read": (C.c_int, [C.c_void_p, C.POINTER(Brick), C.c_void_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "vrb_sv_reduce_top": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "vrb_preint_build": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_double]),
}


def load():
    """Load libvrb200.so; raises VrbError when it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not os.path.exists(p):
        raise VrbError(f"{p} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback)")
    lib = C.CDLL(p, mode=C.RTLD_GLOBAL)
    for name, (res, args) in C_ABI.items():
        fn = getattr(lib, name)   # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def load_host():
    """Load libvrbhost.so (needs libvrb200.so)."""
    global _host
    if _host is not None:
        return _host
    load()
    p = host_lib_path()
    if not os.path.exists(p):
        raise VrbError(f"{p} is missing: run __graft_entry__.build()")
    h = C.CDLL(p)   # RTLD_LOCAL: the host mirrors the reference's class names on purpose
    h.vrbh_last_error.restype = C.c_char_p
    h.vrbh_ctx.restype = C.c_void_p
    h.vrbh_renderer_name.restype = C.c_char_p
    h.vrbh_renderer_name.argtypes = [C.c_int, C.c_int]
    h.vrbh_set_volume.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]
    h.vrbh_set_tf_points.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
    h.vrbh_set_param.argtypes = [C.c_char_p, C.c_double]
    h.vrbh_set_phong.argtypes = [C.c_float] * 4
    h.vrbh_read_rgba.argtypes = [C.c_void_p, C.c_size_t]
    h.vrbh_get_lighting.argtypes = [C.POINTER(Lighting)]
    h.vrbh_tf_create.restype = C.c_void_p
    h.vrbh_tf_create.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
    h.vrbh_tf_read.restype = C.c_void_p
    h.vrbh_tf_read.argtypes = [C.c_char_p]
    h.vrbh_tf_destroy.argtypes = [C.c_void_p]
    h.vrbh_tf_size.argtypes = [C.c_void_p]
    h.vrbh_tf_get.argtypes = [C.c_void_p, C.c_double, C.c_double, C.POINTER(C.c_float)]
    for n in ("vrbh_tf_get_extn", "vrbh_tf_get_opcn"):
        getattr(h, n).restype = C.c_float
        getattr(h, n).argtypes = [C.c_void_p, C.c_double]
    h.vrbh_tf_get_opc.restype = C.c_float
    h.vrbh_tf_get_opc.argtypes = [C.c_void_p, C.c_double, C.c_double]
    h.vrbh_tf_textures.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    h.vrbh_volume_read.restype = C.c_void_p
    h.vrbh_volume_read.argtypes = [C.c_char_p]
    h.vrbh_volume_destroy.argtypes = [C.c_void_p]
    h.vrbh_volume_info.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_ulonglong)]
    h.vrbh_volume_copy.argtypes = [C.c_void_p, C.c_void_p]
    h.vrbh_volume_normalized_sample.restype = C.c_double
    h.vrbh_volume_normalized_sample.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    h.vrbh_dds_decode.restype = C.c_longlong
    h.vrbh_dds_decode.argtypes = [C.c_void_p, C.c_ulonglong, C.c_void_p, C.c_ulonglong]
    h.vrbh_dds_encode.restype = C.c_longlong
    h.vrbh_dds_encode.argtypes = [C.c_void_p, C.c_ulonglong, C.c_uint, C.c_uint, C.c_int, C.c_void_p, C.c_ulonglong]
    h.vrbh_pvm_write.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    h.vrbh_read_camera_states.argtypes = [C.c_char_p, C.c_void_p, C.c_int]
    h.vrbh_read_light_lists.argtypes = [C.c_char_p, C.c_void_p, C.c_int]
    h.vrbh_look_at.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    h.vrbh_tan_fovy.restype = C.c_float
    h.vrbh_init_data.argtypes = [C.c_char_p]
    h.vrbh_load_volume.argtypes = [C.c_char_p]
    h.vrbh_load_tf.argtypes = [C.c_char_p]
    h.vrbh_set_renderer.argtypes = [C.c_char_p]
    h.vrbh_set_camera.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    h.vrbh_set_light_position.argtypes = [C.c_void_p]
    _host = h
    return h


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def make_camera(eye, center, up, width, height, fovy_deg=45.0):
    """Camera uniform block as the reference computes it: glm::lookAt in fp32, tan(fovy/2) in double -> float
    (rc1prenderer.cpp:94-101), aspect = float(w)/float(h) (libs/vis_utils/camera.cpp:307-310).
    The matrix comes from the C++ host (vrb::lookAt), not from Python arithmetic."""
    h = load_host()
    cam = Camera()
    e, c, u = _f32(eye), _f32(center), _f32(up)
    m = np.zeros(16, np.float32)
    h.vrbh_look_at(_ptr(e), _ptr(c), _ptr(u), _ptr(m))
    cam.eye[:] = e.tolist()
    cam.lookat[:] = m.tolist()
    cam.tan_fovy = np.float32(np.tan(np.float64(np.float32(fovy_deg)) * (np.pi / 180.0) / 2.0))
    cam.aspect = np.float32(np.float32(width) / np.float32(height))
    return cam


class Context:
    """Thin RAII wrapper over vrb_ctx; every method is one C-ABI call."""

    def __init__(self, device=0):
        self.lib = load()
        self.h = C.c_void_p()
        self._ck(self.lib.vrb_ctx_create(int(device), C.byref(self.h)))
        self.width = self.height = 0

    def _ck(self, rc):
        if rc != 0:
            raise VrbError(f"vrb error {rc}: {self.lib.vrb_last_error().decode()}")

    def close(self):
        if self.h:
            self.lib.vrb_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- plumbing
    def set_stream(self, cuda_stream_ptr):
        self._ck(self.lib.vrb_ctx_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def synchronize(self):
        self._ck(self.lib.vrb_ctx_synchronize(self.h))

    def set_filter(self, mode):
        """'exact' (software fp32 blends, bit-reproducible) or 'hardware' (texture units, like the reference's GL path)."""
        m = {"exact": 0, "hardware": 1, 0: 0, 1: 1}[mode]
        self._ck(self.lib.vrb_ctx_set_filter(self.h, m))

    def get_filter(self):
        return "hardware" if self.lib.vrb_ctx_get_filter(self.h) == 1 else "exact"

    def sat_set_order(self, order):
        """'reference' (BuildSAT's own fp64 recurrence, bit-identical floats) or 'scan' (three separable scans, faster)."""
        self._ck(self.lib.vrb_sat_set_order(self.h, {"reference": 0, "scan": 1, 0: 0, 1: 1}[order]))

    def sat_get_order(self):
        return "scan" if self.lib.vrb_sat_get_order(self.h) == 1 else "reference"

    def set_partition(self, rank, nranks, tile_w=64, tile_h=64):
        p = Partition(rank, nranks, tile_w, tile_h)
        self._ck(self.lib.vrb_ctx_set_partition(self.h, C.byref(p)))

    def set_kernel_timing(self, on):
        self._ck(self.lib.vrb_ctx_set_kernel_timing(self.h, int(bool(on))))

    def last_kernel_ms(self):
        """(duration in ms, kernel name) of the dominant kernel of the last render call made with kernel timing on."""
        ms = C.c_float(); nm = C.c_char_p()
        self._ck(self.lib.vrb_last_kernel_ms(self.h, C.byref(ms), C.byref(nm)))
        return float(ms.value), (nm.value or b"").decode()

    @property
    def launches(self):
        return int(self.lib.vrb_launch_count(self.h))

    @property
    def last_sample_count(self):
        return int(self.lib.vrb_last_sample_count(self.h))

    # -- inputs
    def volume_upload(self, vox, scale=(1.0, 1.0, 1.0)):
        """vox: numpy array indexed [z, y, x] (x fastest), dtype uint8 or uint16."""
        assert vox.dtype in (np.uint8, np.uint16) and vox.ndim == 3
        vox = np.ascontiguousarray(vox)
        d, h, w = vox.shape
        sc = (C.c_float * 3)(*scale)
        self._ck(self.lib.vrb_volume_upload(self.h, _ptr(vox), w, h, d, vox.dtype.itemsize, sc))

    def volume_upload_device(self, dev_ptr, w, h, d, bpv, scale=(1.0, 1.0, 1.0)):
        sc = (C.c_float * 3)(*scale)
        self._ck(self.lib.vrb_volume_upload_device(self.h, C.c_void_p(dev_ptr), w, h, d, bpv, sc))

    def tf_upload(self, rgbt, rgba=None):
        rgbt = _f32(rgbt).reshape(-1, 4)
        if rgba is not None:
            rgba = _f32(rgba).reshape(-1, 4)
        self._ck(self.lib.vrb_tf_upload(self.h, _ptr(rgbt), _ptr(rgba) if rgba is not None else None, rgbt.shape[0]))

    # -- frame
    def frame_resize(self, w, h):
        self._ck(self.lib.vrb_frame_resize(self.h, w, h))
        self.width, self.height = w, h

    # -- pixel multi-scaling (MULTIPLE_RAYS_PER_PIXEL / DOWN_SCALING_RENDER / UP_SCALING_RENDER)
    def frame_resize_multiscaling(self, screen_w, screen_h, mw, mh):
        """Rendered frame = screen * m (m > 0) or screen / |m| (m < 0) + a screen-sized filtered frame."""
        self._ck(self.lib.vrb_frame_resize_multiscaling(self.h, screen_w, screen_h, mw, mh))
        self.width = screen_w // abs(mw) if mw < 0 else screen_w * mw
        self.height = screen_h // abs(mh) if mh < 0 else screen_h * mh
        self.screen_width, self.screen_height = screen_w, screen_h

    def frame_filter(self, pass_id, kernel=1):
        self._ck(self.lib.vrb_frame_filter(self.h, int(pass_id), int(kernel)))

    def filtered_frame_read(self):
        w = C.c_int(); h = C.c_int()
        self._ck(self.lib.vrb_filtered_frame_info(self.h, None, C.byref(w), C.byref(h)))
        out = np.empty((h.value, w.value, 4), np.float32)
        self._ck(self.lib.vrb_filtered_frame_read_rgba32f(self.h, _ptr(out)))
        return out

    def frame_read(self, out=None):
        if out is None:
            out = np.empty((self.height, self.width, 4), np.float32)
        self._ck(self.lib.vrb_frame_read_rgba32f(self.h, _ptr(out)))
        return out

    def frame_read_into(self, host_ptr):
        self._ck(self.lib.vrb_frame_read_rgba32f(self.h, C.c_void_p(host_ptr)))

    def frame_read_async(self, host_ptr):
        """Queue the read of the current frame into (page-locked) host memory; returns immediately."""
        self._ck(self.lib.vrb_frame_read_rgba32f_async(self.h, C.c_void_p(host_ptr)))

    def frame_read_wait(self, max_in_flight=0):
        self._ck(self.lib.vrb_frame_read_wait(self.h, int(max_in_flight)))

    def frame_set_target(self, dev_ptr):
        """Redirect pixel stores / read-backs to another RGBA16F buffer (None = own frame); see vrb_frame_set_target."""
        self._ck(self.lib.vrb_frame_set_target(self.h, C.c_void_p(dev_ptr) if dev_ptr else None))

    def frame_extra(self, index):
        p = C.c_void_p()
        self._ck(self.lib.vrb_frame_extra(self.h, int(index), C.byref(p)))
        return p.value

    def frame_device_ptr(self):
        p = C.c_void_p(); w = C.c_int(); h = C.c_int()
        self._ck(self.lib.vrb_frame_device_ptr(self.h, C.byref(p), C.byref(w), C.byref(h)))
        return p.value, w.value, h.value

    # -- renderers
    def rc1pass_render(self, cam, step_size=0.5, count_samples=False, skip_empty=False):
        p = Rc1passParams(step_size, int(count_samples), int(skip_empty))
        self._ck(self.lib.vrb_rc1pass_render(self.h, C.byref(cam), C.byref(p)))

    def rc1pass_render_lit(self, cam, light, step_size=0.5, count_samples=False):
        """rc1pass with the lighting uniforms: light.apply_phong = 1 runs ShadeBlinnPhong (needs gradient_build)."""
        p = Rc1passParams(step_size, int(count_samples), 0)
        self._ck(self.lib.vrb_rc1pass_render_lit(self.h, C.byref(cam), C.byref(p), C.byref(light)))

    def iso_render(self, cam, light, params):
        self._ck(self.lib.vrb_iso_render(self.h, C.byref(cam), C.byref(light), C.byref(params)))

    def gradient_build(self, mode):
        self._ck(self.lib.vrb_gradient_build(self.h, int(mode)))

    def gradient_mode(self):
        return int(self.lib.vrb_gradient_mode(self.h))

    def gradient_read(self, shape_zyx):
        """RGB16F gradient texels as float32 (d, h, w, 3)."""
        d, h, w = shape_zyx
        a = np.empty((d, h, w, 3), np.float32)
        self._ck(self.lib.vrb_gradient_read(self.h, _ptr(a)))
        return a

    # -- sort-last
    def rc1pass_render_brick(self, cam, brick, step_size=0.5, count_samples=False):
        p = Rc1passParams(step_size, int(count_samples), 0)
        self._ck(self.lib.vrb_rc1pass_render_brick(self.h, C.byref(cam), C.byref(p), C.byref(brick)))

    def rc1pass_brick_alpha(self, cam, brick, step_size=0.5):
        p = Rc1passParams(step_size, 0, 0)
        self._ck(self.lib.vrb_rc1pass_brick_alpha(self.h, C.byref(cam), C.byref(p), C.byref(brick)))

    def brick_alpha_device_ptr(self):
        p = C.c_void_p()
        self._ck(self.lib.vrb_brick_alpha_device_ptr(self.h, C.byref(p)))
        return p.value

    def rc1pass_render_brick_exact(self, cam, brick, front_alpha_ptrs, step_size=0.5, count_samples=False):
        p = Rc1passParams(step_size, int(count_samples), 0)
        arr = (C.c_void_p * max(1, len(front_alpha_ptrs)))(*front_alpha_ptrs)
        self._ck(self.lib.vrb_rc1pass_render_brick_exact(self.h, C.byref(cam), C.byref(p), C.byref(brick), arr, len(front_alpha_ptrs)))

    def composite_sum(self, partial_ptrs, row0=0, rows=None):
        arr = (C.c_void_p * len(partial_ptrs))(*partial_ptrs)
        self._ck(self.lib.vrb_composite_sum(self.h, arr, len(partial_ptrs), row0, self.height if rows is None else rows))

    def partial_device_ptr(self):
        p = C.c_void_p()
        self._ck(self.lib.vrb_partial_device_ptr(self.h, C.byref(p)))
        return p.value

    def composite_ordered(self, partial_ptrs, row0=0, rows=None):
        arr = (C.c_void_p * len(partial_ptrs))(*partial_ptrs)
        self._ck(self.lib.vrb_composite_ordered(self.h, arr, len(partial_ptrs), row0, self.height if rows is None else rows))

    def ipc_export(self, dev_ptr):
        buf = C.create_string_buffer(64)
        self._ck(self.lib.vrb_ipc_export(self.h, C.c_void_p(dev_ptr), buf))
        return buf.raw

    def ipc_import(self, handle):
        p = C.c_void_p()
        self._ck(self.lib.vrb_ipc_import(self.h, handle, C.byref(p)))
        return p.value

    def ipc_close(self, dev_ptr):
        self._ck(self.lib.vrb_ipc_close(self.h, C.c_void_p(dev_ptr)))

    def sat_build(self, ext_lut):
        lut = _f32(ext_lut)
        self._ck(self.lib.vrb_sat_build(self.h, _ptr(lut), lut.size))

    # -- sharded SAT build (one z-slab per rank; the exchange is in dist.sat_build_sharded)
    def sat_build_slab(self, ext_lut, z_lo, z_hi):
        lut = _f32(ext_lut)
        self._ck(self.lib.vrb_sat_build_slab(self.h, _ptr(lut), lut.size, int(z_lo), int(z_hi)))

    def sat_slab_plane(self):
        """(device pointer, element count) of the slab's last fp64 plane."""
        p = C.c_void_p(); n = C.c_size_t()
        self._ck(self.lib.vrb_sat_slab_plane(self.h, C.byref(p), C.byref(n)))
        return p.value, int(n.value)

    def sat_finish_slab(self, prefix_plane_ptr):
        self._ck(self.lib.vrb_sat_finish_slab(self.h, C.c_void_p(prefix_plane_ptr) if prefix_plane_ptr else None))

    def sat_device_ptr(self):
        p = C.c_void_p(); dims = (C.c_int * 3)()
        self._ck(self.lib.vrb_sat_device_ptr(self.h, C.byref(p), dims))
        return p.value, (dims[0], dims[1], dims[2])

    def sat_commit(self):
        self._ck(self.lib.vrb_sat_commit(self.h))

    def sat_build_u64(self, lut_u32, shape_zyx):
        lut = np.ascontiguousarray(lut_u32, dtype=np.uint32)
        out = np.empty(shape_zyx, np.uint64)
        self._ck(self.lib.vrb_sat_build_u64(self.h, _ptr(lut), lut.size, _ptr(out)))
        return out

    def sat_read(self, shape_zyx):
        d, h, w = shape_zyx
        out = np.empty((d + 2, h + 2, w + 2), np.float32)
        self._ck(self.lib.vrb_sat_read(self.h, _ptr(out)))
        return out

    def ebs_render(self, cam, light, params):
        self._ck(self.lib.vrb_ebs_render(self.h, C.byref(cam), C.byref(light), C.byref(params)))

    @property
    def last_aux_count(self):
        return int(self.lib.vrb_last_aux_count(self.h))

    def extcoef_build(self, sigma0=1.0, res=(128, 128, 128)):
        rw, rh, rd = res if res else (0, 0, 0)
        self._ck(self.lib.vrb_extcoef_build(self.h, sigma0, rw, rh, rd))

    def extcoef_levels(self):
        """Returns [(float32 array [d,h,w]), ...] for every pyramid level (extinction, fp16-rounded)."""
        n = C.c_int()
        dims = np.zeros((16, 3), np.int32)
        self._ck(self.lib.vrb_extcoef_info(self.h, C.byref(n), _ptr(dims), 16))
        out = []
        for l in range(n.value):
            w, h, d = (int(v) for v in dims[l])
            a = np.empty((d, h, w), np.float32)
            self._ck(self.lib.vrb_extcoef_read_level(self.h, l, _ptr(a)))
            out.append(a)
        return out

    def dos_set_cones(self, occ, sdw):
        self._ck(self.lib.vrb_dos_set_cones(self.h, C.byref(occ), C.byref(sdw)))

    def dos_render(self, cam, light, params):
        self._ck(self.lib.vrb_dos_render(self.h, C.byref(cam), C.byref(light), C.byref(params)))

    def dos_light_cache_build(self, eye, eye_up, light, params, res=(32, 32, 32)):
        e = _f32(eye); u = _f32(eye_up)
        self._ck(self.lib.vrb_dos_light_cache_build(self.h, _ptr(e), _ptr(u), C.byref(light), C.byref(params), int(res[0]), int(res[1]), int(res[2])))

    def ebs_light_cache_build(self, light, params, res=(32, 32, 32)):
        self._ck(self.lib.vrb_ebs_light_cache_build(self.h, C.byref(light), C.byref(params), int(res[0]), int(res[1]), int(res[2])))

    def vct_light_cache_build(self, light, params, res=(32, 32, 32)):
        self._ck(self.lib.vrb_vct_light_cache_build(self.h, C.byref(light), C.byref(params), int(res[0]), int(res[1]), int(res[2])))

    def light_cache_read(self):
        dims = (C.c_int * 3)()
        self._ck(self.lib.vrb_light_cache_read(self.h, None, dims))
        out = np.zeros((dims[2], dims[1], dims[0], 2), np.float32)
        self._ck(self.lib.vrb_light_cache_read(self.h, _ptr(out), dims))
        return out

    def obj_march_render(self, cam, light, step_size, apply_occlusion=1, apply_shadow=0, count_samples=False):
        p = ObjParams(step_size, int(apply_occlusion), int(apply_shadow), int(count_samples))
        self._ck(self.lib.vrb_obj_march_render(self.h, C.byref(cam), C.byref(light), C.byref(p)))

    def gt_set_rays(self, occ, sdw):
        occ = _f32(occ).reshape(-1, 3); sdw = _f32(sdw).reshape(-1, 3)
        self._ck(self.lib.vrb_gt_set_rays(self.h, _ptr(occ), occ.shape[0], _ptr(sdw), sdw.shape[0]))

    def gt_render(self, cam, light, params):
        self._ck(self.lib.vrb_gt_render(self.h, C.byref(cam), C.byref(light), C.byref(params)))

    def gt_cube_render(self, cam):
        """RedrawCube of rc1pcrtgt: the bounding-box placeholder frame (vol_intersection.comp)."""
        self._ck(self.lib.vrb_gt_cube_render(self.h, C.byref(cam)))

    def vct_build(self, opc_by_density):
        o = _f32(opc_by_density)
        self._ck(self.lib.vrb_vct_build(self.h, _ptr(o), o.size))

    def vct_info(self):
        n = C.c_int(); lw = C.c_int(); lh = C.c_int(); ms = C.c_float()
        dims = np.zeros((16, 3), np.int32)
        self._ck(self.lib.vrb_vct_info(self.h, C.byref(n), _ptr(dims), 16, C.byref(lw), C.byref(lh), C.byref(ms)))
        return dims[:n.value].copy(), (lw.value, lh.value), ms.value

    def vct_read(self):
        """(levels [(d,h,w,2) float32 ...], lut (h,w) float32, max_stddev)."""
        dims, (lw, lh), ms = self.vct_info()
        levels = []
        for l, (w, h, d) in enumerate(dims):
            a = np.empty((int(d), int(h), int(w), 2), np.float32)
            self._ck(self.lib.vrb_vct_read(self.h, l, _ptr(a)))
            levels.append(a)
        lut = np.empty((lh, lw), np.float32)
        self._ck(self.lib.vrb_vct_read(self.h, -1, _ptr(lut)))
        return levels, lut, ms

    def vct_render(self, cam, light, params):
        self._ck(self.lib.vrb_vct_render(self.h, C.byref(cam), C.byref(light), C.byref(params)))

    # -- sort-last bricks of the VCT renderer
    def sv_build_brick(self, brick, n_levels):
        """Pyramid of the uploaded window (levels 0..n_levels-1) -> largest deviation found in it."""
        m = C.c_double()
        self._ck(self.lib.vrb_sv_build_brick(self.h, C.byref(brick), int(n_levels), C.byref(m)))
        return m.value

    def sv_top_means(self, brick):
        """fp64 means of the brick's owned texels of the last level -> (array (d,h,w), origin (x,y,z) in that level)."""
        dims = (C.c_int * 3)(); org = (C.c_int * 3)()
        self._ck(self.lib.vrb_sv_top_means_read(self.h, C.byref(brick), None, 0, dims, org))
        a = np.empty((dims[2], dims[1], dims[0]), np.float64)
        self._ck(self.lib.vrb_sv_top_means_read(self.h, C.byref(brick), _ptr(a), a.size, dims, org))
        return a, (org[0], org[1], org[2])

    def sv_reduce_top(self, level_means):
        a = np.ascontiguousarray(level_means, np.float64)
        m = C.c_double()
        self._ck(self.lib.vrb_sv_reduce_top(self.h, _ptr(a), a.shape[2], a.shape[1], a.shape[0], C.byref(m)))
        return m.value

    def preint_build(self, opc_by_density, max_stddev):
        o = _f32(opc_by_density)
        self._ck(self.lib.vrb_preint_build(self.h, _ptr(o), o.size, float(max_stddev)))

    def vct_render_brick(self, cam, light, params, brick, mode, front_alpha_ptrs=()):
        arr = (C.c_void_p * max(1, len(front_alpha_ptrs)))(*front_alpha_ptrs)
        self._ck(self.lib.vrb_vct_render_brick(self.h, C.byref(cam), C.byref(light), C.byref(params), C.byref(brick), int(mode),
                                               arr, len(front_alpha_ptrs)))


def host_tf_arrays(tf_points, bpv):
    """RGBt / RGBA transfer-function textures (256 texels) + the per-voxel-value extinction LUT the SAT build reads, from the C++
    host mirror (vis::TransferFunction1D of libvrbhost.so: product code, not the oracle).  tf_points = (rgb points, alpha points)."""
    h = load_host()
    rgb, a = tf_points
    rgb = np.ascontiguousarray(rgb); a = np.ascontiguousarray(a)
    tf = h.vrbh_tf_create(_ptr(rgb), len(rgb), _ptr(a), len(a), 255, 0)
    rgbt = np.zeros((256, 4), np.float32); rgba = np.zeros((256, 4), np.float32)
    assert h.vrbh_tf_textures(tf, _ptr(rgbt), _ptr(rgba), 256) == 256
    nv = 256 if bpv == 1 else 65536
    mx = 255.0 if bpv == 1 else 65535.0
    lut = np.array([h.vrbh_tf_get_extn(tf, v / mx) for v in range(nv)], np.float32)
    h.vrbh_tf_destroy(tf)
    return rgbt, rgba, lut


def host_gt_ray_tables(n_occ, occ_aperture_deg, n_sdw, sdw_aperture_deg):
    """Ray tables as RC1PConeLightGroundTruthSteps::Update of the C++ host draws them."""
    h = load_host()
    h.vrbh_gt_ray_tables.argtypes = [C.c_int, C.c_float, C.c_int, C.c_float, C.c_void_p, C.c_void_p]
    occ = np.zeros((max(n_occ, 1), 3), np.float32); sdw = np.zeros((max(n_sdw, 1), 3), np.float32)
    h.vrbh_gt_ray_tables(n_occ, occ_aperture_deg, n_sdw, sdw_aperture_deg, _ptr(occ), _ptr(sdw))
    return occ[:n_occ], sdw[:n_sdw]


def host_opacity_by_density(tfname_points, bpv):
    """tf->GetOpc(i, maxDensity) for i = 0..maxDensity from the C++ host TransferFunction1D."""
    h = load_host()
    rgb, a = tfname_points
    rgb = np.ascontiguousarray(rgb, np.float64); a = np.ascontiguousarray(a, np.float64)
    tf = h.vrbh_tf_create(_ptr(rgb), len(rgb), _ptr(a), len(a), 255, 0)
    mx = 255 if bpv == 1 else 65535
    out = np.array([h.vrbh_tf_get_opc(tf, float(i), float(mx)) for i in range(mx + 1)], np.float32)
    h.vrbh_tf_destroy(tf)
    return out


def default_gt_params(diagonal, n_occ, n_sdw, step_size=0.5, apply_occlusion=True, apply_shadow=True):
    """Constructor / Init defaults of RC1PConeLightGroundTruthSteps (crtgtrenderer.cpp:35-51,106-113)."""
    p = GtParams()
    p.step_size = step_size
    p.light_ray_initial_gap = 1.0
    p.light_ray_step_size = 0.5
    p.apply_occlusion = int(apply_occlusion); p.occ_num_rays = n_occ; p.occ_cone_distance = np.float32(diagonal * 0.5)
    p.apply_shadow = int(apply_shadow); p.sdw_num_rays = n_sdw; p.sdw_cone_distance = np.float32(diagonal * 0.75)
    p.shadow_type = 0
    p.count_samples = 0
    return p


def default_vct_params(max_density, max_stddev, step_size=0.5):
    """Constructor defaults of RC1PVoxelConeTracingSGPU (vctrenderer.cpp:33-43,145)."""
    p = VctParams()
    p.step_size = step_size
    p.apply_occlusion = 1
    p.apply_shadow = 1
    p.tan_cone_apex_angle = np.float32(np.tan(np.float32(2.0) * np.float32(np.pi) / np.float32(180.0)))
    p.cone_step_size = 2.0
    p.cone_step_increase_rate = 1.0
    p.cone_initial_step = 2.0
    p.opacity_correction_factor = 2.0
    p.apply_opacity_correction = 1
    p.cone_number_of_samples = 50
    p.volume_max_density = max_density
    p.volume_max_stddev = max_stddev
    p.count_samples = 0
    return p


def host_cone_sampler(half_angle, max_packing, covered_distance, ui_weight, initial_step=3.0, d_sigma=1.25, r_sigma=2.0,
                      min_sigma=1.0):
    """ConeGaussianSampler of the C++ host mirror -> (vrb_cone_sampler block, sections array)."""
    h = load_host()
    h.vrbh_cone_sampler_compute.argtypes = [C.c_float] * 2 + [C.c_int] + [C.c_float] * 4 + [C.c_double, C.c_void_p, C.c_int,
                                                                                          C.c_void_p, C.c_void_p, C.c_void_p]
    sec = np.zeros((4096, 4), np.float32)
    counts = (C.c_int * 3)()
    axes = np.zeros((10, 3), np.float32)
    adj = np.zeros(2, np.float32)
    n = h.vrbh_cone_sampler_compute(half_angle, initial_step, max_packing, covered_distance, d_sigma, r_sigma, ui_weight,
                                    float(min_sigma), _ptr(sec), 4096, counts, _ptr(axes), _ptr(adj))
    if n < 0:
        raise VrbError(f"cone sampler failed ({n}): {h.vrbh_last_error().decode()}")
    sec = np.ascontiguousarray(sec[:n])
    cs = ConeSampler()
    cs._keep = sec
    cs.sections = sec.ctypes.data
    cs.n_sections = n
    cs.integration_samples[:] = list(counts)
    cs.initial_step = max(initial_step, 0.0)
    cs.ray7_adj_weight = float(adj[1])
    cs.ui_weight = ui_weight
    for i in range(10):
        for j in range(3):
            cs.ray_axes[i][j] = float(axes[i, j])
    return cs, sec, float(adj[0])


def default_dos_params(step_size=0.5, apply_shadow=False, spot_angle_deg=4.0):
    """Constructor defaults of RC1PConeTracingDirOcclusionShading (dosrcrenderer.cpp:42-59,159)."""
    p = DosParams()
    p.step_size = step_size
    p.apply_occlusion = 1
    p.apply_shadow = int(apply_shadow)
    p.type_of_shadow = 0
    p.spot_cos = np.float32(np.cos(np.float32(np.pi) * np.float32(spot_angle_deg) / np.float32(180.0)))
    p.count_samples = 0
    return p


def default_lighting(light_pos=(0.0, 0.0, 0.0), forward=(0.0, 0.0, 1.0), up=(0.0, 1.0, 0.0), right=(1.0, 0.0, 0.0)):
    """Defaults of renderingparameters.cpp:17-32 and lightsourcelist.cpp:21-37."""
    L = Lighting()
    L.ka, L.kd, L.ks, L.shininess = 0.5, 0.5, 0.8, 30.0
    L.ispecular[:] = [1.0, 1.0, 1.0]
    L.light_pos[:] = list(light_pos)
    L.light_forward[:] = list(forward)
    L.light_up[:] = list(up)
    L.light_right[:] = list(right)
    L.spot_angle_deg = 4.0
    return L


def default_iso_params():
    """Constructor defaults of RayCasting1PassIsoAdapt (rc1pisoadaptrenderer.cpp:13-22)."""
    p = IsoParams()
    p.isovalue, p.step_size_small, p.step_size_large, p.step_size_range = 0.5, 0.05, 1.0, 0.1
    p.color[:] = [0.66, 0.6, 0.05, 1.0]
    p.count_samples = 0
    return p


def default_ebs_params(diagonal, step_size=0.5):
    """Constructor defaults of RC1PExtinctionBasedShading (ebsrenderer.cpp:29-42,105,166)."""
    p = EbsParams()
    p.step_size = step_size
    p.apply_occlusion = 1
    p.apply_shadow = 1
    p.amb_occ_shells = 15
    p.amb_occ_radius = 1.0
    p.sdw_cone_angle_rad = np.float32(1.0 * np.pi / 180.0)
    p.sdw_sample_interval = 2.0
    p.sdw_initial_step = 2.0
    p.sdw_ui_weight = 1.0
    p.sdw_cone_max_distance = np.float32(0.75) * np.float32(diagonal)
    p.type_of_shadow = 0
    p.count_samples = 0
    return p
