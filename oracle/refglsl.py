"""oracle/refglsl.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes binding of oracle/_ref/librefglsl.so: the REFERENCE's own GLSL compute shaders, compiled for the CPU from where
they lie under /root/reference (oracle/glsl_cpu/).  The API is shaped like the GL calls of the reference's renderer
classes: create a program, set uniforms / textures / images by name, dispatch.  Tests use it to pin the oracle's
restatement of every marcher against the shader source itself."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "librefglsl.so")
_lib = None


def build(force=False):
    srcs = [os.path.join(_HERE, "glsl_cpu", f) for f in os.listdir(os.path.join(_HERE, "glsl_cpu"))] + [os.path.join(_HERE, "oracle_common.h")]
    stale = not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if os.path.isdir("/root/reference") and (force or stale):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "refglsl"])


def lib():
    """None when the library was never built and cannot be built here (no /root/reference)."""
    global _lib
    if _lib is None:
        build()
        if not os.path.exists(_SO):
            return None
        L = C.CDLL(_SO)
        L.rg_program_create.restype = C.c_void_p
        L.rg_program_create.argtypes = [C.c_char_p]
        L.rg_program_destroy.argtypes = [C.c_void_p]
        L.rg_program_names.argtypes = [C.c_char_p, C.c_int]
        L.rg_set_f.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int]
        L.rg_set_i.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int]
        L.rg_texture_create.restype = C.c_void_p
        L.rg_texture_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.rg_texture_destroy.argtypes = [C.c_void_p]
        L.rg_set_texture.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        L.rg_image_create.restype = C.c_void_p
        L.rg_image_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.rg_image_destroy.argtypes = [C.c_void_p]
        L.rg_set_image.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        L.rg_dispatch.restype = C.c_long
        L.rg_dispatch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        for n in ("rg_unset_uniforms", "rg_unknown_uniforms"):
            getattr(L, n).argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.rg_first_fault.argtypes = [C.c_char_p, C.c_int]
        _lib = L
    return _lib


def program_names():
    buf = C.create_string_buffer(4096)
    lib().rg_program_names(buf, 4096)
    return buf.value.decode().split()


class Texture:
    """GL texture object: list of mip levels (numpy float32 arrays already rounded to the internal format).
    3-D levels are (d, h, w) or (d, h, w, c); 2-D levels (h, w) or (h, w, c); 1-D levels (n, c)."""

    def __init__(self, levels, dims):
        if not isinstance(levels, (list, tuple)):
            levels = [levels]
        self.levels = [np.ascontiguousarray(l, np.float32) for l in levels]
        l0 = self.levels[0]
        if dims == 1:
            self.channels = l0.shape[1] if l0.ndim == 2 else 1
            whd = [(l.shape[0], 1, 1) for l in self.levels]
        elif dims == 2:
            self.channels = l0.shape[2] if l0.ndim == 3 else 1
            whd = [(l.shape[1], l.shape[0], 1) for l in self.levels]
        elif dims == 3:
            self.channels = l0.shape[3] if l0.ndim == 4 else 1
            whd = [(l.shape[2], l.shape[1], l.shape[0]) for l in self.levels]
        else:
            raise ValueError(dims)
        self._whd = np.array(whd, np.int32)
        self._ptrs = (C.c_void_p * len(self.levels))(*[l.ctypes.data for l in self.levels])
        self.handle = lib().rg_texture_create(dims, self.channels, len(self.levels), self._whd.ctypes.data_as(C.c_void_p), self._ptrs)

    def __del__(self):
        if getattr(self, "handle", None) and _lib is not None:
            _lib.rg_texture_destroy(self.handle)
            self.handle = None


class Image:
    """GL image bound with glBindImageTexture: an (h, w, c) [image2D] or (d, h, w, c) [image3D] float32 array the
    shaders load from / store into (it may be one level of a Texture: same numpy array); half=True rounds stores to
    fp16 (rgba16f / rg16f / r16f)."""

    def __init__(self, array, half=True):
        assert array.dtype == np.float32 and array.flags.c_contiguous and array.ndim in (3, 4)
        self.array = array
        d = array.shape[0] if array.ndim == 4 else 1
        h, w, c = array.shape[-3:]
        self.handle = lib().rg_image_create(array.ctypes.data_as(C.c_void_p), w, h, d, c, 1 if half else 0)

    def __del__(self):
        if getattr(self, "handle", None) and _lib is not None:
            _lib.rg_image_destroy(self.handle)
            self.handle = None


class Program:
    def __init__(self, name):
        self.name = name
        self.handle = lib().rg_program_create(name.encode())
        if not self.handle:
            raise KeyError(f"no reference shader program '{name}' in librefglsl.so ({program_names()})")
        self._keep = {}

    def __del__(self):
        if getattr(self, "handle", None) and _lib is not None:
            _lib.rg_program_destroy(self.handle)
            self.handle = None

    def set(self, name, value):
        """glUniform*: python int / numpy integer arrays set int uniforms, everything else float uniforms."""
        a = np.asarray(value)
        if a.dtype.kind in "iub":
            a = np.ascontiguousarray(a, np.int32).ravel()
            lib().rg_set_i(self.handle, name.encode(), a.ctypes.data_as(C.c_void_p), a.size)
        else:
            a = np.ascontiguousarray(a, np.float32).ravel()
            lib().rg_set_f(self.handle, name.encode(), a.ctypes.data_as(C.c_void_p), a.size)

    def set_many(self, **kw):
        for k, v in kw.items():
            self.set(k, v)

    def texture(self, name, tex):
        self._keep[name] = tex
        lib().rg_set_texture(self.handle, name.encode(), tex.handle)

    def image(self, name, img):
        self._keep[name] = img
        lib().rg_set_image(self.handle, name.encode(), img.handle)

    def dispatch(self, width, height, depth=1, local=(8, 8, 1)):
        """ComputeShader::RecomputeNumberOfGroups + Dispatch: ceil(size / local) groups per axis; `local` is the
        shader's layout(local_size_*) (8 x 8 x 1 for the frame shaders, 8 x 8 x 8 for the volume ones)."""
        local = tuple(local) + (1,) * (3 - len(local))
        groups = np.array([-(-width // local[0]), -(-height // local[1]), -(-depth // local[2])], np.int32)
        loc = np.array(local, np.int32)
        faults = lib().rg_dispatch(self.handle, groups.ctypes.data_as(C.c_void_p), loc.ctypes.data_as(C.c_void_p))
        if faults:
            buf = C.create_string_buffer(256)
            lib().rg_first_fault(buf, 256)
            raise RuntimeError(f"{self.name}: {faults} sampler faults, first: {buf.value.decode()}")

    def _names(self, fn):
        buf = C.create_string_buffer(8192)
        getattr(lib(), fn)(self.handle, buf, 8192)
        return sorted(set(buf.value.decode().split()))

    def unset_uniforms(self):
        """Uniforms some linked shader declares that were never set (valid after a dispatch)."""
        return self._names("rg_unset_uniforms")

    def unknown_uniforms(self):
        """Uniforms that were set but that no linked shader declares."""
        return self._names("rg_unknown_uniforms")


def camera_uniforms(cam):
    """(eye vec3, lookAt mat4 column-major 16 floats, tan(fovy/2), aspect) from an oracle OrcCamera."""
    return (np.array(list(cam.eye), np.float32), np.array(list(cam.lookat), np.float32), np.float32(cam.tan_fovy), np.float32(cam.aspect))


# ---- the reference's renderer classes, replayed: same uniform names, same dispatch sequences ---------------------------
# Arguments mirror oracle/bind.py's wrappers so that a test can call both with the same objects.
def _v3(a):
    return np.array(list(a), np.float32)


def _grid(vox, scale=(1.0, 1.0, 1.0)):
    d, h, w = vox.shape
    return np.array([w * scale[0], h * scale[1], d * scale[2]], np.float32)


def _volume_and_tf(p, vox, tf):
    from oracle import bind
    p.texture("TexVolume", Texture(bind.volume_r16f(vox), 3))
    p.texture("TexTransferFunc", Texture(tf.texture_rgbt(), 1))


def _frame(p, W, H, allowed_unset=(), allowed_unknown=()):
    out = np.zeros((H, W, 4), np.float32)
    p.image("OutputFrag", Image(out))
    p.dispatch(W, H)
    unknown = set(p.unknown_uniforms()) - set(allowed_unknown)
    assert not unknown, f"{p.name}: uniforms set that the shader does not declare: {sorted(unknown)}"
    unset = set(p.unset_uniforms()) - set(allowed_unset)
    assert not unset, f"{p.name}: uniforms the shader declares that were never set: {sorted(unset)}"
    return out


def _lit_uniforms(p, light, eye):
    """ebsrenderer.cpp:223-245, dosrcrenderer.cpp:221-243, vctrenderer.cpp:211-233 (same names in the three)."""
    p.set_many(Kambient=light.ka, Kdiffuse=light.kd, Kspecular=light.ks, Nshininess=light.shininess, Ispecular=_v3(light.ispecular),
               WorldEyePos=eye, WorldLightingPos=_v3(light.light_pos))


def _one_pass_phong_uniforms(p, light, eye):
    """rc1prenderer.cpp:112-132 / rc1pisoadaptrenderer.cpp (same names)."""
    p.set_many(BlinnPhongKa=light.ka, BlinnPhongKd=light.kd, BlinnPhongKs=light.ks, BlinnPhongShininess=light.shininess,
               BlinnPhongIspecular=_v3(light.ispecular), WorldEyePos=eye, LightSourcePosition=_v3(light.light_pos))


def run_rc1pass(vox, tf, cam, light, W, H, step=0.5, scale=(1.0, 1.0, 1.0), grad=None):
    """RayCasting1Pass: CreateRenderingPass (rc1prenderer.cpp:231-262) + Update (:72-138) + Redraw."""
    return _frame(make_rc1pass(vox, tf, cam, light, step, scale, grad), W, H, allowed_unset=("ProjectionMatrix", "VolumeScales", "TexVolumeGradient"))


def make_rc1pass(vox, tf, cam, light, step=0.5, scale=(1.0, 1.0, 1.0), grad=None):
    """The bound program of run_rc1pass, ready to dispatch (bench.py times the dispatch alone)."""
    p = Program("rc1pass")
    _volume_and_tf(p, vox, tf)
    phong = 1 if (grad is not None and light.apply_phong == 1) else 0
    if phong:
        p.texture("TexVolumeGradient", Texture(grad, 3))
    d, h, w = vox.shape
    e, look, tanf, asp = camera_uniforms(cam)
    p.set_many(VolumeGridResolution=np.array([w, h, d], np.float32), VolumeVoxelSize=np.array(scale, np.float32), VolumeGridSize=_grid(vox, scale),
               CameraEye=e, u_CameraLookAt=look, u_TanCameraFovY=tanf, u_CameraAspectRatio=asp, StepSize=step,
               ApplyOcclusion=1, ApplyShadow=1, ApplyGradientPhongShading=phong)
    _one_pass_phong_uniforms(p, light, e)
    return p


def run_iso(vox, cam, light, prm, W, H, grad=None, scale=(1.0, 1.0, 1.0)):
    """RayCasting1PassIsoAdapt (rc1pisoadaptrenderer.cpp: CreateRenderingPass + Update)."""
    from oracle import bind
    p = Program("iso")
    p.texture("TexVolume", Texture(bind.volume_r16f(vox), 3))
    phong = 1 if (grad is not None and light.apply_phong == 1) else 0
    if phong:
        p.texture("TexVolumeGradient", Texture(grad, 3))
    e, look, tanf, asp = camera_uniforms(cam)
    G = _grid(vox, scale)
    d, h, w = vox.shape
    p.set_many(VolumeGridResolution=np.array([w, h, d], np.float32), VolumeVoxelSize=np.array(scale, np.float32), VolumeGridSize=G, CameraEye=e, u_CameraLookAt=look,
               u_TanCameraFovY=tanf, u_CameraAspectRatio=asp, Isovalue=prm.isovalue, StepSizeSmall=prm.step_size_small,
               StepSizeLarge=prm.step_size_large, StepSizeRange=prm.step_size_range, Color=np.array(list(prm.color), np.float32),
               ApplyGradientPhongShading=phong)
    _one_pass_phong_uniforms(p, light, e)
    return _frame(p, W, H, allowed_unset=("ProjectionMatrix", "VolumeScales", "TexVolumeGradient"))


def run_ebs(vox, tf, sat, cam, light, prm, W, H, grad=None, scale=(1.0, 1.0, 1.0)):
    """RC1PExtinctionBasedShading: CreateRenderingShaders (ebsrenderer.cpp:557-590) + Update (:125-247).
    DirSdwConeSamples = 120 (:36) is uploaded but never read by the shader."""
    return _frame(make_ebs(vox, tf, sat, cam, light, prm, grad, scale), W, H, allowed_unset=("ProjectionMatrix", "TexVolumeGradient"))


def make_ebs(vox, tf, sat, cam, light, prm, grad=None, scale=(1.0, 1.0, 1.0)):
    """The bound program of run_ebs, ready to dispatch (bench.py times the dispatch alone)."""
    p = Program("ebs")
    _volume_and_tf(p, vox, tf)
    phong = 1 if (grad is not None and light.apply_phong == 1) else 0
    if phong:
        p.texture("TexVolumeGradient", Texture(grad, 3))
    p.texture("TexVolumeSAT3D", Texture(sat, 3))
    e, look, tanf, asp = camera_uniforms(cam)
    p.set_many(VolumeScales=np.array(scale, np.float32), VolumeScaledSizes=_grid(vox, scale),
               u_sat_width=sat.shape[2], u_sat_height=sat.shape[1], u_sat_depth=sat.shape[0],
               AmbOccShells=int(prm.amb_occ_shells), AmbOccRadius=prm.amb_occ_radius, DirSdwConeSamples=120,
               DirSdwConeAngle=prm.sdw_cone_angle_rad, DirSdwSampleInterval=prm.sdw_sample_interval, DirSdwInitialStep=prm.sdw_initial_step,
               DirSdwUserInterfaceWeight=prm.sdw_ui_weight, DirSdwConeMaxDistance=prm.sdw_cone_max_distance,
               LightCamForward=_v3(light.light_forward), TypeOfShadow=int(prm.type_of_shadow),
               CameraEye=e, ViewMatrix=look, fov_y_tangent=tanf, aspect_ratio=asp,
               ApplyOcclusion=int(prm.apply_occlusion), ApplyShadow=int(prm.apply_shadow), StepSize=prm.step_size, ApplyPhongShading=phong)
    _lit_uniforms(p, light, e)
    return p


def run_gt_cube(vox_shape, cam, W, H, scale=(1.0, 1.0, 1.0)):
    """RC1PConeLightGroundTruthSteps::RedrawCube (crtgtrenderer.cpp:327-338): vol_intersection.comp with the uniforms of
    CreateRenderingPass (:583-600) and Update (:247-255)."""
    p = Program("gt_volint")
    d, h, w = vox_shape
    e, look, tanf, asp = camera_uniforms(cam)
    p.set_many(VolumeGridSize=np.array([w * scale[0], h * scale[1], d * scale[2]], np.float32), VolumeGridResolution=np.array([w, h, d], np.float32),
               CameraEye=e, CameraLookAt=look, CameraAspectRatio=asp, TanCameraFovY=tanf)
    return _frame(p, W, H, allowed_unset=("CameraProjection",))


def pyramid_levels(pyr, dims):
    """Split oracle.bind.extcoef_build's concatenated pyramid into per-level (d, h, w) arrays."""
    levels, off = [], 0
    for w, h, d in ((int(a), int(b), int(c)) for a, b, c in dims):
        levels.append(pyr[off:off + w * h * d].reshape(d, h, w).copy())
        off += w * h * d
    return levels


def bind_dos_cone(p, prefix, cone):
    """BindConeOcclusionUniforms / BindConeShadowUniforms (dosrcrenderer.cpp:823-985); cone = oracle.bind.OrcDosCone."""
    p.texture(f"Tex{prefix}ConeSectionsInfo", Texture(cone._keep, 1))
    p.set_many(**{f"{prefix}InitialStep": cone.initial_step, f"{prefix}Ray7AdjWeight": cone.ray7_adj_weight,
                  f"{prefix}ConeRayAxes": np.array([[cone.axes[i][j] for j in range(3)] for i in range(10)], np.float32),
                  f"{prefix}ConeIntegrationSamples": np.array(list(cone.counts), np.int32), f"{prefix}UIWeight": cone.ui_weight})


def run_dos(vox, tf, pyr, dims, cam, light, occ, sdw, prm, W, H, grad=None, scale=(1.0, 1.0, 1.0)):
    """RC1PConeTracingDirOcclusionShading: CreateRenderingPass (dosrcrenderer.cpp:659-700) + Update (:134-247)."""
    p = Program("dos")
    _volume_and_tf(p, vox, tf)
    phong = 1 if (grad is not None and light.apply_phong == 1) else 0
    if phong:
        p.texture("TexVolumeGradient", Texture(grad, 3))
    p.texture("TexVolumeOfGaussians", Texture(pyramid_levels(pyr, dims), 3))
    bind_dos_cone(p, "Occ", occ)
    bind_dos_cone(p, "Sdw", sdw)
    e, look, tanf, asp = camera_uniforms(cam)
    p.set_many(VolumeScales=np.array(scale, np.float32), VolumeScaledSizes=_grid(vox, scale),
               SpotLightMaxAngle=prm.spot_cos, TypeOfShadow=int(prm.type_of_shadow),
               LightCamForward=_v3(light.light_forward), LightCamUp=_v3(light.light_up), LightCamRight=_v3(light.light_right),
               CameraEye=e, ViewMatrix=look, fov_y_tangent=tanf, aspect_ratio=asp,
               ApplyOcclusion=int(prm.apply_occlusion), ApplyShadow=int(prm.apply_shadow), Shade=1 if (prm.apply_occlusion or prm.apply_shadow) else 0,
               StepSize=prm.step_size, ApplyPhongShading=phong)
    _lit_uniforms(p, light, e)
    return _frame(p, W, H, allowed_unset=("ProjectionMatrix", "TexVolumeGradient"))


def run_vct(vox, tf, levels, lut, cam, light, prm, W, H, grad=None, scale=(1.0, 1.0, 1.0)):
    """RC1PVoxelConeTracingSGPU: CreateRenderingPass (vctrenderer.cpp:517-560) + Update (:124-237)."""
    p = Program("vct")
    _volume_and_tf(p, vox, tf)
    phong = 1 if (grad is not None and light.apply_phong == 1) else 0
    if phong:
        p.texture("TexVolumeGradient", Texture(grad, 3))
    p.texture("TexSuperVoxelsVolume", Texture(levels, 3))
    p.texture("TexPreIntegrationLookup", Texture(lut, 2))
    e, look, tanf, asp = camera_uniforms(cam)
    p.set_many(VolumeScaledSizes=_grid(vox, scale), VolumeScales=np.array(scale, np.float32),
               TanRadiusConeApexAngle=prm.tan_cone_apex_angle, ConeStepSize=prm.cone_step_size, ConeStepIncreaseRate=prm.cone_step_increase_rate,
               ConeInitialStep=prm.cone_initial_step, OpacityCorrectionFactor=prm.opacity_correction_factor,
               ApplyOpacityCorrectionFactor=int(prm.apply_opacity_correction), ConeNumberOfSamples=int(prm.cone_number_of_samples),
               VolumeMaxDensity=prm.volume_max_density, VolumeMaxStandardDeviation=prm.volume_max_stddev,
               CameraEye=e, ViewMatrix=look, fov_y_tangent=tanf, aspect_ratio=asp,
               ApplyOcclusion=int(prm.apply_occlusion), ApplyShadow=int(prm.apply_shadow), StepSize=prm.step_size, ApplyPhongShading=phong)
    _lit_uniforms(p, light, e)
    return _frame(p, W, H, allowed_unset=("ProjectionMatrix", "TexVolumeGradient"))


def run_gt(vox, tf, cam, light, prm, occ_rays, sdw_rays, W, H, grad=None, max_dispatches=4000, scale=(1.0, 1.0, 1.0)):
    """RC1PConeLightGroundTruthSteps: CreateRenderingPass (crtgtrenderer.cpp:545-600), Update (:189-245), then PreRedraw's
    clear (:262-270) and RedrawFrameTexture's loop (:272-325): one dispatch = one primary sample per pixel, colour kept in
    the rgba16f frame, ray parameter + done flag in an rg16f image, until no pixel is pending.  Returns (frame,
    dispatches, stalled) where stalled counts pixels whose done flag can never be written (see tests/test_refglsl.py)."""
    p = Program("gt")
    _volume_and_tf(p, vox, tf)
    phong = 1 if (grad is not None and light.apply_phong == 1) else 0
    if phong:
        p.texture("TexVolumeGradient", Texture(grad, 3))
    r16 = lambda a: np.ascontiguousarray(np.asarray(a, np.float32).reshape(-1, 3).astype(np.float16).astype(np.float32))
    p.texture("TexOccRaysSampledVectors", Texture(r16(occ_rays), 1))
    p.texture("TexSdwRaysSampledVectors", Texture(r16(sdw_rays), 1))
    e, look, tanf, asp = camera_uniforms(cam)
    d, h, w = vox.shape
    # the aperture angles only feed the host's ray tables
    p.set_many(VolumeGridSize=_grid(vox, scale), VolumeGridResolution=np.array([w, h, d], np.float32),
               CameraEye=e, CameraLookAt=look, CameraAspectRatio=asp, TanCameraFovY=tanf, StepSize=prm.step_size,
               LightRayInitialGap=prm.light_ray_initial_gap, LightRayStepSize=prm.light_ray_step_size,
               ApplyConeOcclusion=int(prm.apply_occlusion), OccNumberOfSampledRays=int(prm.occ_num_rays), OccConeApertureAngle=90.0,
               OccConeDistanceEvaluation=prm.occ_cone_distance,
               ApplyConeShadow=int(prm.apply_shadow), SdwNumberOfSampledRays=int(prm.sdw_num_rays), SdwConeApertureAngle=10.0,
               SdwConeDistanceEvaluation=prm.sdw_cone_distance, SdwShadowType=int(prm.shadow_type),
               ApplyGradientPhongShading=phong, LightSourcePosition=_v3(light.light_pos), LightCamForward=_v3(light.light_forward),
               LightCamUp=_v3(light.light_up), LightCamRight=_v3(light.light_right),
               BlinnPhongKa=light.ka, BlinnPhongKd=light.kd, BlinnPhongKs=light.ks, BlinnPhongShininess=light.shininess)
    out = np.zeros((H, W, 4), np.float32)
    state = np.zeros((H, W, 2), np.float32)
    p.image("OutputFrag", Image(out))
    p.image("StateFrag", Image(state))
    dispatches, stalled = 0, 0
    while True:
        before = (out.copy(), state.copy())
        p.dispatch(W, H)
        dispatches += 1
        pending = state[..., 1] < 0.5
        if not pending.any():                    # RedrawFrameTexture's stop test (:314-322)
            break
        if np.array_equal(before[0], out) and np.array_equal(before[1], state):
            stalled = int(pending.sum())
            break
        assert dispatches < max_dispatches
    assert p.unknown_uniforms() == [], p.unknown_uniforms()
    assert set(p.unset_uniforms()) <= {"CameraProjection", "TexVolumeGradient"}, p.unset_uniforms()
    return out, dispatches, stalled


def run_obj(vox, tf, cam, light, apply_occlusion, apply_shadow, step, cache, W, H, grad=None, scale=(1.0, 1.0, 1.0)):
    """_common_shaders/obj_ray_marching.comp over a light cache [rd, rh, rw, 2], as the DOS / EBS / VCT renderers dispatch it
    while PreIlluminationStructuredVolume is active (e.g. dosrcrenderer.cpp:134-141,659-700).  `Shade` is uploaded by the DOS
    host although this shader does not declare it (glGetUniformLocation == -1: ignored)."""
    p = Program("obj")
    _volume_and_tf(p, vox, tf)
    phong = 1 if (grad is not None and light.apply_phong == 1) else 0
    if phong:
        p.texture("TexVolumeGradient", Texture(grad, 3))
    p.texture("TexVolumeLightCache", Texture(cache, 3))
    e, look, tanf, asp = camera_uniforms(cam)
    p.set_many(VolumeScales=np.array(scale, np.float32), VolumeScaledSizes=_grid(vox, scale), CameraEye=e, ViewMatrix=look, fov_y_tangent=tanf, aspect_ratio=asp,
               ApplyOcclusion=int(apply_occlusion), ApplyShadow=int(apply_shadow), Shade=1, StepSize=step, ApplyPhongShading=phong)
    _lit_uniforms(p, light, e)
    return _frame(p, W, H, allowed_unset=("ProjectionMatrix", "TexVolumeGradient"), allowed_unknown=("Shade",))


# ---- pre-passes, light caches, image filters --------------------------------------------------------------------------
def mip_dims(w, h, d):
    """Level sizes of a complete GL mip chain (max(1, N >> l) down to 1 x 1 x 1)."""
    dims = [(w, h, d)]
    while max(dims[-1]) > 1:
        dims.append(tuple(max(1, v >> 1) for v in dims[-1]))
    return dims


def run_extcoef_pyramid_samesize(vox, tf, sigma0=1.0, scale=(1.0, 1.0, 1.0)):
    """ExtinctionCoefficientVolume::GenerateExtinctionCoefficientVolumeSameSize (extcoefvolumegenerator.cpp:92-228), the
    branch BuildMipMappedTexture takes when no custom resolution is set: gen_extcoefvol_samesize.comp for the base level
    (grid position = (i + 0.5) * VoxelSize, normalised by VolumeResolution * VoxelSize), the same level and back-to-tau
    shaders as the any-size build."""
    return run_extcoef_pyramid(vox, tf, sigma0, None, scale)


def run_extcoef_pyramid(vox, tf, sigma0=1.0, res=(128, 128, 128), scale=(1.0, 1.0, 1.0)):
    """ExtinctionCoefficientVolume::GenerateExtinctionCoefficientVolumeAnySize + TransformTexOpacityToExtinction
    (extcoefvolumegenerator.cpp:230-408) on the reference's three shaders: Gaussian-filtered opacity at the base level,
    every further level filtered from the fp16 level above through textureLod, then -log(1 - a) in place.
    Returns the list of r16f levels [(d, h, w) float32]."""
    from oracle import bind
    G = _grid(vox, scale)
    same_size = res is None
    if same_size:
        res = (vox.shape[2], vox.shape[1], vox.shape[0])
    rw, rh, rd = res
    dims = mip_dims(rw, rh, rd)
    levels = [np.zeros((d, h, w), np.float32) for (w, h, d) in dims]
    images = [Image(l.reshape(l.shape + (1,))) for l in levels]
    tex = Texture(levels, 3)                      # the images alias the texture's levels, as in GL
    base = Program("extcoef_base_same" if same_size else "extcoef_base")
    base.texture("TexInputVolume", Texture(bind.volume_r16f(vox), 3))
    base.texture("TexInputTransferFunc", Texture(tf.texture_rgba(), 1))
    base.image("TexBaseLevelExtCoefVolume", images[0])
    if same_size:
        base.set_many(VolumeResolution=np.array(res, np.float32), VoxelSize=np.array(scale, np.float32), S0=sigma0)
    else:
        base.set_many(ExtCoefVolumeResolution=np.array(res, np.float32), ExtCoefVoxelSize=G / np.array(res, np.float32), S0=sigma0, VolumeGridSize=G)
    base.dispatch(rw, rh, rd, local=(8, 8, 8))
    assert base.unset_uniforms() == [] and base.unknown_uniforms() == []
    lev = Program("extcoef_level_same" if same_size else "extcoef_level")
    lev.set_many(S0=sigma0, VolumeGridSize=G)
    lev.texture("TexExtinctionCoefficientVolume", tex)
    for i, (w, h, d) in enumerate(dims):
        if i == 0:
            continue
        lev.image("TexMipMapLevelExtCoefVolume", images[i])
        lev.set_many(PreviousMipMapLevel=float(i - 1), SubLevelVolumeResolution=np.array([w, h, d], np.float32),
                     Si=np.float32(sigma0) * np.float32(2.0) ** np.float32(i))
        lev.dispatch(w, h, d, local=(8, 8, 8))
    if len(dims) > 1:
        assert lev.unset_uniforms() == [] and lev.unknown_uniforms() == []
    back = Program("extcoef_backtotau")
    for i, (w, h, d) in enumerate(dims):
        back.image("TexExtinctionCoefficientVolume", images[i])
        back.set_many(MMLevelVolResolution=np.array([w, h, d], np.float32), MMLevel=i, S0=sigma0)
        back.dispatch(w, h, d, local=(8, 8, 8))
    assert back.unknown_uniforms() == []
    return levels


def run_sobel(vox):
    """DataManager::GenerateStructuredGradientTexture, compute-shader branch (datamanager.cpp:623-717): three r16f images
    written by sobelfeldman_generator.comp, interleaved into the RGB16F gradient texture [(d, h, w, 3)]."""
    from oracle import bind
    d, h, w = vox.shape
    chans = [np.zeros((d, h, w, 1), np.float32) for _ in range(3)]
    p = Program("sobel")
    p.texture("TexVolume", Texture(bind.volume_r16f(vox), 3))
    p.set("VolumeDimensions", np.array([w, h, d], np.float32))
    for name, c in zip(("TexGradient_RED", "TexGradient_GREEN", "TexGradient_BLUE"), chans):
        p.image(name, Image(c))
    p.dispatch(w, h, d, local=(8, 8, 8))
    assert p.unset_uniforms() == [] and p.unknown_uniforms() == []
    return np.concatenate(chans, axis=-1)


def camera_vectors(eye, center, up):
    """Camera::GetCameraVectors (libs/vis_utils/camera.cpp:336-341): forward = -dir, right = up x forward, up = forward x right."""
    e = np.asarray(eye, np.float32); c = np.asarray(center, np.float32); u = np.asarray(up, np.float32)
    d = c - e
    d = d / np.sqrt(np.sum(d * d, dtype=np.float32))
    f = -d
    r = np.cross(u, f).astype(np.float32); r /= np.sqrt(np.sum(r * r, dtype=np.float32))
    v = np.cross(f, r).astype(np.float32); v /= np.sqrt(np.sum(v * v, dtype=np.float32))
    return f.astype(np.float32), v.astype(np.float32), r.astype(np.float32)


def _light_cache_program(name, vox, tf, res, scale=(1.0, 1.0, 1.0)):
    d, h, w = vox.shape
    p = Program(name)
    cache = np.zeros((res[2], res[1], res[0], 2), np.float32)
    _volume_and_tf(p, vox, tf)                                      # bound by the host although the shaders never read them
    p.image("TexLightCache", Image(cache))
    p.set_many(LightCacheDimensions=np.array(res, np.float32), VolumeDimensions=np.array([w, h, d], np.float32),
               VolumeScales=np.array(scale, np.float32), VolumeScaledSizes=_grid(vox, scale))
    return p, cache


def run_dos_light_cache(vox, tf, pyr, dims, eye, center, up, light, occ, sdw, prm, res, scale=(1.0, 1.0, 1.0)):
    """K6 rc1pdosct/lightcachecomputation.comp as PreComputeLightCache dispatches it (dosrcrenderer.cpp:555-657,700-735)."""
    fwd_v, up_v, right_v = camera_vectors(eye, center, up)
    p, cache = _light_cache_program("dos_lightcache", vox, tf, res, scale)
    p.texture("TexVolumeOfGaussians", Texture(pyramid_levels(pyr, dims), 3))
    bind_dos_cone(p, "Occ", occ)
    bind_dos_cone(p, "Sdw", sdw)
    p.set_many(ApplyOcclusion=int(prm.apply_occlusion), ApplyShadow=int(prm.apply_shadow), WorldEyePos=_v3(eye), WorldLightingPos=_v3(light.light_pos),
               SpotLightMaxAngle=prm.spot_cos, TypeOfShadow=int(prm.type_of_shadow),
               LightCamForward=_v3(light.light_forward), LightCamUp=_v3(light.light_up), LightCamRight=_v3(light.light_right),
               EyeCamForward=fwd_v, EyeCamUp=up_v, EyeCamRight=right_v)
    p.dispatch(res[0], res[1], res[2], local=(8, 8, 8))
    assert p.unknown_uniforms() == [] and p.unset_uniforms() == [], (p.unknown_uniforms(), p.unset_uniforms())
    return cache


def run_ebs_light_cache(vox, tf, sat, eye, light, prm, res, scale=(1.0, 1.0, 1.0)):
    """K9 rc1pextbsd/lightcachecomputation.comp as PreComputeLightCache dispatches it (ebsrenderer.cpp:441-555)."""
    p, cache = _light_cache_program("ebs_lightcache", vox, tf, res, scale)
    p.texture("TexVolumeSAT3D", Texture(sat, 3))
    p.set_many(AmbOccShells=int(prm.amb_occ_shells), AmbOccRadius=prm.amb_occ_radius, DirSdwConeSamples=120, DirSdwConeAngle=prm.sdw_cone_angle_rad,
               DirSdwSampleInterval=prm.sdw_sample_interval, DirSdwInitialStep=prm.sdw_initial_step, DirSdwUserInterfaceWeight=prm.sdw_ui_weight,
               DirSdwConeMaxDistance=prm.sdw_cone_max_distance, LightCamForward=_v3(light.light_forward), TypeOfShadow=int(prm.type_of_shadow),
               WorldEyePos=_v3(eye), WorldLightingPos=_v3(light.light_pos), ApplyOcclusion=int(prm.apply_occlusion), ApplyShadow=int(prm.apply_shadow))
    p.dispatch(res[0], res[1], res[2], local=(8, 8, 8))
    assert set(p.unknown_uniforms()) <= {"TexVolume", "TexTransferFunc"}, p.unknown_uniforms()
    assert set(p.unset_uniforms()) <= {"u_sat_width", "u_sat_height", "u_sat_depth"}, p.unset_uniforms()
    return cache


def run_vct_light_cache(vox, tf, levels, lut, light, prm, res, apex_angle_deg=2.0, scale=(1.0, 1.0, 1.0)):
    """K13 rc1pvctsg/lightcachecomputation.comp as PreComputeLightCache dispatches it (vctrenderer.cpp:393-515)."""
    p, cache = _light_cache_program("vct_lightcache", vox, tf, res, scale)
    p.texture("TexSuperVoxelsVolume", Texture(levels, 3))
    p.texture("TexPreIntegrationLookup", Texture(lut, 2))
    p.set_many(ConeStepSize=prm.cone_step_size, ConeStepIncreaseRate=prm.cone_step_increase_rate, ConeInitialStep=prm.cone_initial_step,
               RadiusConeApexAngle=np.float32(apex_angle_deg) * np.float32(np.pi) / np.float32(180.0), TanRadiusConeApexAngle=prm.tan_cone_apex_angle,
               ApplyOpacityCorrectionFactor=int(prm.apply_opacity_correction), OpacityCorrectionFactor=prm.opacity_correction_factor,
               ConeNumberOfSamples=int(prm.cone_number_of_samples), VolumeMaxDensity=prm.volume_max_density, VolumeMaxStandardDeviation=prm.volume_max_stddev,
               WorldLightingPos=_v3(light.light_pos), ApplyOcclusion=int(prm.apply_occlusion), ApplyShadow=int(prm.apply_shadow))
    p.dispatch(res[0], res[1], res[2], local=(8, 8, 8))
    assert set(p.unknown_uniforms()) <= {"TexVolume", "TexTransferFunc"}, p.unknown_uniforms()
    assert p.unset_uniforms() == [], p.unset_uniforms()
    return cache


FILTER_KERNELS = ["box", "hat", "catmullrom", "mitchell", "cbs", "comoms"]         # vis::IMAGE_FILTER_KERNEL order (defines.h:16-17)


def _digital_filter(kernel_name, img):
    """The two in-place dispatches of renderoutputframe.cpp:388-408 / 483-502 (rows, then columns)."""
    p = Program("ff_digital_" + kernel_name)
    p.image("OutputTex", Image(img))
    h, w = img.shape[:2]
    p.set_many(TexWidth=w, TexHeight=h, FilterDirection=0)
    p.dispatch(h, 1, 1, local=(8, 1, 1))
    p.set("FilterDirection", 1)
    p.dispatch(w, 1, 1, local=(8, 1, 1))
    assert p.unknown_uniforms() == [] and p.unset_uniforms() == []


def run_frame_filter(src, out_w, out_h, pass_id, kernel=1):
    """RenderFrameToScreen's pixel multi-scaling passes on the reference's shaders (renderoutputframe.cpp:265-540):
    pass 1 multisample_filter.comp; 2 the kernel file linked with downscaling_filter.comp (+ digital filter on the result for
    the cardinal kernels); 3 (digital filter on a copy of the rendered frame first, then) upscaling_filter.comp."""
    src = np.ascontiguousarray(src, np.float32)
    out = np.zeros((out_h, out_w, 4), np.float32)
    if pass_id == 1:
        p = Program("ff_multisample")
        p.texture("TexGeneratedFrame", Texture(src, 2))
        p.image("OutputFrag", Image(out))
        p.dispatch(out_w, out_h)
        return out
    name = FILTER_KERNELS[kernel]
    cardinal = name in ("cbs", "comoms")
    direction = "down" if pass_id == 2 else "up"
    work = src.copy()
    if direction == "up" and cardinal:
        _digital_filter(name, work)
    p = Program(f"ff_{direction}_{name}")
    p.texture("TexGeneratedFrame", Texture(work, 2))
    p.image("OutputFrag", Image(out))
    p.set_many(TexGeneratedWidth=src.shape[1], TexGeneratedHeight=src.shape[0], TargetWidth=out_w, TargetHeight=out_h)
    p.dispatch(out_w, out_h)
    assert p.unknown_uniforms() == [] and p.unset_uniforms() == []
    if direction == "down" and cardinal:
        _digital_filter(name, out)
    return out
