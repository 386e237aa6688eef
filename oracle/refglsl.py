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


def run_iso(vox, cam, light, prm, W, H, grad=None):
    """RayCasting1PassIsoAdapt (rc1pisoadaptrenderer.cpp: CreateRenderingPass + Update)."""
    from oracle import bind
    p = Program("iso")
    p.texture("TexVolume", Texture(bind.volume_r16f(vox), 3))
    phong = 1 if (grad is not None and light.apply_phong == 1) else 0
    if phong:
        p.texture("TexVolumeGradient", Texture(grad, 3))
    e, look, tanf, asp = camera_uniforms(cam)
    G = _grid(vox)
    p.set_many(VolumeGridResolution=G, VolumeVoxelSize=np.ones(3, np.float32), VolumeGridSize=G, CameraEye=e, u_CameraLookAt=look,
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


def run_obj(vox, tf, cam, light, apply_occlusion, apply_shadow, step, cache, W, H, grad=None):
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
    p.set_many(VolumeScales=np.ones(3, np.float32), VolumeScaledSizes=_grid(vox), CameraEye=e, ViewMatrix=look, fov_y_tangent=tanf, aspect_ratio=asp,
               ApplyOcclusion=int(apply_occlusion), ApplyShadow=int(apply_shadow), Shade=1, StepSize=step, ApplyPhongShading=phong)
    _lit_uniforms(p, light, e)
    return _frame(p, W, H, allowed_unset=("ProjectionMatrix", "TexVolumeGradient"), allowed_unknown=("Shade",))
