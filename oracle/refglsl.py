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
