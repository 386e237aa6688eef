"""vrb200 -- B200-native (sm_100a) replacement of the GPU hot path of lquatrin/cpp_volume_rendering.

The product is two shared libraries built in-tree:
  libvrb200.so   hand-written CUDA kernels behind the C ABI of include/vrb200.h
  libvrbhost.so  C++ host mirror of the reference plugin API (BaseVolumeRenderer, RenderingManager, DataManager ...)
This Python package only binds them with ctypes (tests, bench.py, multi-GPU plumbing with torch.distributed).
There is no CPU fallback: if the libraries are missing, importing the bindings raises.
"""
from . import capi  # noqa: F401
from .capi import Context, VrbError, lib_path, host_lib_path  # noqa: F401
