"""CPU tests of the product's host side: the C ABI library loads and exports every symbol of include/vrb200.h, the
C++ host mirror (libvrbhost.so) reproduces the reference's TransferFunction1D / readers / list parsers.  No compute
call is made (no GPU here)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from cpp_volume_rendering_b200 import capi, synth
from oracle import bind

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_c_abi_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "vrb200.h")).read()
    declared = set(re.findall(r"\b(vrb_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    lib = capi.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/vrb200.h but not exported by libvrb200.so"
    assert declared == set(capi.C_ABI), (declared ^ set(capi.C_ABI))
    assert b"sm_100a" in lib.vrb_version()


def test_library_is_sm100a_only(built):
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "--list-elf", capi.lib_path()], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_error_reporting_without_device(built):
    """No GPU in the CPU container: ctx creation must fail with an error code and message, never exit()."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = capi.load()
    h = C.c_void_p()
    rc = lib.vrb_ctx_create(0, C.byref(h))
    assert rc != 0 and len(lib.vrb_last_error()) > 0
    with pytest.raises(capi.VrbError):
        capi.Context(0)
    assert lib.vrb_rc1pass_render(None, None, None) == 1      # VRB_ERR_INVALID


def test_renderer_registry_names(built):
    h = capi.load_host()
    names = []
    i = 0
    while True:
        a = h.vrbh_renderer_name(i, 1)
        if not a:
            break
        names.append(a.decode())
        i += 1
    # abbreviations kept from the reference
    assert "s_1rc" in names and "s_1rc_eb" in names


@pytest.mark.parametrize("name", ["bonsai", "ramp", "sparse", "thin"])
@pytest.mark.parametrize("ext", [0, 1])
def test_host_transfer_function_matches_oracle_and_reference(built, name, ext):
    h = capi.load_host()
    rgb, a = synth.TFS[name]
    rgb = np.ascontiguousarray(rgb); a = np.ascontiguousarray(a)
    tfh = h.vrbh_tf_create(_p(rgb), len(rgb), _p(a), len(a), 255, ext)
    tfo = bind.TF(rgb, a, 255, ext)
    r = bind.ref()
    tfr = r.ref_tf_create(_p(rgb), len(rgb), _p(a), len(a), 255, ext) if r is not None else None
    try:
        assert h.vrbh_tf_size(tfh) == 256
        rgbt = np.zeros((256, 4), np.float32); rgba = np.zeros((256, 4), np.float32)
        assert h.vrbh_tf_textures(tfh, _p(rgbt), _p(rgba), 256) == 256
        assert np.array_equal(rgbt, tfo.floats_rgbt()) and np.array_equal(rgba, tfo.floats_rgba())
        for x in [v / 255.0 for v in range(256)] + [0.3333, 0.77777, 1.0]:
            e = h.vrbh_tf_get_extn(tfh, x)
            assert e == tfo.get_extn(x)
            assert h.vrbh_tf_get_opcn(tfh, x) == tfo.get_opcn(x)
            if tfr is not None:
                assert e == r.ref_tf_get_extn(tfr, x)
        out = (C.c_float * 4)()
        h.vrbh_tf_get(tfh, 300.0, -1.0, out)          # out of range -> 0
        assert list(out) == [0.0, 0.0, 0.0, 0.0]
    finally:
        h.vrbh_tf_destroy(tfh)
        if tfr is not None:
            r.ref_tf_destroy(tfr)


def test_tf1d_reader_all_three_header_types(built, tmp_path):
    h = capi.load_host()
    for maxd, extf, want_n in ((None, None, 256), (63, None, 64), (127, 1, 128)):
        p = tmp_path / f"t_{maxd}_{extf}.tf1d"
        rgb = np.array([[0.1, 0.2, 0.3, 0], [0.9, 0.8, 0.7, want_n - 1]])
        a = np.array([[0.0, 0], [0.5, want_n - 1]])
        synth.write_tf1d(str(p), (rgb, a), maxd, extf)
        tf = h.vrbh_tf_read(str(p).encode())
        assert tf, h.vrbh_last_error()
        assert h.vrbh_tf_size(tf) == want_n
        o = bind.TF(rgb, a, want_n - 1, 1 if extf else 0)
        rgbt = np.zeros((want_n, 4), np.float32); rgba = np.zeros((want_n, 4), np.float32)
        assert h.vrbh_tf_textures(tf, _p(rgbt), _p(rgba), want_n) == want_n
        assert np.array_equal(rgbt, o.floats_rgbt()) and np.array_equal(rgba, o.floats_rgba())
        h.vrbh_tf_destroy(tf)
    assert not h.vrbh_tf_read(str(tmp_path / "missing.tf1d").encode())
    assert b"cannot open" in h.vrbh_last_error()


def _read_volume(h, path):
    v = h.vrbh_volume_read(str(path).encode())
    if not v:
        return None
    dims = (C.c_int * 3)(); sc = (C.c_double * 3)(); bpv = C.c_int(); cs = C.c_ulonglong()
    h.vrbh_volume_info(v, dims, sc, C.byref(bpv), C.byref(cs))
    arr = np.empty((dims[2], dims[1], dims[0]), np.uint8 if bpv.value == 1 else np.uint16)
    h.vrbh_volume_copy(v, _p(arr))
    ns = h.vrbh_volume_normalized_sample(v, dims[0] - 1, dims[1] - 1, dims[2] - 1)
    h.vrbh_volume_destroy(v)
    return arr, tuple(sc), cs.value, ns


def test_volume_readers_raw_syn_pvm(built, tmp_path):
    h = capi.load_host()
    rng = np.random.default_rng(5)
    # .raw, 8 and 16 bit, sizes parsed from the file name from the right (reader.cpp:172-205)
    for dt, b in ((np.uint8, 1), (np.uint16, 2)):
        vox = rng.integers(0, np.iinfo(dt).max + 1, (5, 6, 7)).astype(dt)
        p = tmp_path / f"My.Volume.{b}.7x6x5.raw"
        vox.tofile(p)
        arr, sc, cs, ns = _read_volume(h, p)
        assert np.array_equal(arr, vox) and sc == (1.0, 1.0, 1.0) and cs == int(vox.astype(np.uint64).sum())
        assert ns == float(vox[-1, -1, -1]) / float(np.iinfo(dt).max)
    assert _read_volume(h, tmp_path / "bad.raw") is None
    (tmp_path / "Short.1.4x4x4.raw").write_bytes(b"123")
    assert _read_volume(h, tmp_path / "Short.1.4x4x4.raw") is None and b"shorter" in h.vrbh_last_error()
    # .syn: box records (half-open) and single-voxel records; the buffer is zero-initialised
    p = tmp_path / "boxes.syn"
    synth.write_syn_boxes(str(p), 24, count=9, seed=7)
    with open(p, "a") as f:
        f.write("0 3 4 5 77\n")
    want = synth.volume_boxes(24, 9, 7)
    want[5, 4, 3] = 77
    arr, _, _, _ = _read_volume(h, p)
    assert np.array_equal(arr, want)
    # .pvm: PVM (8 bit), PVM3 with spacing and 16-bit little-endian payload (pvm.cpp:80-109)
    v8 = rng.integers(0, 256, (3, 4, 5)).astype(np.uint8)
    (tmp_path / "a.pvm").write_bytes(b"PVM\n5 4 3\n1\n" + v8.tobytes())
    arr, sc, _, _ = _read_volume(h, tmp_path / "a.pvm")
    assert np.array_equal(arr, v8) and sc == (1.0, 1.0, 1.0)
    v16 = rng.integers(0, 65536, (3, 4, 5)).astype("<u2")
    (tmp_path / "b.pvm").write_bytes(b"PVM3\n5 4 3\n1 1 1.5\n2\n" + v16.tobytes())
    arr, sc, _, _ = _read_volume(h, tmp_path / "b.pvm")
    assert np.array_equal(arr, v16) and sc == (1.0, 1.0, 1.5)
    # DDS-compressed .pvm: tests/test_pvm_dds.py


def test_camera_and_light_list_parsers(built, tmp_path):
    h = capi.load_host()
    cams = tmp_path / "#list_camera_states"
    with open(cams, "w") as f:
        for name, eye, center, up in synth.CAMERA_STATES_256:
            f.write(f"{name}\nARCBALL\n {eye[0]} {eye[1]} {eye[2]}\n {center[0]} {center[1]} {center[2]}\n {up[0]} {up[1]} {up[2]}\n")
    out = np.zeros((16, 9), np.float32)
    n = h.vrbh_read_camera_states(str(cams).encode(), _p(out), 16)
    assert n == len(synth.CAMERA_STATES_256)
    for i, (_, eye, center, up) in enumerate(synth.CAMERA_STATES_256):
        assert np.allclose(out[i], np.array(eye + center + up, np.float32))
    lights = tmp_path / "#list_light_sources"
    lights.write_text("Shadow Comparison Synthetic Bars\n1\n-206.873 -51.0699 557.011\n-0.346883 -0.0856335 0.933991\n"
                      "-0.0298143 0.996327 0.0802758\n0.937434 -0 0.348162\n20\nTwo\n2\n1 2 3\n0 0 1\n0 1 0\n1 0 0\n5\n"
                      "4 5 6\n0 0 -1\n0 1 0\n-1 0 0\n7\n")
    lo = np.zeros((8, 13), np.float32)
    assert h.vrbh_read_light_lists(str(lights).encode(), _p(lo), 8) == 3
    assert np.allclose(lo[0, :3], [-206.873, -51.0699, 557.011]) and lo[0, 12] == 20.0
    assert np.allclose(lo[0, 3:6], [-0.346883, -0.0856335, 0.933991])
    assert np.allclose(lo[2, :3], [4, 5, 6]) and lo[2, 12] == 7.0
    assert h.vrbh_read_camera_states(str(tmp_path / "nope").encode(), _p(out), 16) == -1


def test_camera_block_matches_oracle(built):
    """vrb::lookAt / tan(fovy/2) / aspect of the C++ host equal the oracle's restatement of glm::lookAt bit for bit."""
    for i in range(len(synth.CAMERA_STATES_256)):
        eye, center, up = synth.camera_state(i, 256)
        a = capi.make_camera(eye, center, up, 1920, 1080)
        b = bind.camera(eye, center, up, 1920, 1080)
        assert bytes(a) == bytes(b)
    assert capi.load_host().vrbh_tan_fovy() == np.float32(np.tan(np.pi / 8))


def test_list_parsers_match_the_reference_parsers(built, tmp_path):
    """CameraStateList::ReadCameraStates / LightSourceList::ReadLightSourceLists of the host mirror against the reference's
    own parsers (libs/volvis_utils/camerastatelist.cpp, lightsourcelist.cpp compiled in place into oracle/_ref/libref.so)
    on the reference's own data files and on a synthetic list: identical floats, identical counts."""
    r = bind.ref()
    if r is None or not hasattr(r, "ref_read_camera_states"):
        pytest.skip("oracle/_ref/libref.so with the list parsers is not available")
    h = capi.load_host()
    r.ref_read_camera_states.argtypes = [C.c_char_p, C.c_void_p, C.c_int]
    r.ref_read_light_lists.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    cam_files, light_files = [], []
    data = "/root/reference/data"
    if os.path.isdir(data):
        cam_files.append(os.path.join(data, "#list_camera_states"))
        light_files.append(os.path.join(data, "#list_light_sources"))
    cams = tmp_path / "cams"
    cam_text = "".join(f"{name}\nARCBALL\n {eye[0]} {eye[1]} {eye[2]}\n {center[0]} {center[1]} {center[2]}\n {up[0]} {up[1]} {up[2]}\n"
                       for name, eye, center, up in synth.CAMERA_STATES_256)
    cams.write_text(cam_text.rstrip("\n"))                     # like the reference's data files: no newline after the last state
    cam_files.append(str(cams))
    lights = tmp_path / "lights"
    light_text = ("One\n1\n-206.873 -51.0699 557.011\n-0.346883 -0.0856335 0.933991\n-0.0298143 0.996327 0.0802758\n0.937434 -0 0.348162\n20\n"
                  "Two\n2\n1 2 3\n0 0 1\n0 1 0\n1 0 0\n5\n4 5 6\n0 0 -1\n0 1 0\n-1 0 0\n7.5")
    lights.write_text(light_text)
    light_files.append(str(lights))
    for path in cam_files:
        a = np.zeros((64, 9), np.float32); b = np.zeros((64, 9), np.float32)
        na, nb = h.vrbh_read_camera_states(path.encode(), _p(a), 64), r.ref_read_camera_states(path.encode(), _p(b), 64)
        assert na == nb and na > 0, (path, na, nb)
        assert np.array_equal(a, b), path
    for path in light_files:
        a = np.zeros((64, 13), np.float32); b = np.zeros((64, 13), np.float32)
        nl = C.c_int(0)
        na, nb = h.vrbh_read_light_lists(path.encode(), _p(a), 64), r.ref_read_light_lists(path.encode(), _p(b), 64, C.byref(nl))
        assert na == nb and na > 0 and nl.value > 0, (path, na, nb)
        assert np.array_equal(a, b), path
    # Deliberate deviation: the reference loops `while (!eof())` (camerastatelist.cpp:40), so a file that ends with a newline
    # yields one more, empty state (name "", eye = center = up = 0).  The host mirror skips blank lines instead.
    trailing = tmp_path / "cams_nl"
    trailing.write_text(cam_text)
    a = np.zeros((64, 9), np.float32); b = np.ones((64, 9), np.float32)
    na, nb = h.vrbh_read_camera_states(str(trailing).encode(), _p(a), 64), r.ref_read_camera_states(str(trailing).encode(), _p(b), 64)
    assert nb == na + 1 and np.array_equal(a[:na], b[:na]) and np.all(b[na] == 0)


def test_readers_reject_implausible_sizes(built, tmp_path):
    """Sizes come from file names and headers; they are checked before anything is allocated (found by the sanitizer fuzz
    runs of tools/fuzz: the reference allocates first and dies)."""
    h = capi.load_host()
    (tmp_path / "huge.syn").write_text("16 1600000000 16\n0 1 1 1 5\n")
    assert _read_volume(h, tmp_path / "huge.syn") is None and b"implausible" in h.vrbh_last_error()
    (tmp_path / "Huge.1.99999x99999x99999.raw").write_bytes(b"abc")
    assert _read_volume(h, tmp_path / "Huge.1.99999x99999x99999.raw") is None and b"implausible" in h.vrbh_last_error()
    (tmp_path / "Short.2.300x300x300.raw").write_bytes(b"abc")           # plausible sizes, tiny file: rejected before allocating
    assert _read_volume(h, tmp_path / "Short.2.300x300x300.raw") is None and b"shorter" in h.vrbh_last_error()
    (tmp_path / "huge.pvm").write_bytes(b"PVM\n70000 70000 70000\n1\n" + bytes(16))
    assert _read_volume(h, tmp_path / "huge.pvm") is None and b"implausible" in h.vrbh_last_error()
    (tmp_path / "maxd.tf1d").write_text("linear\n1\n2000000000\n1\n0 0 0 0\n1\n0 0\n")
    assert not h.vrbh_tf_read(str(tmp_path / "maxd.tf1d").encode()) and b"malformed header" in h.vrbh_last_error()
    (tmp_path / "count.tf1d").write_text("linear\n0\n2000000000\n0.1 0.2 0.3 0\n")
    assert not h.vrbh_tf_read(str(tmp_path / "count.tf1d").encode()) and b"malformed" in h.vrbh_last_error()


def test_every_c_abi_entry_point_survives_null_arguments(built):
    """SURVEY.md section 8b, errors: every ABI call returns an int (0 = ok) + vrb_last_error(), never exit() or a crash.
    All 77 entry points are called with NULL pointers and zero scalars in a child process; no GPU is needed for that."""
    code = r'''
import sys, ctypes as C
sys.path.insert(0, %r)
from cpp_volume_rendering_b200 import capi
lib = capi.load()
errors = 0
for name in sorted(capi.C_ABI):
    restype, argtypes = capi.C_ABI[name]
    args = [0 if t in (C.c_int, C.c_uint, C.c_longlong, C.c_ulonglong, C.c_size_t) else 0.0 if t in (C.c_float, C.c_double) else None for t in argtypes]
    r = getattr(lib, name)(*args)
    if restype is C.c_int and r != 0:
        errors += 1
        assert lib.vrb_last_error(), name
print("CALLED", len(capi.C_ABI), "ERRORS", errors)
''' % ROOT
    p = subprocess.run([os.sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, (p.returncode, p.stderr[-1500:])
    called, errors = (int(v) for v in re.search(r"CALLED (\d+) ERRORS (\d+)", p.stdout).groups())
    assert called >= 77 and errors >= 60, p.stdout


def test_headless_driver_help_and_loud_failure(built):
    exe = os.path.join(ROOT, "cpp_volume_rendering_b200", "vrb_headless")
    p = subprocess.run([exe, "--help"], capture_output=True, text=True, timeout=60)
    assert p.returncode == 0 and "usage:" in p.stdout and "--renderer" in p.stdout and "DDS" in p.stdout
    p = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert p.returncode == 2 and "usage:" in p.stderr


def test_step_multiples_exact_predicate(built):
    """When may a sort-last brick jump to s = k0 * step?  Only if the repeated fp32 addition s = s + step produces exactly
    the multiples of the step (vrb_step_multiples_exact_f, used by the brick kernels); checked against the additions."""
    lib = capi.load()
    def walks_exactly(step, length):
        s = np.float32(0.0); k = 0
        step = np.float32(step)
        while s < np.float32(length):
            if s != np.float32(k) * step or np.float64(s) != np.float64(k) * np.float64(step):
                return False
            s = np.float32(s + step); k += 1
        return True
    for step, length, want in [(0.5, 4000.0, True), (0.25, 4000.0, True), (1.0, 7000.0, True), (0.75, 3000.0, True),
                               (0.3, 100.0, False), (1.0 / 3.0, 50.0, False), (0.1, 10.0, False), (0.5, 1.0e8, False),
                               (0.0, 10.0, False), (-0.5, 10.0, False)]:
        got = bool(lib.vrb_step_multiples_exact_f(step, length))
        assert got == want, (step, length, got)
        if got and length <= 8000:
            assert walks_exactly(step, length), (step, length)     # the predicate never says yes when the walk disagrees
