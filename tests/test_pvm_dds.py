"""DDS-compressed .pvm volumes: the product's independent decoder / encoder
(cpp_volume_rendering_b200/host/host_pvm.cpp) against the REFERENCE's own decoder (libs/file_utils/pvm.cpp compiled in
place into oracle/_ref/libref.so, test-only) -- byte for byte, on encoder output and on arbitrary bit streams -- plus
round trips and the reader's error paths.  CPU only."""
import ctypes as C

import numpy as np
import pytest

from cpp_volume_rendering_b200 import capi, synth
from oracle import bind


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _ref():
    r = bind.ref()
    if r is None or not hasattr(r, "ref_dds_read"):
        pytest.skip("oracle/_ref/libref.so with the reference's pvm.cpp is not available")
    r.ref_dds_read.restype = C.c_longlong
    r.ref_dds_read.argtypes = [C.c_char_p, C.c_void_p, C.c_ulonglong]
    r.ref_pvm_read.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_ulonglong]
    return r


def _encode(h, data, skip, strip, version):
    data = np.ascontiguousarray(data, np.uint8)
    need = h.vrbh_dds_encode(_p(data), data.size, skip, strip, version, None, 0)
    out = np.empty(need, np.uint8)
    assert h.vrbh_dds_encode(_p(data), data.size, skip, strip, version, _p(out), need) == need
    return out


def _decode(h, file_image, cap):
    file_image = np.ascontiguousarray(file_image, np.uint8)
    out = np.empty(max(cap, 1), np.uint8)
    n = h.vrbh_dds_decode(_p(file_image), file_image.size, _p(out), out.size)
    assert n >= 0, n
    return out[:n]


def _ref_decode(r, path, cap):
    out = np.empty(max(cap, 1), np.uint8)
    n = r.ref_dds_read(str(path).encode(), _p(out), out.size)
    assert 0 <= n <= cap, n
    return out[:n]


def _payloads(rng):
    g = synth.volume_gauss(20, np.uint8).ravel()
    g16 = synth.volume_gauss(12, np.uint16).astype("<u2").view(np.uint8).ravel()
    return {
        "empty": np.zeros(0, np.uint8),
        "one": np.array([200], np.uint8),
        "zeros": np.zeros(1000, np.uint8),
        "ramp": (np.arange(5000) % 256).astype(np.uint8),
        "noise": rng.integers(0, 256, 4099).astype(np.uint8),
        "gauss8": g,
        "gauss16": g16,
        "steps": np.repeat(rng.integers(0, 256, 40), rng.integers(1, 300, 40)).astype(np.uint8),
    }


@pytest.mark.parametrize("version", [1, 2])
def test_encoder_output_is_read_identically_by_the_reference_decoder(built, tmp_path, version):
    h, r = capi.load_host(), _ref()
    rng = np.random.default_rng(11)
    for name, data in _payloads(rng).items():
        for skip, strip in ((1, 1), (1, 20), (2, 24), (2, 1), (3, 7), (4, 65536), (2, 5000)):
            img = _encode(h, data, skip, strip, version)
            assert bytes(img[:8]) == (b"DDS v3d\n" if version == 1 else b"DDS v3e\n")
            path = tmp_path / f"{name}_{skip}_{strip}.dds"
            img.tofile(path)
            mine = _decode(h, img, data.size + 16)
            assert np.array_equal(mine, data), (name, skip, strip)
            if data.size:            # the reference treats an empty result as an I/O error
                theirs = _ref_decode(r, path, data.size + 16)
                assert np.array_equal(theirs, data), (name, skip, strip)


def test_arbitrary_bit_streams_decode_like_the_reference(built, tmp_path):
    """Decoder equivalence beyond what the encoder emits: random bytes are valid streams (random skip, strip, run
    lengths and widths, wrap-around of the running value, ragged last shuffle block, zero tail past the end)."""
    h, r = capi.load_host(), _ref()
    rng = np.random.default_rng(2024)
    total = 0
    for i in range(60):
        body = rng.integers(0, 256, int(rng.integers(1, 3000))).astype(np.uint8)
        if i % 3 == 0:      # force short strips so that the two-row predictor is exercised early
            body[0] = (body[0] & 0xC0)
            body[1] = 0
            body[2] = (body[2] & 0x3F) | 0x40
        for magic in (b"DDS v3d\n", b"DDS v3e\n"):
            img = np.concatenate([np.frombuffer(magic, np.uint8), body])
            path = tmp_path / "fuzz.dds"
            img.tofile(path)
            cap = body.size * 8 * 13 + 64
            out = np.empty(cap, np.uint8)
            n_ref = r.ref_dds_read(str(path).encode(), _p(out), cap)
            mine = _decode(h, img, cap)
            if n_ref < 0:       # nothing decoded: the reference reports an error, the product an empty result
                assert mine.size == 0
                continue
            assert n_ref == mine.size and np.array_equal(out[:n_ref], mine), (i, magic, n_ref, mine.size)
            total += n_ref
    assert total > 100000


def test_v3e_shuffles_per_block_of_2_pow_24_samples(built, tmp_path):
    """A payload longer than skip * 2^24 bytes: 'DDS v3e' shuffles each block on its own and the last block is ragged."""
    h, r = capi.load_host(), _ref()
    n = 2 * (1 << 24) + 12345
    data = (np.arange(n, dtype=np.uint32) * 2654435761 >> 27).astype(np.uint8)   # cheap, compressible, not periodic in 2
    img = _encode(h, data, 2, 512, 2)
    path = tmp_path / "big.dds"
    img.tofile(path)
    assert np.array_equal(_decode(h, img, n), data)
    assert np.array_equal(_ref_decode(r, path, n), data)
    img1 = _encode(h, data, 2, 512, 1)
    assert not np.array_equal(img1[8:], img[8:])          # whole-stream shuffle differs from the blocked one


def _read_volume(h, path):
    v = h.vrbh_volume_read(str(path).encode())
    if not v:
        return None
    dims = (C.c_int * 3)(); sc = (C.c_double * 3)(); bpv = C.c_int(); cs = C.c_ulonglong()
    h.vrbh_volume_info(v, dims, sc, C.byref(bpv), C.byref(cs))
    arr = np.empty((dims[2], dims[1], dims[0]), np.uint8 if bpv.value == 1 else np.uint16)
    h.vrbh_volume_copy(v, _p(arr))
    h.vrbh_volume_destroy(v)
    return arr, tuple(sc)


@pytest.mark.parametrize("dds_version", [0, 1, 2])
@pytest.mark.parametrize("dtype", [np.uint8, np.uint16])
def test_pvm_files_read_like_the_reference_pvm_class(built, tmp_path, dtype, dds_version):
    h, r = capi.load_host(), _ref()
    rng = np.random.default_rng(3)
    w, hh, d = 21, 13, 9
    vox = (synth.volume_gauss(32, dtype)[:d, :hh, :w].astype(np.int64) + rng.integers(0, 3, (d, hh, w))).astype(dtype)
    scale = np.array([1.0, 0.7, 1.5])
    path = tmp_path / f"v_{dds_version}.pvm"
    assert h.vrbh_pvm_write(str(path).encode(), _p(np.ascontiguousarray(vox)), w, hh, d, vox.itemsize, _p(scale), dds_version) == 0, h.vrbh_last_error()
    if dds_version:
        assert path.read_bytes()[:8] == (b"DDS v3d\n" if dds_version == 1 else b"DDS v3e\n")
        assert path.stat().st_size < vox.nbytes             # it does compress
    arr, sc = _read_volume(h, path)
    assert np.array_equal(arr, vox)
    dims = (C.c_uint * 3)(); rsc = (C.c_double * 3)()
    out = np.empty(vox.shape, dtype)
    comp = r.ref_pvm_read(str(path).encode(), dims, rsc, _p(out), out.nbytes)
    assert comp == vox.itemsize and tuple(dims) == (w, hh, d)
    assert np.array_equal(out, vox)
    assert sc == tuple(rsc) == (1.0, float(np.float32(0.7)), 1.5)    # the scale goes through a float (pvm.cpp:61-66)


def test_pvm_reader_error_paths(built, tmp_path):
    h = capi.load_host()
    cases = {
        "junk_dds.pvm": (b"DDS v3d\nxxxx", b"bad magic"),             # decodes to bytes that are not a PVM header
        "dds_v9.pvm": (b"DDS v9z\nxxxx", b"unknown DDS"),
        "magic.pvm": (b"PVX\n2 2 2\n1\n" + bytes(8), b"bad magic"),
        "dims.pvm": (b"PVM\n2 0 2\n1\n" + bytes(8), b"bad header"),
        "scale.pvm": (b"PVM2\n2 2 2\n1 -1 1\n1\n" + bytes(8), b"bad header"),
        "comp.pvm": (b"PVM\n2 2 2\n3\n" + bytes(24), b"1- and 2-component"),
        "short.pvm": (b"PVM\n2 2 2\n2\n" + bytes(15), b"truncated"),
        "nohdr.pvm": (b"PVM\n2 2 2", b"bad header"),
    }
    for name, (blob, msg) in cases.items():
        (tmp_path / name).write_bytes(blob)
        assert _read_volume(h, tmp_path / name) is None, name
        assert msg in h.vrbh_last_error(), (name, h.vrbh_last_error())
    assert _read_volume(h, tmp_path / "missing.pvm") is None and b"cannot open" in h.vrbh_last_error()
    # '#' comment lines after a plain "PVM" magic are skipped (pvm.cpp:239-241)
    (tmp_path / "c.pvm").write_bytes(b"PVM\n# a comment\n# another\n2 1 1\n1\n\x07\x09")
    arr, sc = _read_volume(h, tmp_path / "c.pvm")
    assert arr.ravel().tolist() == [7, 9] and sc == (1.0, 1.0, 1.0)
