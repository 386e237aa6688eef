"""GPU parity tests of the rc1pass marcher: CUDA path (through the C ABI and through the C++ host mirror) against the
CPU oracle on the same seeded inputs.  Tolerance is BASELINE.json's: max abs <= 2/255 and PSNR >= 50 dB on float RGBA."""
import ctypes as C

import numpy as np
import pytest

from cpp_volume_rendering_b200 import capi, synth
from oracle import bind
from conftest import assert_image_parity, hardware_filter_bounds

pytestmark = pytest.mark.gpu


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def ctx(built):
    c = capi.Context(0)
    yield c
    c.close()


def _render(ctx, vox, tf, eye, center, up, W, H, step=0.5, scale=(1.0, 1.0, 1.0), count=True):
    ctx.volume_upload(vox, scale)
    ctx.tf_upload(tf.floats_rgbt(), tf.floats_rgba())
    ctx.frame_resize(W, H)
    ctx.rc1pass_render(capi.make_camera(eye, center, up, W, H), step, count_samples=count)
    return ctx.frame_read()


CASES = [
    # name, volume maker, tf, camera state, W, H, step
    ("gauss64-bonsai", lambda: synth.volume_gauss(64), "bonsai", 0, 160, 160, 0.5),
    ("gauss64-bonsai-cam1", lambda: synth.volume_gauss(64), "bonsai", 1, 160, 120, 0.5),
    ("noise64-ramp", lambda: synth.volume_noise(64), "ramp", 2, 128, 128, 0.5),
    ("boxes64-sparse", lambda: synth.volume_boxes(64), "sparse", 3, 128, 128, 0.3),
    ("gauss48-u16-bonsai", lambda: synth.volume_gauss(48, np.uint16), "bonsai", 4, 128, 96, 0.7),
    ("noise40-u16-thin", lambda: synth.volume_noise(40, np.uint16), "thin", 5, 96, 96, 1.3),
]


@pytest.mark.parametrize("name,mk,tfname,cam_id,W,H,step", CASES, ids=[c[0] for c in CASES])
def test_rc1pass_matches_oracle(ctx, name, mk, tfname, cam_id, W, H, step):
    vox = mk()
    n = vox.shape[0]
    tf = bind.TF(*synth.TFS[tfname])
    eye, center, up = synth.camera_state(cam_id, n)
    img = _render(ctx, vox, tf, eye, center, up, W, H, step)
    ref, ns = bind.rc1pass(vox, tf, bind.camera(eye, center, up, W, H), W, H, step, count=True)
    assert (ns > 0).sum() > 100, "camera misses the volume: useless case"
    assert_image_parity(img, ref, what=name)
    # same rays hit, same number of loop iterations (the ray set-up is bit-identical by construction)
    assert np.array_equal(img[..., 3] > 0, ref[..., 3] > 0) or tfname in ("sparse",)
    assert abs(ctx.last_sample_count - int(ns.sum())) <= max(2, int(ns.sum()) // 100000)


def test_rc1pass_config1_full_size(ctx):
    """BASELINE config 1: 256^3 u8 V-gauss + bonsai TF at 768x768, step 0.5."""
    n, W, H = 256, 768, 768
    vox = synth.volume_gauss(n)
    tf = bind.TF(*synth.TF_BONSAI)
    eye, center, up = synth.camera_state(0, n)
    img = _render(ctx, vox, tf, eye, center, up, W, H, 0.5)
    ref, ns = bind.rc1pass(vox, tf, bind.camera(eye, center, up, W, H), W, H, 0.5, count=True)
    assert_image_parity(img, ref, what="config1")
    assert ctx.last_sample_count == int(ns.sum())
    assert (ns > 0).sum() == 240599        # rays that hit the box at config 1


def test_rc1pass_anisotropic_scale_and_nonsquare(ctx):
    vox = synth.volume_noise(48)[:32, :40, :]
    tf = bind.TF(*synth.TF_RAMP)
    scale = (1.0, 1.5, 2.0)
    eye, center, up = (90.0, 70.0, 160.0), (0, 0, 0), (0, 1, 0)
    W, H = 200, 120
    img = _render(ctx, vox, tf, eye, center, up, W, H, 0.4, scale)
    ref = bind.rc1pass(vox, tf, bind.camera(eye, center, up, W, H), W, H, 0.4, scale)
    assert_image_parity(img, ref, what="anisotropic")


def test_rc1pass_edge_cases(ctx):
    # camera inside the volume (tnear clamps to 0), 1-voxel-thick volume, frame size not a multiple of the 8x8 tile
    vox = synth.volume_noise(32)
    tf = bind.TF(*synth.TF_THIN)
    W, H = 67, 45
    eye, center, up = (3.0, -2.0, 5.0), (0, 0, -40), (0, 1, 0)
    img = _render(ctx, vox, tf, eye, center, up, W, H, 0.5)
    ref = bind.rc1pass(vox, tf, bind.camera(eye, center, up, W, H), W, H, 0.5)
    assert (ref[..., 3] > 0).all()
    assert_image_parity(img, ref, what="inside")
    thin = synth.volume_noise(32)[:1]
    eye = (10.0, 12.0, 60.0)
    img = _render(ctx, thin, tf, eye, (0, 0, 0), up, W, H, 0.25)
    ref = bind.rc1pass(thin, tf, bind.camera(eye, (0, 0, 0), up, W, H), W, H, 0.25)
    assert_image_parity(img, ref, what="slab")
    # camera looking away from the box.  Reference quirk kept: IntersectBox tests tfar > tnear BEFORE tnear is clamped
    # to 0 (ray_bbox_intersection.comp:29-47), so a box entirely behind the eye still "hits" and the shader marches
    # |tfar| through clamp-to-edge texels.  Kernel and oracle must agree on that too.
    img = _render(ctx, vox, tf, (0.0, 0.0, 100.0), (0, 0, 200), up, W, H, 0.5)
    ref = bind.rc1pass(vox, tf, bind.camera((0.0, 0.0, 100.0), (0, 0, 200), up, W, H), W, H, 0.5)
    assert (ref[..., 3] > 0).any()
    assert_image_parity(img, ref, what="box behind the eye")
    # camera looking sideways past the box: nothing hit, frame stays cleared
    img = _render(ctx, vox, tf, (0.0, 0.0, 100.0), (500, 0, 100), up, W, H, 0.5)
    assert np.all(img == 0.0)


def test_rc1pass_large_tf_table_goes_through_global_memory(ctx):
    """.tf1d type 1 with max_density 4095: 4096 texels > the shared-memory staging limit."""
    vox = synth.volume_gauss(40, np.uint16)
    rgb = np.array([[0.0, 0.2, 0.9, 0], [0.9, 0.9, 0.1, 2000], [1.0, 0.1, 0.1, 4095]], np.float64)
    a = np.array([[0.0, 0], [0.0, 500], [0.4, 3000], [0.7, 4095]], np.float64)
    tf = bind.TF(rgb, a, 4095, 0)
    eye, center, up = synth.camera_state(0, 40)
    img = _render(ctx, vox, tf, eye, center, up, 96, 96, 0.5)
    ref = bind.rc1pass(vox, tf, bind.camera(eye, center, up, 96, 96), 96, 96, 0.5)
    assert_image_parity(img, ref, what="tf4096")


def test_rc1pass_sort_first_partition_is_bit_identical(ctx):
    """Rendering the image as N interleaved tile sets and summing them equals the single-context image bit for bit."""
    vox = synth.volume_gauss(48)
    tf = bind.TF(*synth.TF_BONSAI)
    eye, center, up = synth.camera_state(0, 48)
    W, H = 200, 136
    full = _render(ctx, vox, tf, eye, center, up, W, H, 0.5).copy()
    acc = np.zeros_like(full)
    cam = capi.make_camera(eye, center, up, W, H)
    for r in range(3):
        ctx.set_partition(r, 3, 32, 32)
        ctx.rc1pass_render(cam, 0.5)
        part = ctx.frame_read()
        assert np.all((part == 0) | (acc == 0))      # tile sets are disjoint
        acc += part
    ctx.set_partition(0, 1)
    assert np.array_equal(acc, full)


def test_rc1pass_through_cpp_host_mirror(ctx, built, tmp_path):
    """RenderingManager -> RayCasting1Pass (C++ host, same call sequence as the reference) -> C ABI -> kernel."""
    h = capi.load_host()
    n, W, H = 56, 144, 112
    vox = synth.volume_gauss(n)
    assert h.vrbh_init(0) == 0, h.vrbh_last_error()
    try:
        assert h.vrbh_set_volume(_p(vox), n, n, n, 1, 1.0, 1.0, 1.0) == 0, h.vrbh_last_error()
        rgb, a = synth.TF_BONSAI
        assert h.vrbh_set_tf_points(_p(np.ascontiguousarray(rgb)), len(rgb), _p(np.ascontiguousarray(a)), len(a), 255, 0) == 0
        assert h.vrbh_bind_data() == 0, h.vrbh_last_error()
        assert h.vrbh_reshape(W, H) == 0
        assert h.vrbh_set_renderer(b"s_1rc") == 0, h.vrbh_last_error()
        eye, center, up = synth.camera_state(0, n)
        e = np.array(eye, np.float32); c = np.array(center, np.float32); u = np.array(up, np.float32)
        h.vrbh_set_camera(_p(e), _p(c), _p(u))
        assert h.vrbh_display() == 0, h.vrbh_last_error()
        img = np.zeros((H, W, 4), np.float32)
        assert h.vrbh_read_rgba(_p(img), img.size) == 0, h.vrbh_last_error()
        tf = bind.TF(rgb, a)
        ref = bind.rc1pass(vox, tf, bind.camera(eye, center, up, W, H), W, H, 0.5)   # default step 0.5/sqrt(3)*|scale| = 0.5
        assert_image_parity(img, ref, what="host mirror")
        # parameter surface of the reference's FillParameterSpace: StepSize
        assert h.vrbh_set_param(b"StepSize", 1.0) == 0
        assert h.vrbh_display() == 0
        assert h.vrbh_read_rgba(_p(img), img.size) == 0
        ref = bind.rc1pass(vox, tf, bind.camera(eye, center, up, W, H), W, H, 1.0)
        assert_image_parity(img, ref, what="host mirror step 1.0")
        assert h.vrbh_set_param(b"NoSuchParameter", 1.0) != 0
    finally:
        h.vrbh_shutdown()


def test_state_errors_are_reported_not_fatal(built):
    c = capi.Context(0)
    cam = capi.make_camera((0, 0, 100), (0, 0, 0), (0, 1, 0), 32, 32)
    with pytest.raises(capi.VrbError, match="no volume"):
        c.rc1pass_render(cam, 0.5)
    c.volume_upload(synth.volume_gauss(16))
    with pytest.raises(capi.VrbError, match="transfer function"):
        c.rc1pass_render(cam, 0.5)
    tf = bind.TF(*synth.TF_RAMP)
    c.tf_upload(tf.floats_rgbt(), tf.floats_rgba())
    with pytest.raises(capi.VrbError, match="frame"):
        c.rc1pass_render(cam, 0.5)
    c.frame_resize(32, 32)
    with pytest.raises(capi.VrbError, match="step_size"):
        c.rc1pass_render(cam, 0.0)
    c.rc1pass_render(cam, 0.5)
    c.close()


SKIP_CASES = [
    ("gauss64-bonsai", lambda: synth.volume_gauss(64), "bonsai", 0, 160, 160, 0.5, (1.0, 1.0, 1.0)),
    ("gauss96-bonsai-cam4-fine", lambda: synth.volume_gauss(96), "bonsai", 4, 200, 160, 0.11, (1.0, 1.0, 1.0)),
    ("boxes64-sparse", lambda: synth.volume_boxes(64), "sparse", 3, 128, 128, 0.3, (1.0, 1.0, 1.0)),
    ("boxes72-bonsai-aniso", lambda: synth.volume_boxes(72)[:40, :56, :], "bonsai", 1, 150, 130, 0.45, (1.0, 1.7, 0.6)),
    ("noise64-sparse-u16", lambda: synth.volume_noise(64, np.uint16), "sparse", 2, 128, 128, 0.5, (1.0, 1.0, 1.0)),
    ("gauss64-zero", lambda: synth.volume_gauss(64), "zero", 5, 96, 96, 0.5, (1.0, 1.0, 1.0)),
]


@pytest.mark.parametrize("name,mk,tfname,cam_id,W,H,step,scale", SKIP_CASES, ids=[c[0] for c in SKIP_CASES])
def test_rc1pass_empty_space_skipping_preserves_the_result(ctx, name, mk, tfname, cam_id, W, H, step, scale):
    """skip_empty drops fetches only where alpha is exactly 0: pixels and loop counts are BIT-identical."""
    vox = np.ascontiguousarray(mk())
    tf = bind.TF(*synth.TFS[tfname])
    eye, center, up = synth.camera_state(cam_id, max(vox.shape))
    ctx.volume_upload(vox, scale)
    ctx.tf_upload(tf.floats_rgbt(), tf.floats_rgba())
    ctx.frame_resize(W, H)
    cam = capi.make_camera(eye, center, up, W, H)
    ctx.rc1pass_render(cam, step, count_samples=True)
    plain, n_plain = ctx.frame_read(), ctx.last_sample_count
    ctx.rc1pass_render(cam, step, count_samples=True, skip_empty=True)
    skipped, n_skip = ctx.frame_read(), ctx.last_sample_count
    assert n_plain > 1000
    assert n_skip == n_plain
    assert np.array_equal(plain.view(np.uint32), skipped.view(np.uint32))
    # a new transfer function must rebuild the cell flags
    tf2 = bind.TF(*synth.TFS["ramp"])
    ctx.tf_upload(tf2.floats_rgbt(), tf2.floats_rgba())
    ctx.rc1pass_render(cam, step, count_samples=True)
    plain2 = ctx.frame_read()
    ctx.rc1pass_render(cam, step, count_samples=True, skip_empty=True)
    assert np.array_equal(plain2.view(np.uint32), ctx.frame_read().view(np.uint32))


@pytest.mark.parametrize("name,mk,tfname,cam_id,W,H,step", CASES, ids=[c[0] for c in CASES])
def test_rc1pass_hardware_filter_within_tolerance(ctx, name, mk, tfname, cam_id, W, H, step):
    """VRB_FILTER_HARDWARE (texture-unit trilinear, like the reference's GL sampler) against the fp32 oracle."""
    vox = mk()
    n = vox.shape[0]
    tf = bind.TF(*synth.TFS[tfname])
    eye, center, up = synth.camera_state(cam_id, n)
    ctx.set_filter("hardware")
    try:
        img = _render(ctx, vox, tf, eye, center, up, W, H, step)
        n_hw = ctx.last_sample_count
        ctx.rc1pass_render(capi.make_camera(eye, center, up, W, H), step, count_samples=True, skip_empty=True)
        img_skip = ctx.frame_read()
    finally:
        ctx.set_filter("exact")
    ref, ns = bind.rc1pass(vox, tf, bind.camera(eye, center, up, W, H), W, H, step, count=True)
    assert_image_parity(img, ref, what=name + " (hardware filter)", **hardware_filter_bounds(name))
    assert abs(n_hw - int(ns.sum())) <= int(ns.sum()) // 200
    assert_image_parity(img_skip, ref, what=name + " (hardware filter + skipping)", **hardware_filter_bounds(name))


def test_pipelined_frame_read_equals_blocking_read(ctx):
    """vrb_frame_read_rgba32f_async + vrb_frame_read_wait return the same floats as the blocking read, with two reads in
    flight while later frames are being rendered."""
    import torch
    n, W, H = 48, 160, 120
    vox = synth.volume_gauss(n)
    tf = bind.TF(*synth.TF_BONSAI)
    ctx.volume_upload(vox)
    ctx.tf_upload(tf.floats_rgbt(), tf.floats_rgba())
    ctx.frame_resize(W, H)
    cams = [capi.make_camera(*synth.camera_state(i, n), W, H) for i in range(5)]
    want = []
    for cam in cams:
        ctx.rc1pass_render(cam, 0.5)
        want.append(ctx.frame_read().copy())
    bufs = [torch.zeros((H, W, 4), dtype=torch.float32).pin_memory() for _ in range(2)]
    got = []
    for i, cam in enumerate(cams):
        ctx.rc1pass_render(cam, 0.5)
        ctx.frame_read_async(bufs[i & 1].data_ptr())
        ctx.frame_read_wait(1)
        if i >= 1:
            got.append(bufs[(i - 1) & 1].numpy().copy())     # frame i-1 has landed; its buffer is reused at i+1
    ctx.frame_read_wait(0)
    got.append(bufs[(len(cams) - 1) & 1].numpy().copy())
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert np.array_equal(g, w)
    assert float(np.abs(want[0] - want[1]).max()) > 0.01      # the frames do differ
