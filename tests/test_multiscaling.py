"""Pixel multi-scaling of BaseVolumeRenderer / vis::RenderFrameToScreen:
MULTIPLE_RAYS_PER_PIXEL (2x2 rays + multisample_filter.comp), DOWN_SCALING_RENDER (downscaling_filter.comp) and
UP_SCALING_RENDER (upscaling_filter.comp) with the six reconstruction kernels and the recursive digital filters of the
two cardinal ones.  CPU: known answers of the oracle.  GPU: vrb_frame_filter against the oracle, bit for bit, on frames
the marchers rendered; the C++ host mirror's MultiSampleRedraw / DownScalingRedraw / UpScalingRedraw."""
import ctypes as C

import numpy as np
import pytest

from cpp_volume_rendering_b200 import capi, synth
from oracle import bind
from conftest import assert_image_parity

KERNELS = ["box", "hat", "catmull-rom", "mitchell-netravali", "cardinal-bspline3", "cardinal-omoms3"]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _rand_frame(h, w, seed=3):
    rng = np.random.default_rng(seed)
    return rng.random((h, w, 4)).astype(np.float16).astype(np.float32)


# ---------------------------------------------------------------------------------------------------------------- CPU
def test_oracle_multisample_is_the_mean_of_the_2x2_block():
    src = _rand_frame(32, 48)
    out = bind.frame_filter(src, 24, 16, 1)
    want = src.reshape(16, 2, 24, 2, 4).mean((1, 3))
    assert np.abs(out - want).max() <= 2.0 ** -11                      # one fp16 rounding of a value below 1
    # 3x3: the output centre falls on a texel centre, GL_LINEAR returns that texel
    src3 = _rand_frame(27, 30)
    assert np.array_equal(bind.frame_filter(src3, 10, 9, 1), src3[1::3, 1::3])


@pytest.mark.parametrize("k", range(6), ids=KERNELS)
def test_oracle_kernels_reproduce_a_constant_image(k):
    c = np.full((32, 48, 4), 0.75, np.float32)
    down = bind.frame_filter(c, 24, 16, 2, k)
    up = bind.frame_filter(c[:16, :24], 48, 32, 3, k)
    tol = 0.0 if k < 4 else 0.01                                        # cardinal kernels: kernel sum 6 x digital filter ~ 1/6, fp16 steps
    assert np.abs(down[4:-4, 4:-4] - 0.75).max() <= tol + 1e-3
    assert np.abs(up[8:-8, 8:-8] - 0.75).max() <= tol + 1e-3
    # texels outside the image count as zero: the border darkens for every kernel wider than a pixel
    if k >= 1:
        assert up[0, 0, 0] < 0.75


def test_oracle_box_downscale_is_block_mean_and_hat_upscale_is_bilinear():
    src = _rand_frame(20, 28, 5)
    down = bind.frame_filter(src, 14, 10, 2, 0)
    assert np.abs(down - src.reshape(10, 2, 14, 2, 4).mean((1, 3))).max() <= 2.0 ** -11
    up = bind.frame_filter(src, 56, 40, 3, 1)
    # interior output pixel (2i+1, 2j+1) sits a quarter texel past texel (i, j): weights 0.75 / 0.25 per axis
    i, j = 4, 6
    want = (0.75 * 0.75 * src[i, j] + 0.75 * 0.25 * src[i, j + 1] + 0.25 * 0.75 * src[i + 1, j] + 0.25 * 0.25 * src[i + 1, j + 1])
    assert np.abs(up[2 * i + 1, 2 * j + 1] - want).max() <= 2.0 ** -10


def test_oracle_interpolating_kernels_keep_the_samples_when_the_size_does_not_change():
    src = _rand_frame(12, 16, 9)
    for k in (0, 1, 2):                                                 # box, hat, Catmull-Rom interpolate
        assert np.array_equal(bind.frame_filter(src, 16, 12, 3, k)[3:-3, 3:-3], src[3:-3, 3:-3]), KERNELS[k]


# ---------------------------------------------------------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def scene(built):
    n = 32
    vox = synth.volume_gauss(n)
    tf = bind.TF(*synth.TFS["bonsai"])
    eye, center, up = synth.camera_state(1, n)
    c = capi.Context(0)
    c.volume_upload(vox); c.tf_upload(tf.floats_rgbt(), tf.floats_rgba())
    yield c, vox, tf, (eye, center, up)
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("screen", [(96, 64), (75, 53)], ids=["96x64", "75x53"])
def test_multisample_pass_matches_oracle(scene, screen):
    ctx, vox, tf, (eye, center, up) = scene
    W, H = screen
    ctx.frame_resize_multiscaling(W, H, 2, 2)
    assert (ctx.width, ctx.height) == (2 * W, 2 * H)
    ctx.rc1pass_render(capi.make_camera(eye, center, up, W, H), 0.5)    # aspect of the SCREEN, rays of the 2x frame
    hi = ctx.frame_read().copy()
    assert hi.shape == (2 * H, 2 * W, 4) and hi[..., 3].max() > 0.5
    ctx.frame_filter(capi.FILTER_PASS_MULTISAMPLE)
    got = ctx.filtered_frame_read()
    assert got.shape == (H, W, 4)
    assert np.array_equal(got, bind.frame_filter(hi, W, H, 1))
    # the supersampled image is the one-ray image up to edge antialiasing
    ctx.frame_resize(W, H)
    ctx.rc1pass_render(capi.make_camera(eye, center, up, W, H), 0.5)
    one = ctx.frame_read()
    assert np.abs(got - one).mean() < 0.01 and not np.array_equal(got, one)
    with pytest.raises(capi.VrbError, match="no multi-scaling frames"):
        ctx.frame_filter(capi.FILTER_PASS_MULTISAMPLE)                  # vrb_frame_resize dropped the filtered frame


@pytest.mark.gpu
@pytest.mark.parametrize("k", range(6), ids=KERNELS)
@pytest.mark.parametrize("screen", [(96, 64), (75, 53)], ids=["96x64", "75x53"])
def test_downscale_and_upscale_passes_match_oracle(scene, screen, k):
    ctx, vox, tf, (eye, center, up) = scene
    W, H = screen
    cam = capi.make_camera(eye, center, up, W, H)
    # DOWN_SCALING_RENDER: 2x frame filtered down to the screen
    ctx.frame_resize_multiscaling(W, H, 2, 2)
    ctx.rc1pass_render(cam, 0.5)
    hi = ctx.frame_read().copy()
    ctx.frame_filter(capi.FILTER_PASS_DOWNSCALE, k)
    got = ctx.filtered_frame_read()
    want = bind.frame_filter(hi, W, H, 2, k)
    assert np.array_equal(got, want), (KERNELS[k], float(np.abs(got - want).max()))
    assert np.array_equal(ctx.frame_read(), hi)                         # the rendered frame is untouched
    # UP_SCALING_RENDER: half-resolution frame (integer division) filtered up to the screen
    ctx.frame_resize_multiscaling(W, H, -2, -2)
    assert (ctx.width, ctx.height) == (W // 2, H // 2)
    ctx.rc1pass_render(cam, 0.5)
    lo = ctx.frame_read().copy()
    ctx.frame_filter(capi.FILTER_PASS_UPSCALE, k)
    got = ctx.filtered_frame_read()
    want = bind.frame_filter(lo, W, H, 3, k)
    assert got.shape == (H, W, 4)
    assert np.array_equal(got, want), (KERNELS[k], float(np.abs(got - want).max()))
    if k >= 4:
        assert not np.array_equal(ctx.frame_read(), lo)                 # the digital pre-filter ran in place (renderoutputframe.cpp:472-502)
    else:
        assert np.array_equal(ctx.frame_read(), lo)


@pytest.mark.gpu
def test_multiscaling_argument_errors(scene):
    ctx = scene[0]
    with pytest.raises(capi.VrbError, match="bad multipliers"):
        ctx.frame_resize_multiscaling(64, 64, 0, 2)
    with pytest.raises(capi.VrbError, match="rendered frame would be"):
        ctx.frame_resize_multiscaling(3, 3, -4, -4)
    ctx.frame_resize_multiscaling(64, 48, 2, 2)
    with pytest.raises(capi.VrbError, match="kernel"):
        ctx.frame_filter(capi.FILTER_PASS_DOWNSCALE, 6)
    with pytest.raises(capi.VrbError, match="pass"):
        ctx.frame_filter(0, 1)
    ctx.frame_resize(64, 48)


@pytest.mark.gpu
def test_multiscaling_modes_through_cpp_host_mirror(built):
    """SetCurrentMultiScalingMode + Reshape + MultiSampleRedraw / DownScalingRedraw / UpScalingRedraw (volrenderbase.cpp:42-68,
    121-197; rc1prenderer.cpp:153-190), driven like the radio buttons of AddImGuiMultiSampleOptions."""
    h = capi.load_host()
    n, W, H = 32, 88, 60
    vox = synth.volume_gauss(n)
    tf = bind.TF(*synth.TFS["bonsai"])
    rgb, a = synth.TF_BONSAI
    eye, center, up = synth.camera_state(1, n)
    assert h.vrbh_init(0) == 0, h.vrbh_last_error()
    try:
        assert h.vrbh_set_volume(_p(vox), n, n, n, 1, C.c_double(1.0), C.c_double(1.0), C.c_double(1.0)) == 0, h.vrbh_last_error()
        assert h.vrbh_set_tf_points(_p(np.ascontiguousarray(rgb)), len(rgb), _p(np.ascontiguousarray(a)), len(a), 255, 0) == 0
        assert h.vrbh_bind_data() == 0, h.vrbh_last_error()
        assert h.vrbh_reshape(W, H) == 0
        e = np.array(eye, np.float32); c = np.array(center, np.float32); u = np.array(up, np.float32)
        h.vrbh_set_camera(_p(e), _p(c), _p(u))
        assert h.vrbh_set_renderer(b"s_1rc") == 0, h.vrbh_last_error()
        img = np.zeros((H, W, 4), np.float32)

        def frame():
            assert h.vrbh_display() == 0, h.vrbh_last_error()
            assert h.vrbh_read_rgba(_p(img), img.size) == 0, h.vrbh_last_error()
            return img.copy()
        one = frame()
        ocam = bind.camera(eye, center, up, W, H)                      # aspect of the screen in every mode
        step = float(np.float32(0.5))
        hi = bind.rc1pass(vox, tf, ocam, 2 * W, 2 * H, step=step)
        lo = bind.rc1pass(vox, tf, ocam, W // 2, H // 2, step=step)
        assert h.vrbh_set_param(b"MultiScalingMode", C.c_double(1.0)) == 0, h.vrbh_last_error()
        assert_image_parity(frame(), bind.frame_filter(hi, W, H, 1), what="MULTIPLE_RAYS_PER_PIXEL")
        assert h.vrbh_set_param(b"MultiScalingMode", C.c_double(2.0)) == 0
        assert_image_parity(frame(), bind.frame_filter(hi, W, H, 2, 1), what="DOWN_SCALING_RENDER, hat (the default kernel)")
        assert h.vrbh_set_param(b"ImageKernelFilter", C.c_double(3.0)) == 0
        assert_image_parity(frame(), bind.frame_filter(hi, W, H, 2, 3), what="DOWN_SCALING_RENDER, Mitchell-Netravali")
        assert h.vrbh_set_param(b"MultiScalingMode", C.c_double(3.0)) == 0
        assert_image_parity(frame(), bind.frame_filter(lo, W, H, 3, 3), what="UP_SCALING_RENDER, Mitchell-Netravali")
        # a window resize keeps the mode (Reshape, volrenderbase.cpp:42-62)
        assert h.vrbh_reshape(W + 8, H + 4) == 0
        img = np.zeros((H + 4, W + 8, 4), np.float32)
        ocam2 = bind.camera(eye, center, up, W + 8, H + 4)
        lo2 = bind.rc1pass(vox, tf, ocam2, (W + 8) // 2, (H + 4) // 2, step=step)
        assert_image_parity(frame(), bind.frame_filter(lo2, W + 8, H + 4, 3, 3), what="UP_SCALING_RENDER after Reshape")
        assert h.vrbh_reshape(W, H) == 0
        img = np.zeros((H, W, 4), np.float32)
        assert h.vrbh_set_param(b"MultiScalingMode", C.c_double(0.0)) == 0
        assert np.array_equal(frame(), one)
    finally:
        h.vrbh_shutdown()
