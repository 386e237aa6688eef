"""The opt-in kernels (environment switches kept for A/B runs) against the default ones: every pair must give bit-identical
frames and equal counters, so that the measurements quoted in DESIGN.md compare like with like.
  rc1pass   VRB_RC1_KERNEL=list     persistent march kernel with in-place compositing      vs k_rc1pass
  rc1pextbsd VRB_EBS_KERNEL=deferred / ray   march -> k_ebs_shade -> composite / one thread per ray   vs k_ebs_coop
  rc1pdosct VRB_DOS_KERNEL=ray / compact     round 1's k_dos / M-lane k_dos_compact                  vs the deferred frame
  rc1pcrtgt VRB_GT_KERNEL=ray, VRB_GT_SHADE=entry                                                   vs k_gt_shade (task)
  rc1pvctsg VRB_VCT_KERNEL=ray                                                                      vs k_vct_shade
and the march's own switches (VRB_LIST_SKIP=0, VRB_VOL_QUADS=0, VRB_MARCH_REFILL, VRB_BRICK_JUMP=0 is covered by test_dist)."""
import os

import numpy as np
import pytest

from cpp_volume_rendering_b200 import capi, synth
from oracle import bind

pytestmark = pytest.mark.gpu

N, W, H = 48, 112, 80


def _scene(ctx, dtype=np.uint8, volume="gauss_noise"):
    vox = getattr(synth, "volume_" + volume)(N, dtype) if volume != "boxes" else synth.volume_boxes(N)
    tf = bind.TF(*synth.TF_BONSAI)
    eye, center, up = synth.camera_state(0, N)
    ctx.volume_upload(vox)
    ctx.tf_upload(tf.floats_rgbt(), tf.floats_rgba())
    ctx.frame_resize(W, H)
    cam = capi.make_camera(eye, center, up, W, H)
    light = capi.default_lighting(light_pos=synth.light_position(N), forward=synth.camera_forward(eye, center))
    return vox, tf, cam, light


def _with_env(env, fn):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return fn()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _frame(ctx, render):
    render()
    return ctx.frame_read().copy(), (ctx.last_sample_count, ctx.last_aux_count)


def _same(a, b):
    return np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32)) and a[1] == b[1]


@pytest.fixture()
def ctx(built):
    c = capi.Context(0)
    yield c
    c.close()


def test_rc1pass_persistent_kernel_equals_default(ctx):
    _, _, cam, _ = _scene(ctx)
    ref = _frame(ctx, lambda: ctx.rc1pass_render(cam, 0.5, count_samples=True))
    for env in ({"VRB_RC1_KERNEL": "list"}, {"VRB_RC1_KERNEL": "list", "VRB_LIST_SKIP": "0"}, {"VRB_VOL_QUADS": "0"}):
        c2 = capi.Context(0)                                   # the quad copy is decided per context
        _scene(c2)
        got = _with_env(env, lambda: _frame(c2, lambda: c2.rc1pass_render(cam, 0.5, count_samples=True)))
        c2.close()
        assert _same(got, ref), env


def test_ebs_kernels_agree(ctx):
    vox, tf, cam, light = _scene(ctx)
    ctx.sat_build(tf.ext_lut(1))
    prm = capi.default_ebs_params(float(np.sqrt(3.0) * N))
    prm.count_samples = 1
    ref = _frame(ctx, lambda: ctx.ebs_render(cam, light, prm))
    for env in ({"VRB_EBS_KERNEL": "deferred"}, {"VRB_EBS_KERNEL": "ray"}, {"VRB_EBS_KERNEL": "deferred", "VRB_LIST_SKIP": "0"}):
        got = _with_env(env, lambda: _frame(ctx, lambda: ctx.ebs_render(cam, light, prm)))
        assert np.array_equal(np.nan_to_num(got[0]), np.nan_to_num(ref[0])) and got[1] == ref[1], env


def test_dos_kernels_agree(ctx):
    _, _, cam, light = _scene(ctx, np.uint16)
    diag = float(np.sqrt(3.0) * N)
    ctx.extcoef_build(1.0, (16, 16, 16))
    occ, _, _ = capi.host_cone_sampler(20.0, 1, 0.5 * diag, 0.35)
    sdw, _, _ = capi.host_cone_sampler(0.5, 0, 0.75 * diag, 1.0)
    ctx.dos_set_cones(occ, sdw)
    prm = capi.default_dos_params(0.5, apply_shadow=True)
    prm.count_samples = 1
    ref = _frame(ctx, lambda: ctx.dos_render(cam, light, prm))
    for env in ({"VRB_DOS_KERNEL": "ray"}, {"VRB_DOS_KERNEL": "compact"}, {"VRB_LIST_SKIP": "0"}, {"VRB_LIST_SKIP": "force"}):
        got = _with_env(env, lambda: _frame(ctx, lambda: ctx.dos_render(cam, light, prm)))
        assert _same(got, ref), env


def test_gt_kernels_agree(ctx):
    _, _, cam, light = _scene(ctx, volume="boxes")
    occ_r, sdw_r = capi.host_gt_ray_tables(8, 90.0, 8, 1.0)
    ctx.gt_set_rays(occ_r, sdw_r)
    prm = capi.default_gt_params(float(np.sqrt(3.0) * N), 8, 8)
    prm.count_samples = 1
    ref = _frame(ctx, lambda: ctx.gt_render(cam, light, prm))
    for env in ({"VRB_GT_KERNEL": "ray"}, {"VRB_GT_SHADE": "entry", "VRB_GT_ILP": "1"}, {"VRB_GT_SHADE": "entry", "VRB_GT_ILP": "2"},
                {"VRB_GT_SHADE": "entry", "VRB_GT_ILP": "4"}, {"VRB_LIST_SKIP": "0"}):
        got = _with_env(env, lambda: _frame(ctx, lambda: ctx.gt_render(cam, light, prm)))
        assert np.array_equal(got[0].view(np.uint32), ref[0].view(np.uint32)) and got[1][0] == ref[1][0], env
        assert abs(got[1][1] - ref[1][1]) <= max(16, ref[1][1] // 100000), env      # secondary steps: one ulp of expf can move one step


def test_vct_kernels_agree(ctx):
    vox, _, cam, light = _scene(ctx, np.uint16, "noise")
    ctx.vct_build(capi.host_opacity_by_density(synth.TF_BONSAI, 2))
    _, _, ms = ctx.vct_info()
    prm = capi.default_vct_params(65535.0, ms, 0.5)
    prm.count_samples = 1
    ref = _frame(ctx, lambda: ctx.vct_render(cam, light, prm))
    got = _with_env({"VRB_VCT_KERNEL": "ray"}, lambda: _frame(ctx, lambda: ctx.vct_render(cam, light, prm)))
    assert _same(got, ref)
