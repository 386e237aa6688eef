"""Parity at BASELINE.json's FULL sizes (the other parity tests use sizes the oracle finishes in a second).
Config 2 is the bench default: rc1pextbsd, 512^3 u8 V-noise, bonsai TF, 1920x1080.  The CPU oracle needs ~1 min for it
(volume synthesis, fp64 SAT, 4.8 M primary samples x 16 SAT box queries on all cores)."""
import numpy as np
import pytest

from cpp_volume_rendering_b200 import capi, synth
from oracle import bind
from conftest import psnr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(built):
    c = capi.Context(0)
    yield c
    c.close()


def test_config2_full_size_ebs_matches_oracle(ctx):
    n, W, H, step = 512, 1920, 1080, 0.5
    vox = synth.volume_noise(n)
    tf = bind.TF(*synth.TF_BONSAI)
    eye, center, up = synth.camera_state(0, n)
    lut = tf.ext_lut(1)
    ctx.volume_upload(vox)
    ctx.tf_upload(tf.floats_rgbt(), tf.floats_rgba())
    ctx.sat_build(lut)
    ctx.frame_resize(W, H)
    prm = capi.default_ebs_params(float(np.sqrt(3.0) * n), step)
    prm.count_samples = 1
    light = capi.default_lighting(light_pos=synth.light_position(n), forward=synth.camera_forward(eye, center))
    ctx.ebs_render(capi.make_camera(eye, center, up, W, H), light, prm)
    img = ctx.frame_read()
    n_gpu = ctx.last_sample_count
    # the GPU SAT (reference-order wavefront build) must be the reference algorithm's SAT bit for bit: the marcher's box
    # queries are differences of these ~1e7-1e8-sized floats and amplify a single differing ulp
    sat_ref = bind.sat_build(vox, lut)
    sat_gpu = ctx.sat_read(vox.shape)
    assert sat_gpu.shape == sat_ref.shape
    neq = int((sat_gpu != sat_ref).sum())
    print(f"SAT texels differing from the reference-order build: {neq} of {sat_ref.size}")
    assert neq == 0
    del sat_gpu
    ref, ns = bind.ebs(vox, tf, sat_ref, bind.camera(eye, center, up, W, H), bind.copy_struct(light, bind.OrcLighting),
                       bind.copy_struct(prm, bind.OrcEbsParams), W, H, count=True)
    assert n_gpu == int(ns.sum())
    # fp32 cancellation in the SAT queries makes exp(-Stau) overflow RGBA16F at some pixels: the reference's own result
    # (SURVEY.md section 8a12).  Same pixels must be non-finite; the finite ones are held to the parity bar.
    fin_ref, fin_img = np.isfinite(ref), np.isfinite(img)
    print(f"non-finite values: oracle {int((~fin_ref).sum())}, gpu {int((~fin_img).sum())}, differing {int((fin_ref != fin_img).sum())}")
    both = fin_ref & fin_img
    d = np.abs(np.where(both, img, 0.0).astype(np.float64) - np.where(both, ref, 0.0).astype(np.float64))
    print(f"finite pixels: max abs err {d.max():.6f}, values over 2/255: {int((d > 2.0 / 255.0).sum())}, PSNR {psnr(np.where(both, img, 0.0), np.where(both, ref, 0.0)):.1f} dB")
    wi = np.unravel_index(int(np.argmax(d)), d.shape)
    print(f"worst value at (row, col, channel) = {wi}: gpu {img[wi[0], wi[1]]}, oracle {ref[wi[0], wi[1]]}")
    assert np.array_equal(fin_ref, fin_img), f"{int((fin_ref != fin_img).sum())} pixels differ in finiteness"
    assert 0 < int((~fin_ref).sum()) < ref.size // 100
    a = np.where(fin_ref, img, 0.0).astype(np.float64); b = np.where(fin_ref, ref, 0.0).astype(np.float64)
    # The same cancellation leaves a few finite but absurd values (|v| up to 1e4, outside the displayable [0,1] range) where
    # exp() of the noise blows up: there one ulp of expf() is worth more than 2/255 in absolute terms.  Displayable range:
    # BASELINE's absolute bar on the clamped image; out-of-range values: relative 1e-2.
    ac, bc = np.clip(a, 0.0, 1.0), np.clip(b, 0.0, 1.0)
    err = float(np.abs(ac - bc).max())
    assert err <= 2.0 / 255.0, err
    assert psnr(ac, bc) >= 50.0
    wild = np.abs(b) > 1.0
    assert int(wild.sum()) < b.size // 100
    if wild.any():
        rel = float((np.abs(a - b)[wild] / np.abs(b)[wild]).max())
        assert rel <= 1e-2, rel
    ctx.volume_upload(synth.volume_gauss(16))           # release the big buffers
