"""The empty-cell shortcut of k_list_march (cpp_volume_rendering_b200/csrc/march_list.cu) restated in numpy float32 and checked on
random rays: once a sample has been found in an empty 8^3 occupancy cell, the kernel skips the samples with
t < t_safe = min over axes of (b * ia + ca) without looking at the volume.  The claim the kernel's bit-exactness rests on:
EVERY such sample has its padded floor index inside that cell (so its alpha is exactly 0 and dropping its fetches changes
nothing).  Here the positions and indices are computed exactly as the kernel computes them (fp32, one rounding per operation,
fmaf for the index) for rays of every orientation, including nearly axis-parallel ones, and voxel scales other than 1."""
import numpy as np

F = np.float32
INDEX_EPS = F(0.0078125)


def _fmaf(a, b, c):
    """fp32 fused multiply-add: the product of two floats is exact in fp64, the sum rounds once to fp64 (error far below half an
    fp32 ulp of the result at these magnitudes), then to fp32."""
    return (a.astype(np.float64) * b.astype(np.float64) + np.float64(c)).astype(F)


def _index(w, d, t, k, n):
    """vrb_volume_coords' padded floor index of the sample at ray parameter t (lit shaders: product and sum rounded separately)."""
    p = (w + (d * t).astype(F)).astype(F)
    u = _fmaf(p, k, 0.5)
    u = np.minimum(np.maximum(u, F(0.0)), F(n) + F(0.999))
    return np.floor(u).astype(np.int64)


def test_skipped_samples_stay_in_their_cell():
    rng = np.random.default_rng(2024)
    rays = 40000
    checked = 0
    for n, scale in ((512, 1.0), (200, 0.37), (1024, 2.5), (64, 1.0)):
        G = F(n * scale)
        k = F(F(n) / G)
        # ray = entry point on the box + direction; some directions nearly parallel to an axis
        w = (rng.random((rays, 3)) * float(G)).astype(F)
        d = rng.normal(size=(rays, 3))
        tiny = rng.random(rays) < 0.3
        axis = rng.integers(0, 3, rays)
        d[tiny, axis[tiny]] *= 10.0 ** rng.uniform(-9, -2, tiny.sum())
        d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(F)
        step = F(0.5)
        # per-ray constants as the kernel computes them: ia = 1 / (d k), ca = (-0.5 - w k) ia - INDEX_EPS |ia|
        with np.errstate(divide="ignore", invalid="ignore"):
            ia = (F(1.0) / (d * k).astype(F)).astype(F)
            ca = (((F(-0.5) - (w * k).astype(F)).astype(F) * ia).astype(F) - (INDEX_EPS * np.abs(ia)).astype(F)).astype(F)
        ia = np.where(d != 0, ia, F(0.0)).astype(F)
        ca = np.where(d != 0, ca, F(3.0e38)).astype(F)
        # the first sample of the shortcut: any sample of the ray (here: a random multiple of the step)
        k0 = rng.integers(0, 3 * n, rays)
        t_first = (k0.astype(F) * step + step * F(0.5)).astype(F)
        idx0 = np.stack([_index(w[:, a], d[:, a], t_first, k, n) for a in range(3)], 1)
        b = ((idx0 & ~7) + np.where(d > 0, 8, 0)).astype(F)
        t_axis = (b.astype(np.float64) * ia.astype(np.float64) + ca.astype(np.float64)).astype(F)       # fmaf(b, ia, ca)
        t_safe = t_axis.min(1)
        cell0 = idx0 >> 3
        # the following samples, as the loop produces them: s = s + step, t = s + step / 2
        s = (k0.astype(F) * step).astype(F)
        for _ in range(40):
            s = (s + step).astype(F)
            t = (s + step * F(0.5)).astype(F)
            skipped = t < t_safe
            if not skipped.any():
                break
            idx = np.stack([_index(w[:, a], d[:, a], t, k, n) for a in range(3)], 1)
            bad = skipped & ((idx >> 3) != cell0).any(1)
            assert not bad.any(), (n, scale, int(bad.sum()), w[bad][:2], d[bad][:2], t[bad][:2], t_safe[bad][:2])
            checked += int(skipped.sum())
    assert checked > 200000, checked        # the shortcut is actually exercised: most rays skip several samples per cell
