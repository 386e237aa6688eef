"""The span jump of the sort-last brick kernels (vrb_brick_span in cpp_volume_rendering_b200/csrc/vrb_internal.cuh, used by
k_rc1pass_brick and k_vct_brick) restated in numpy float32: a brick starts its ray loop at sample k0 and stops after ray
parameter s_end instead of walking the whole ray.  That is only allowed if NO sample outside [k0, s_end] can be owned by the
brick; here the ownership test is computed exactly as the kernels compute it (cell = clamp(floor(q k)), owned iff lo <= cell
< hi) for every sample of random rays through random bricks, and compared with the span."""
import numpy as np

F = np.float32


def _span(w, d, D, step, lo, hi, k):
    """vrb_brick_span, vectorised over rays: (any, k0, s_end)."""
    t0 = np.zeros(len(D), F); t1 = D.copy()
    for a in range(3):
        lw = F((F(lo[a]) - F(1.5)) / k); hw = F((F(hi[a]) + F(1.5)) / k)
        big = np.abs(d[:, a]) > F(1e-20)
        with np.errstate(divide="ignore", invalid="ignore"):
            ta = ((lw - w[:, a]) / d[:, a]).astype(F); tb = ((hw - w[:, a]) / d[:, a]).astype(F)
        t0 = np.where(big, np.maximum(t0, np.minimum(ta, tb)), t0).astype(F)
        t1 = np.where(big, np.minimum(t1, np.maximum(ta, tb)), np.where((w[:, a] < lw) | (w[:, a] > hw), F(-1.0), t1)).astype(F)
    ok = t1 >= t0
    k0 = np.maximum(0, np.floor((t0 / step).astype(F)).astype(np.int64) - 2)
    return ok, k0, (t1 + F(2.0) * step).astype(F)


def test_no_owned_sample_outside_the_span():
    rng = np.random.default_rng(77)
    rays = 20000
    owned_total = 0
    for n, scale, step in ((256, 1.0, 0.5), (200, 0.5, 0.25), (96, 2.0, 1.0)):
        G = F(n * scale); k = F(F(n) / G); step = F(step)
        for _ in range(4):
            lo = rng.integers(0, n // 2, 3); hi = lo + rng.integers(8, n // 2, 3)          # a brick's owned cells
            # rays through the volume: a point on a random face towards a random interior point
            p0 = (rng.random((rays, 3)) * float(G)).astype(F)
            face = rng.integers(0, 3, rays); side = rng.integers(0, 2, rays)
            p0[np.arange(rays), face] = np.where(side == 1, G, F(0.0))
            p1 = (rng.random((rays, 3)) * float(G)).astype(F)
            d = (p1 - p0).astype(np.float64)
            thin = rng.random(rays) < 0.25
            d[thin, rng.integers(0, 3, thin.sum())] *= 1e-7
            D = np.linalg.norm(d, axis=1).astype(F)
            d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(F)
            ok, k0, s_end = _span(p0, d, D, step, lo, hi, k)
            nmax = int(np.ceil(float(D.max()) / float(step))) + 2
            for kk in range(nmax):
                s = F(kk) * step                                  # exact multiples (the jump is only taken when they are)
                live = s < D
                if not live.any():
                    break
                h = np.minimum(step, (D - s).astype(F)).astype(F)
                t = (s + h * F(0.5)).astype(F)
                q = (p0.astype(np.float64) + d.astype(np.float64) * t[:, None].astype(np.float64)).astype(F)   # fmaf(d, t, w) / w + d t: both within an ulp
                cell = np.clip(np.floor((q * k).astype(F)).astype(np.int64), 0, n - 1)
                owned = live & ((cell >= lo) & (cell < hi)).all(1)
                owned_total += int(owned.sum())
                outside = owned & (~ok | (kk < k0) | (s > s_end))
                assert not outside.any(), (n, scale, kk, int(outside.sum()), lo, hi)
    assert owned_total > 100000, owned_total
