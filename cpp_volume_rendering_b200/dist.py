"""Multi-GPU plumbing (one process per GPU, torch.distributed): host-side logic of the two partitions of SURVEY.md 8e.

  sort-first  image tiles round-robin over ranks, volume replicated, partial frames summed to rank 0 (tile sets are
              disjoint, so the sum is exact): tile_owner_map / reduce_frame.
  sort-last   axis-aligned bricks with one ghost layer, front-to-back visibility order from the eye, ordered RGBA
              compositing of the per-brick partial frames: brick_plan / visibility_order / composite_reference and the
              strip exchange all_to_all_strips (NCCL or gloo) used when peers cannot map each other's memory.

Nothing here renders; the kernels are behind include/vrb200.h.  The functions work on numpy arrays / torch tensors so
that the N > 1 logic is testable on CPU with the gloo backend (tests/test_dist_cpu.py).
"""
import numpy as np


# ---- sort-first ------------------------------------------------------------------------------------------------------
def tile_owner_map(width, height, nranks, tile_w=32, tile_h=32):
    """Owner rank of every pixel: tiles numbered row-major, owner = tile % nranks (vrb_owns_pixel in vrb_internal.cuh)."""
    tiles_x = (width + tile_w - 1) // tile_w
    ty, tx = np.meshgrid(np.arange(height) // tile_h, np.arange(width) // tile_w, indexing="ij")
    return ((ty * tiles_x + tx) % max(nranks, 1)).astype(np.int32)


def reduce_frame(frame_tensor, dst=0):
    """Sum of the disjoint partial frames on rank dst (x + 0 is exact, so the result is bit-identical to one GPU)."""
    import torch.distributed as dist
    dist.reduce(frame_tensor, dst=dst, op=dist.ReduceOp.SUM)
    return frame_tensor


# ---- sort-last -------------------------------------------------------------------------------------------------------
def split_counts(n):
    """Brick grid for n ranks: powers of two are split x, then y, then z (8 -> 2x2x2)."""
    g = [1, 1, 1]
    a = 0
    m = n
    while m > 1:
        if m % 2:
            raise ValueError("sort-last brick grid needs a power-of-two rank count, got %d" % n)
        g[a % 3] *= 2
        m //= 2
        a += 1
    return tuple(g)


def brick_plan(dims_xyz, nranks, ghost=1):
    """List (one per rank) of dicts: origin / owned / ghost_lo / ghost_hi in voxels, and the numpy slices (z, y, x) of the
    sub-array to upload.  Owned regions tile the volume exactly; ghost layers exist only on interior faces."""
    g = split_counts(nranks)
    plans = []
    for r in range(nranks):
        idx = (r % g[0], (r // g[0]) % g[1], r // (g[0] * g[1]))
        origin, owned, glo, ghi = [], [], [], []
        for a in range(3):
            n = dims_xyz[a]
            lo = (n * idx[a]) // g[a]
            hi = (n * (idx[a] + 1)) // g[a]
            origin.append(lo); owned.append(hi - lo)
            glo.append(min(ghost, lo)); ghi.append(min(ghost, n - hi))
        sl = tuple(slice(origin[a] - glo[a], origin[a] + owned[a] + ghi[a]) for a in (2, 1, 0))
        plans.append(dict(rank=r, grid_index=idx, origin=tuple(origin), owned=tuple(owned), ghost_lo=tuple(glo), ghost_hi=tuple(ghi),
                          slices_zyx=sl, global_dims=tuple(dims_xyz)))
    return plans


def visibility_order(plans, eye_world, dims_xyz, scale=(1.0, 1.0, 1.0)):
    """Front-to-back order of the bricks for an eye position (world space, volume centred at the origin): along every
    axis bricks are visited from the eye's side outwards; for an axis-aligned brick grid seen from a point this order
    is valid for every ray.  Returns a list of ranks."""
    G = [dims_xyz[a] * scale[a] for a in range(3)]
    e = [eye_world[a] + 0.5 * G[a] for a in range(3)]          # texture space

    def axis_key(p, a):
        lo = p["origin"][a] * scale[a]
        hi = (p["origin"][a] + p["owned"][a]) * scale[a]
        if e[a] < lo:
            return lo - e[a]
        if e[a] > hi:
            return e[a] - hi
        return 0.0                                             # eye inside the slab: that slab first
    # sort by per-axis distance of the slab from the eye; ties keep rank order (stable)
    return [p["rank"] for p in sorted(plans, key=lambda p: (axis_key(p, 2), axis_key(p, 1), axis_key(p, 0)))]


def composite_reference(partials_in_order):
    """numpy restatement of k_composite_ordered: over() in order, 0.99 cut between segments, fp16 rounding of the result."""
    dst = np.zeros_like(partials_in_order[0], dtype=np.float32)
    done = np.zeros(dst.shape[:-1], bool)
    for p in partials_in_order:
        p = p.astype(np.float32)
        nz = (p > 0).any(-1) & ~done
        om = (1.0 - dst[..., 3:4])
        upd = dst + om * p
        dst = np.where(nz[..., None], upd, dst).astype(np.float32)
        done |= dst[..., 3] > 0.99
    with np.errstate(over="ignore"):
        return dst.astype(np.float16).astype(np.float32)


def strip_rows(height, nranks):
    """Rows [r0, r1) of the final image that each rank composites (direct-send: image cut into N strips)."""
    return [((height * r) // nranks, (height * (r + 1)) // nranks) for r in range(nranks)]


def all_to_all_strips(partial_tensor, nranks):
    """Exchange path without peer mappings: rank r receives strip r of every rank's partial frame.
    partial_tensor: [H, W, 4] float32 with H divisible by nranks.  Returns [nranks, H/nranks, W, 4] (index = source rank)."""
    import torch
    import torch.distributed as dist
    H = partial_tensor.shape[0]
    assert H % nranks == 0, "all_to_all_strips needs H divisible by the rank count"
    out = torch.empty_like(partial_tensor)
    dist.all_to_all_single(out, partial_tensor.contiguous())
    return out.view(nranks, H // nranks, *partial_tensor.shape[1:])
