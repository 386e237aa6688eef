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
    """Owner rank of every pixel: tiles numbered row-major with every tile row rotated by three tiles against the row above
    (a rank's tiles lie on diagonals, not in columns), owner = number % nranks (vrb_owns_pixel in vrb_internal.cuh)."""
    tiles_x = (width + tile_w - 1) // tile_w
    ty, tx = np.meshgrid(np.arange(height) // tile_h, np.arange(width) // tile_w, indexing="ij")
    return ((ty * tiles_x + (tx + 3 * ty) % tiles_x) % max(nranks, 1)).astype(np.int32)


def reduce_frame(frame_tensor, dst=0):
    """Sum of the disjoint partial frames on rank dst (x + 0 is exact, so the result is bit-identical to one GPU)."""
    import torch.distributed as dist
    dist.reduce(frame_tensor, dst=dst, op=dist.ReduceOp.SUM)
    return frame_tensor


def slab_bounds(n_slices, nranks):
    """z-slabs of the bordered SAT grid, one per rank: [(z_lo, z_hi), ...] covering [0, n_slices)."""
    return [((n_slices * r) // nranks, (n_slices * (r + 1)) // nranks) for r in range(nranks)]


def _device_tensor(ptr, count, typestr, device):
    import torch

    class _Wrap:
        __cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (ptr, False), "version": 2}
    return torch.as_tensor(_Wrap(), device=device)


def sat_build_sharded(ctx, ext_lut, volume_depth, rank, nranks, device):
    """SURVEY.md section 8e, row "SAT build": every rank scans one z-slab of the bordered (volume_depth + 2)-slice grid, ONE
    all-gather of the slabs' last fp64 planes gives each rank its prefix plane, then the float slabs are exchanged so that
    every rank holds the whole table (the volume is replicated in a sort-first run) and the gather atlas is built.  Needs an
    initialised process group (NCCL).  Returns the slab bounds."""
    import torch
    import torch.distributed as dist
    n_slices = int(volume_depth) + 2
    bounds = slab_bounds(n_slices, nranks)
    z_lo, z_hi = bounds[rank]
    ctx.sat_build_slab(ext_lut, z_lo, z_hi)
    pptr, pcount = ctx.sat_slab_plane()
    mine = _device_tensor(pptr, pcount, "<f8", device)
    planes = [torch.empty_like(mine) for _ in range(nranks)]
    dist.all_gather(planes, mine.clone())
    prefix = None
    if rank > 0:
        prefix = planes[0].clone()
        for r in range(1, rank):
            prefix += planes[r]                               # fp64, in slab order
    torch.cuda.synchronize()                                  # the prefix plane is complete before the context's stream reads it
    ctx.sat_finish_slab(prefix.data_ptr() if prefix is not None else None)
    sptr, (w, h, d) = ctx.sat_device_ptr()
    sat = _device_tensor(sptr, w * h * d, "<f4", device).view(d, h * w)
    for r, (lo, hi) in enumerate(bounds):
        dist.broadcast(sat[lo:hi], src=r)                     # slabs may differ by one slice: one broadcast per slab
    torch.cuda.synchronize()
    ctx.sat_commit()
    return bounds


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


def brick_plan(dims_xyz, nranks, ghost=1, align=1):
    """List (one per rank) of dicts: origin / owned / ghost_lo / ghost_hi in voxels, and the numpy slices (z, y, x) of the
    sub-array to upload.  Owned regions tile the volume exactly; ghost layers exist only on interior faces.  align > 1:
    interior cuts fall on multiples of `align` (the VCT pyramid windows need that, see vct_brick_plan)."""
    g = split_counts(nranks)
    plans = []
    for r in range(nranks):
        idx = (r % g[0], (r // g[0]) % g[1], r // (g[0] * g[1]))
        origin, owned, glo, ghi = [], [], [], []
        for a in range(3):
            n = dims_xyz[a]
            lo = (n * idx[a]) // g[a]
            hi = (n * (idx[a] + 1)) // g[a]
            if align > 1:
                lo = 0 if idx[a] == 0 else min(n, (lo + align // 2) // align * align)
                hi = n if idx[a] == g[a] - 1 else min(n, (hi + align // 2) // align * align)
                if hi <= lo:
                    raise ValueError("axis %d (%d voxels) cannot be cut into %d bricks aligned to %d" % (a, n, g[a], align))
            origin.append(lo); owned.append(hi - lo)
            glo.append(min(ghost, lo)); ghi.append(min(ghost, n - hi))
        sl = tuple(slice(origin[a] - glo[a], origin[a] + owned[a] + ghi[a]) for a in (2, 1, 0))
        plans.append(dict(rank=r, grid_index=idx, origin=tuple(origin), owned=tuple(owned), ghost_lo=tuple(glo), ghost_hi=tuple(ghi),
                          slices_zyx=sl, global_dims=tuple(dims_xyz)))
    return plans


def vct_cone_reach(cone_initial_step, cone_step_size, cone_step_increase_rate, cone_number_of_samples, tan_cone_apex_angle):
    """Farthest cone tap of EvaluationVoxelConeTracing (vct_ray_bbox_marching.comp:97-144) from its sample, in world
    units, and the largest mip level it asks for: tap i sits at apex_i + step_i / 2, lod = log2(2 * x * tan)."""
    apex, step, reach = float(cone_initial_step), float(cone_step_size), 0.0
    for _ in range(int(cone_number_of_samples)):
        reach = apex + 0.5 * step
        apex += step
        step *= float(cone_step_increase_rate)
    lod = float(np.log2(max(2.0 * reach * float(tan_cone_apex_angle), 1e-30))) if reach > 0 else 0.0
    return reach, max(lod, 0.0)


def vct_brick_plan(dims_xyz, nranks, vct_params, scale=(1.0, 1.0, 1.0)):
    """Brick plan for the VCT renderer: (plans, n_levels, halo).  A shaded sample looks `reach` world units towards the
    light and blends the two levels around its lod, so every brick carries levels 0..ceil(lod) (one more when the lod
    is within rounding of an integer, so that fp32 and this estimate cannot disagree on the last level) and a halo of
    reach / min(scale) voxels + the trilinear footprint of the coarsest level (1.5 texels) + 1, rounded up to the
    coarsest level's texel size, which is also the alignment of the cuts."""
    p = vct_params
    reach, lod = vct_cone_reach(p.cone_initial_step, p.cone_step_size, p.cone_step_increase_rate, p.cone_number_of_samples,
                                p.tan_cone_apex_angle)
    top = int(np.ceil(lod))                                  # coarsest level any tap can touch
    if top - lod < 1e-3:
        top += 1
    n_levels = max(top + 1, 2)
    full_levels = 1
    a = [d // 2 for d in dims_xyz]
    while a[0] * a[1] * a[2] >= 1:
        full_levels += 1
        a = [v // 2 for v in a]
    if n_levels > full_levels:
        raise ValueError("the cones reach mip level %d but the volume has only %d levels: render it on one GPU" % (top, full_levels))
    align = 1 << (n_levels - 1)
    need = reach / min(scale) + 1.5 * align + 1.0
    halo = int(np.ceil(need / align)) * align
    plans = brick_plan(dims_xyz, nranks, ghost=halo, align=align)
    return plans, n_levels, halo


def assemble_top_level(parts, dims_zyx):
    """Whole level from the bricks' owned parts: parts = [(array (d,h,w), origin (x,y,z)), ...] -> float64 (D,H,W)."""
    out = np.full(dims_zyx, np.nan, np.float64)
    for a, (ox, oy, oz) in parts:
        d, h, w = a.shape
        out[oz:oz + d, oy:oy + h, ox:ox + w] = a
    if np.isnan(out).any():
        raise ValueError("the bricks' owned parts do not cover the level")
    return out


def level_dims(dims_xyz, level):
    """Resolution of a pyramid level (each level halves with floor, as PreProcessSuperVoxels does)."""
    d = list(dims_xyz)
    for _ in range(level):
        d = [v // 2 for v in d]
    return tuple(d)


def vct_global_max_stddev(ctx, local_maxes, parts, dims_xyz, n_levels):
    """Deviation range of the WHOLE volume's pyramid from per-brick pieces: the maximum over every brick's window levels
    (local_maxes) and over the levels above the windows, which are reduced on `ctx` from the assembled last window level
    (parts: every brick's (owned means, origin) from Context.sv_top_means).  All ranks call this with the same gathered
    inputs and get the same number."""
    w, h, d = level_dims(dims_xyz, n_levels - 1)
    top = assemble_top_level(parts, (d, h, w))
    m = max(local_maxes)
    if (w // 2) * (h // 2) * (d // 2) >= 1:
        m = max(m, ctx.sv_reduce_top(top))
    return m


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
