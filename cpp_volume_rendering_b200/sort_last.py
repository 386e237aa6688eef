"""Sort-last frame loop over N GPUs (one process per GPU): bricks + compositing through CUDA-IPC peer loads.
BASELINE config 5 (2048^3 u16 at 3840x2160: rc1pass + voxel-cone-traced shadows) is `run(..., renderer="vct")`; bench.py
calls it for the `workloads.cfg5` object of its JSON line and tools/sort_last_run.py is the command-line front end.

Every rank owns one brick of a seeded synthetic volume (dist.brick_plan / dist.vct_brick_plan), renders its partial frame
(vrb_rc1pass_render_brick* / vrb_vct_render_brick), publishes the buffers with vrb_ipc_export; every rank then composites
its strip of the image from ALL partial frames with ONE kernel that loads the peers' pixels over NVLink and stores the
strip straight into rank 0's frame (vrb_composite_sum / vrb_composite_ordered with vrb_frame_set_target): two stream-ordered
barriers per frame, no gather.  The reference has no multi-GPU path
 and cannot load this volume at all (libs/volvis_utils/utils.cpp:25-29)."""
import time

import numpy as np

from . import capi, synth
from . import dist as vdist
from .capi import Context


def run(env, n, W, H, dtype="u16", renderer="vct", steps=5, filter_mode="exact", gen="device", ordered=False, check=False,
        volume="noise", tf="bonsai"):
    """env: an object with torch, dist, rank, world, local, stream, max_over_ranks, sum_over_ranks (bench.Env is one).
    Returns the result dict on rank 0."""
    torch, dist, rank, world, local, stream = env.torch, env.dist, env.rank, env.world, env.local, env.stream
    H -= H % world                                            # strips of equal height
    bpv = 1 if dtype == "u8" else 2
    vox = synth.make_volume(volume, dtype, n) if gen == "host" else None
    rgbt, rgba, _ = capi.host_tf_arrays(synth.TFS[tf], bpv)
    eye, center, up = synth.camera_state(0, n)
    cam = capi.make_camera(eye, center, up, W, H)
    vct = renderer == "vct"
    n_levels = halo = 0
    light = prm = None
    if vct:
        opc = capi.host_opacity_by_density(synth.TFS[tf], bpv)
        prm = capi.default_vct_params(255.0 if bpv == 1 else 65535.0, 1.0, 0.5)      # max_stddev filled in after the pre-pass
        light = capi.default_lighting(light_pos=synth.light_position(n))
        plans, n_levels, halo = vdist.vct_brick_plan((n, n, n), world, prm)
    else:
        plans = vdist.brick_plan((n, n, n), world)
    order = vdist.visibility_order(plans, eye, (n, n, n))
    p = plans[rank]
    brick = capi.Brick()
    brick.global_dims[:] = [n, n, n]; brick.origin[:] = list(p["origin"]); brick.owned[:] = list(p["owned"])
    brick.ghost_lo[:] = list(p["ghost_lo"]); brick.ghost_hi[:] = list(p["ghost_hi"])
    ctx = Context(local)
    ctx.set_stream(stream.cuda_stream)
    t0 = time.perf_counter()
    if gen == "host":
        ctx.volume_upload(np.ascontiguousarray(vox[p["slices_zyx"]]))
    else:
        blk = synth.volume_noise_torch(n, p["slices_zyx"], dtype, device=torch.device("cuda", local))
        torch.cuda.synchronize()
        ctx.volume_upload_device(blk.data_ptr(), blk.shape[2], blk.shape[1], blk.shape[0], bpv)
        ctx.synchronize()
        del blk
        torch.cuda.empty_cache()
    upload_s = time.perf_counter() - t0
    ctx.tf_upload(rgbt, rgba); ctx.frame_resize(W, H)
    prepass_ms = None
    if vct:
        # pre-pass: window pyramid per brick, the levels above from the gathered last window level, one LUT for all
        torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
        lmax = ctx.sv_build_brick(brick, n_levels)
        gathered = [None] * world
        dist.all_gather_object(gathered, (lmax, ctx.sv_top_means(brick)))
        gmax = vdist.vct_global_max_stddev(ctx, [g[0] for g in gathered], [g[1] for g in gathered], (n, n, n), n_levels)
        ctx.preint_build(opc, gmax)
        ctx.synchronize(); dist.barrier(); prepass_ms = (time.perf_counter() - t0) * 1e3
        prm.volume_max_stddev = np.float32(gmax)
        ctx.set_filter(filter_mode)
        ctx.vct_render_brick(cam, light, prm, brick, capi.BRICK_ALPHA)   # allocates the partial frame
    else:
        ctx.rc1pass_render_brick(cam, brick, 0.5)             # allocates the partial frame
    ctx.synchronize()
    my_ptr = ctx.partial_device_ptr()
    my_alpha = ctx.brick_alpha_device_ptr()
    handles = [None] * world
    dist.all_gather_object(handles, (ctx.ipc_export(my_ptr), ctx.ipc_export(my_alpha)))
    ptrs = [my_ptr if r == rank else ctx.ipc_import(handles[r][0]) for r in range(world)]
    aptrs = [my_alpha if r == rank else ctx.ipc_import(handles[r][1]) for r in range(world)]
    front = [aptrs[r] for r in order[:order.index(rank)]]
    r0, r1 = vdist.strip_rows(H, world)[rank]
    # the strips are composited STRAIGHT INTO rank 0's frame: every rank's composite kernel loads the partial frames of all
    # bricks (peer loads) and stores its rows through a CUDA-IPC pointer into one of two target buffers on rank 0 (they
    # alternate, so frame i+1 never lands in the buffer frame i is still being read from).  No gather, no staging copy.
    if rank == 0:
        targets = [ctx.frame_extra(0), ctx.frame_extra(1)]
        box = [[ctx.ipc_export(t) for t in targets]]
    else:
        box = [None]
    dist.broadcast_object_list(box, src=0)
    if rank != 0:
        targets = [ctx.ipc_import(h) for h in box[0]]
    token = torch.zeros(1, device="cuda")
    frame_no = [0]

    def barrier():
        dist.all_reduce(token)                                # stream-ordered: later kernels of this rank wait for every rank's earlier ones

    def frame():
        """Two barriers per frame.  Why the buffers are safe without a third: partial (i) is rewritten by the exact pass of
        frame i+1, which sits behind barrier 1 of frame i+1, and every rank reaches that barrier only after its own
        composite (i); the opacity buffer (i) is rewritten by the alpha pass (i+1), behind barrier 2 (i)."""
        ctx.frame_set_target(targets[frame_no[0] & 1])
        frame_no[0] += 1
        if vct:
            if ordered:
                ctx.vct_render_brick(cam, light, prm, brick, capi.BRICK_SEGMENT)
                barrier()
                ctx.composite_ordered([ptrs[r] for r in order], r0, r1 - r0)
            else:
                ctx.vct_render_brick(cam, light, prm, brick, capi.BRICK_ALPHA)
                barrier()
                ctx.vct_render_brick(cam, light, prm, brick, capi.BRICK_EXACT, front)
                barrier()
                ctx.composite_sum(ptrs, r0, r1 - r0)
        elif ordered:
            ctx.rc1pass_render_brick(cam, brick, 0.5)
            barrier()                                         # every partial frame is complete before anyone reads it
            ctx.composite_ordered([ptrs[r] for r in order], r0, r1 - r0)
        else:
            ctx.rc1pass_brick_alpha(cam, brick, 0.5)          # pass 1: opacity of my segment
            barrier()
            ctx.rc1pass_render_brick_exact(cam, brick, front, 0.5)   # pass 2 reads the front bricks' opacity over NVLink
            barrier()
            ctx.composite_sum(ptrs, r0, r1 - r0)

    for _ in range(3):
        frame()
    barrier()
    torch.cuda.synchronize(); dist.barrier()
    l0 = ctx.launches
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        frame()
    barrier()                                                 # the last frame has landed on rank 0
    e1.record(stream)
    torch.cuda.synchronize(); dist.barrier()
    ms = env.max_over_ranks(e0.elapsed_time(e1) / steps)
    launches = ctx.launches - l0
    # e2e: rank 0 additionally reads every assembled frame back as float RGBA into pinned host memory (pipelined: the copy
    # of frame i overlaps frame i+1; every frame lands inside the timed region)
    pinned2 = [torch.empty((H, W, 4), dtype=torch.float32).pin_memory() for _ in range(2)] if rank == 0 else None
    for _ in range(2):
        frame(); barrier()
        if rank == 0:
            ctx.frame_read_into(pinned2[0].data_ptr())
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        frame()
        barrier()
        if rank == 0:
            ctx.frame_read_async(pinned2[i & 1].data_ptr())
            ctx.frame_read_wait(1)
    if rank == 0:
        ctx.frame_read_wait(0)
    torch.cuda.synchronize(); dist.barrier()
    e2e_ms = env.max_over_ranks((time.perf_counter() - t0) * 1e3 / steps)
    # where a frame's time goes, per rank (exact two-pass mode): CUDA events around the phases of three more frames
    phases = None
    if not ordered:
        names = ["opacity_pass", "barrier_1", "shaded_pass", "barrier_2", "composite"]
        acc = np.zeros(len(names))
        for _ in range(3):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
            ctx.frame_set_target(targets[frame_no[0] & 1]); frame_no[0] += 1
            ev[0].record(stream)
            if vct:
                ctx.vct_render_brick(cam, light, prm, brick, capi.BRICK_ALPHA)
            else:
                ctx.rc1pass_brick_alpha(cam, brick, 0.5)
            ev[1].record(stream); barrier(); ev[2].record(stream)
            if vct:
                ctx.vct_render_brick(cam, light, prm, brick, capi.BRICK_EXACT, front)
            else:
                ctx.rc1pass_render_brick_exact(cam, brick, front, 0.5)
            ev[3].record(stream); barrier(); ev[4].record(stream)
            ctx.composite_sum(ptrs, r0, r1 - r0)
            ev[5].record(stream)
            torch.cuda.synchronize()
            acc += np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(len(names))])
        barrier(); torch.cuda.synchronize()
        t = torch.zeros((world, len(names)), dtype=torch.float64, device="cuda")
        t[rank] = torch.from_numpy(acc / 3.0).to(t.device)
        dist.all_reduce(t)
        t = t.cpu().numpy()
        phases = {nm: {"max_ms": float(t[:, i].max()), "mean_ms": float(t[:, i].mean()), "by_rank_ms": [round(float(v), 3) for v in t[:, i]]}
                  for i, nm in enumerate(names)}
    # loop iterations of the whole frame (every brick counts the samples it owns)
    if vct:
        prm.count_samples = 1
        ctx.vct_render_brick(cam, light, prm, brick, capi.BRICK_EXACT, front)
        prm.count_samples = 0
    else:
        ctx.rc1pass_render_brick_exact(cam, brick, front, 0.5, count_samples=True)
    samples, aux = env.sum_over_ranks([ctx.last_sample_count, ctx.last_aux_count])
    result = None
    if rank == 0:
        img = pinned2[(steps - 1) & 1].numpy().copy()
        result = {"sort_last": True, "n_gpus": world, "volume": f"{n}^3 {dtype} V-{volume} ({gen}-generated)", "frame": [W, H],
                  "ms_per_step": ms, "steps": steps, "samples_per_frame": samples, "secondary_units_per_frame": aux,
                  "value": samples / (ms * 1e-3) / 1e9, "unit": "Gsamples/s",
                  "e2e": {"value": samples / (e2e_ms * 1e-3) / 1e9, "unit": "Gsamples/s", "ms_per_step": e2e_ms,
                          "d2h_bytes_per_step": W * H * 16, "h2d_bytes_per_step": 0},
                  "gpu_launches": int(launches), "brick_grid": vdist.split_counts(world), "visibility_order": order,
                  "mode": "ordered-over" if ordered else "exact two-pass", "renderer": renderer, "filter": filter_mode,
                  "upload_s_rank0": upload_s, "checksum": float(np.nan_to_num(img, nan=0.0, posinf=0.0, neginf=0.0).sum())}
        if phases:
            result["phases"] = phases
        if vct:
            result.update(pyramid_levels_per_brick=n_levels, halo_voxels=halo, window=[int(s.stop - s.start) for s in p["slices_zyx"]][::-1],
                          prepass_ms=prepass_ms, max_stddev=float(prm.volume_max_stddev))
        if check:
            full = Context(local)
            if gen == "host":
                full.volume_upload(vox)
            else:
                fv = synth.volume_noise_torch(n, None, dtype, device=torch.device("cuda", local))
                torch.cuda.synchronize()
                full.volume_upload_device(fv.data_ptr(), n, n, n, bpv); full.synchronize(); del fv
            full.tf_upload(rgbt, rgba); full.frame_resize(W, H)
            if vct:
                full.vct_build(opc)
                assert np.float32(full.vct_info()[2]) == np.float32(prm.volume_max_stddev), (full.vct_info()[2], prm.volume_max_stddev)
                prm.count_samples = 1
                full.set_filter(filter_mode)
                full.vct_render(cam, light, prm)
            else:
                full.rc1pass_render(cam, 0.5, count_samples=True)
            want = full.frame_read()
            err = float(np.abs(img - want).max())
            mse = float(np.mean((img.astype(np.float64) - want) ** 2))
            result.update(max_abs_err=err, psnr_db=(float("inf") if mse == 0 else float(10 * np.log10(1.0 / mse))),
                          parity_ok=bool(err <= 2.0 / 255.0), samples_per_frame_single_gpu=full.last_sample_count)
            full.close()
    torch.cuda.synchronize(); dist.barrier()
    ctx.frame_set_target(None)
    for r in range(world):
        if r != rank:
            ctx.ipc_close(ptrs[r]); ctx.ipc_close(aptrs[r])
    if rank != 0:
        for t in targets:
            ctx.ipc_close(t)
    dist.barrier()
    ctx.close()
    return result
