#!/usr/bin/env python
"""bench.py -- ray samples/s (Gsamples/s) and ms/frame of the B200 volume marcher.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--impl vrb|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one frame of the workload: Update(camera) + Redraw of the renderer through the C ABI, volume / transfer
function / SAT resident in HBM (they are built once in Init(), like the reference does).  Default workload = config 2
of BASELINE.json: extinction-based shading (rc1pextbsd) of a 512^3 uint8 volume at 1920x1080.

  value     Gsamples/s = primary-ray loop iterations the single-GPU algorithm executes per frame (counted by the
            kernel's own counter, equal to the oracle's count) x K / device time, whole job over all N GPUs.
  e2e       same metric, each step additionally reading the float RGBA frame back into pinned host memory through
            vrb_frame_read_rgba32f (the reference's glGetTexImage(GL_RGBA, GL_FLOAT)); camera uniforms are the H2D.
  roofline  dominant kernel (the marcher): algorithmic L1 bytes / CUDA-event duration, against
            the L1 bandwidth measured in this run; roofline_hbm: unique bytes / duration against MEASURED_PEAKS.json.
  cpu_baseline  the reference's own GLSL marcher compiled for the CPU (oracle/_ref/librefglsl.so, kind "reference"; the
            OpenMP oracle, kind "port", when that library is absent) on a bounded sample of the same workload, all host
            threads, rank 0, N=1.
  --impl reference  the reference arm: that CPU execution of the reference's shader + the reference's own
            SummedAreaTable3D for the SAT, on the host cores, same metric/config.

N > 1: sort-first image tiles (volume replicated), partial frames summed to rank 0 with one NCCL reduce per frame
inside the timed region ("scaling": "strong": the frame is fixed, ranks split its tiles).
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: renderer, volume, dtype, n, W, H, tf, camera state
    "cfg1": dict(renderer="rc1pass", volume="gauss", dtype="u8", n=256, W=768, H=768, tf="bonsai", cam=0,
                 desc="rc1pass 256^3 u8 V-gauss + bonsai_01.tf1d @768x768 step 0.5"),
    "cfg2": dict(renderer="ebs", volume="noise", dtype="u8", n=512, W=1920, H=1080, tf="bonsai", cam=0,
                 desc="rc1pextbsd (15 AO shells + 1deg cone shadows) 512^3 u8 V-noise + bonsai_01.tf1d @1920x1080 step 0.5"),
    "cfg3": dict(renderer="dos", volume="gauss_noise", dtype="u16", n=512, W=1920, H=1080, tf="bonsai", cam=0,
                 desc="rc1pdosct (AO 20deg/3 rays + cone shadow 0.5deg, 128^3 extinction pyramid) 512^3 u16 V-gauss+noise @1920x1080"),
    "cfg4": dict(renderer="gt", volume="boxes", dtype="u8", n=256, W=1920, H=1080, tf="ramp", cam=0, rays=64,
                 desc="rc1pcrtgt cone ground truth, 64 occlusion + 64 shadow rays per sample, 256^3 u8 V-boxes @1920x1080"),
    "cfg5-1gpu": dict(renderer="vct", volume="noise", dtype="u16", n=512, W=1920, H=1080, tf="bonsai", cam=0,
                      desc="rc1pvctsg voxel-cone-traced shadows 512^3 u16 V-noise @1920x1080 (single-GPU stand-in for config 5)"),
    "cfg2-small": dict(renderer="ebs", volume="noise", dtype="u8", n=128, W=480, H=270, tf="bonsai", cam=0,
                       desc="reduced config 2 for quick checks (NOT a bench line)"),
}


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def make_volume(wl):
    from cpp_volume_rendering_b200 import synth
    dt = np.uint8 if wl["dtype"] == "u8" else np.uint16
    n = wl["n"]
    if wl["volume"] == "gauss":
        return synth.volume_gauss(n, dt)
    if wl["volume"] == "noise":
        return synth.volume_noise(n, dt)
    if wl["volume"] == "gauss_noise":
        return synth.volume_gauss_noise(n, dt)
    if wl["volume"] == "boxes":
        return synth.volume_boxes(n)
    raise ValueError(wl["volume"])


def host_tf_arrays(tfname, bpv):
    """TF textures + per-voxel-value extinction LUT from the C++ host mirror (product code, not the oracle)."""
    from cpp_volume_rendering_b200 import capi, synth
    h = capi.load_host()
    rgb, a = synth.TFS[tfname]
    rgb = np.ascontiguousarray(rgb); a = np.ascontiguousarray(a)
    tf = h.vrbh_tf_create(_p(rgb), len(rgb), _p(a), len(a), 255, 0)
    rgbt = np.zeros((256, 4), np.float32); rgba = np.zeros((256, 4), np.float32)
    assert h.vrbh_tf_textures(tf, _p(rgbt), _p(rgba), 256) == 256
    nv = 256 if bpv == 1 else 65536
    mx = 255.0 if bpv == 1 else 65535.0
    lut = np.array([h.vrbh_tf_get_extn(tf, v / mx) for v in range(nv)], np.float32)
    h.vrbh_tf_destroy(tf)
    return rgbt, rgba, lut


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons with NVML during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = False
        self.sm = []
        self.reasons = set()
        self.max_sm = None
        self.power = []

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            hdl = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(hdl, pynvml.NVML_CLOCK_SM)
            names = {pynvml.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                     pynvml.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     pynvml.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     pynvml.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                     pynvml.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
            while not self.stop_flag:
                self.sm.append(pynvml.nvmlDeviceGetClockInfo(hdl, pynvml.NVML_CLOCK_SM))
                try:
                    self.power.append(pynvml.nvmlDeviceGetPowerUsage(hdl) / 1000.0)
                except Exception:
                    pass
                r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(hdl)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
                time.sleep(0.05)
        except Exception as e:  # NVML missing: report nothing rather than fail the bench
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_sm,
                "power_w_max": max(self.power) if self.power else None, "reasons": sorted(self.reasons), "samples": len(self.sm)}


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def probe_llvmpipe():
    """the reference's GLSL path can only be timed when a software GL (Mesa llvmpipe through
    surfaceless EGL or OSMesa) exists on the box.  Probed at run time, never assumed: returns a one-line verdict."""
    import ctypes.util
    import shutil
    import subprocess
    found = {}
    try:
        cache = subprocess.run(["ldconfig", "-p"], capture_output=True, text=True, timeout=20).stdout
    except Exception:
        cache = ""
    for lib in ("libEGL.so", "libOSMesa.so", "libGL.so", "libGLX_mesa.so", "libgallium"):
        hit = [ln.split("=>")[-1].strip() for ln in cache.splitlines() if lib in ln]
        if hit:
            found[lib] = hit[0]
    for name in ("EGL", "OSMesa", "GL"):
        f = ctypes.util.find_library(name)
        if f:
            found.setdefault("lib" + name, f)
    dri = [d for d in ("/usr/lib/x86_64-linux-gnu/dri", "/usr/lib64/dri", "/usr/lib/dri") if os.path.isdir(d)]
    swrast = [os.path.join(d, f) for d in dri for f in os.listdir(d) if "swrast" in f or "llvmpipe" in f]
    mesa_only = {k: v for k, v in found.items() if "nvidia" not in v.lower()}
    if swrast and (("libEGL.so" in mesa_only) or ("libOSMesa.so" in mesa_only)):
        return "present (%s; %s) -- the GLSL harness is not part of this round" % (swrast[0], sorted(mesa_only))
    return ("absent: no Mesa software rasteriser on this box (ldconfig/ctypes found %s; dri dirs %s; glxinfo %s) -- "
            "the reference's GLSL cannot execute here" % (sorted(found) or "no EGL/OSMesa/GL library", dri or "none",
                                                         "present" if shutil.which("glxinfo") else "absent"))


# ------------------------------------------------------------------------------------------------------------------
def oracle_sample(wl, vox, steps, warmup, with_sat_reference, reference_shader=False, sub=8):
    """CPU leg: a bounded sample of the workload (the same view, every sub-th ray per axis: 1/64 of the rays in the main arm, 1/16 in the reference arm) on the host cores.
    reference_shader=False: the oracle (OpenMP restatement, kind "port").  reference_shader=True: the REFERENCE'S OWN
    GLSL compute shader compiled for the CPU (oracle/_ref/librefglsl.so, oracle/glsl_cpu; kind "reference"), one
    invocation per ray over all host threads; the sample count comes from one untimed oracle run of the same rays (the
    two produce bit-identical frames, tests/test_refglsl.py).  Falls back to the oracle when that library is absent."""
    from cpp_volume_rendering_b200 import capi, synth
    from oracle import bind
    n, W, H = wl["n"], wl["W"], wl["H"]
    sw, sh = max(8, W // sub), max(8, H // sub)          # every sub-th ray per axis of the same view
    tf = bind.TF(*synth.TFS[wl["tf"]])
    eye, center, up = synth.camera_state(wl["cam"], n)
    cam = bind.camera(eye, center, up, sw, sh)
    # aspect must be the full frame's so the rays are a subsample of the same view
    cam.aspect = np.float32(np.float32(W) / np.float32(H))
    extra = {}
    if wl["renderer"] == "ebs":
        lut = tf.ext_lut(vox.dtype.itemsize)
        sat = None
        t0 = time.perf_counter()
        if with_sat_reference:
            try:                                              # the reference's own CPU SAT (oracle/_ref/libref.so) when it is there
                if bind.ref() is not None:
                    sat = np.empty((n + 2, n + 2, n + 2), np.float32)
                    bind.ref().ref_sat3d_from_volume(_p(vox), n, n, n, vox.dtype.itemsize, _p(lut), _p(sat))
                    kind = "reference (SummedAreaTable3D<double> from libs/vis_utils/summedareatable.h, 1 core as written)"
            except Exception as exc:                          # never lose the bench line over the CPU baseline
                sat = None
                extra["sat_build_reference_error"] = repr(exc)
        if sat is None:
            t0 = time.perf_counter()
            sat = bind.sat_build(vox, lut)
            kind = "port (oracle restatement of the same recurrence, 1 core)"
        extra["sat_build_cpu_s"] = time.perf_counter() - t0
        extra["sat_build_cpu_kind"] = kind
        light = capi.default_lighting(light_pos=synth.light_position(n), forward=synth.camera_forward(eye, center))
        prm = capi.default_ebs_params(float(np.sqrt(3.0) * n))
        ol, op = bind.copy_struct(light, bind.OrcLighting), bind.copy_struct(prm, bind.OrcEbsParams)
        tex = bind.volume_r16f(vox)
        rgbt = tf.texture_rgbt()
        sc = np.ones(3, np.float32)
        out = np.zeros((sh, sw, 4), np.float32); ns = np.zeros((sh, sw), np.uint32)

        def run():
            bind.orc().orc_ebs_render(_p(tex), n, n, n, _p(sc), _p(sat), _p(rgbt), tf.n, C.byref(cam), C.byref(ol), C.byref(op),
                                      sw, sh, _p(out), _p(ns))
    else:
        tex = bind.volume_r16f(vox)
        rgbt = tf.texture_rgbt()
        G = np.array([n, n, n], np.float32)
        out = np.zeros((sh, sw, 4), np.float32); ns = np.zeros((sh, sw), np.uint32)

        def run():
            bind.orc().orc_rc1pass_render(_p(tex), n, n, n, _p(G), _p(rgbt), tf.n, C.byref(cam), C.c_float(0.5), sw, sh, _p(out), _p(ns))
    kind = "port"
    engine = "OpenMP CPU restatement of the shader (oracle/)"
    if reference_shader:
        oracle_run = run
        try:
            from oracle import refglsl
            if refglsl.lib() is not None:
                run()                                         # untimed: the oracle's loop-iteration count of these rays
                ol_ = bind.OrcLighting() if wl["renderer"] != "ebs" else ol
                if wl["renderer"] == "ebs":
                    prog = refglsl.make_ebs(vox, tf, sat, cam, ol_, op)
                else:
                    prog = refglsl.make_rc1pass(vox, tf, cam, ol_, 0.5)
                frame = np.zeros((sh, sw, 4), np.float32)
                prog.image("OutputFrag", refglsl.Image(frame))
                oracle_frame = out.copy()

                def run():
                    prog.dispatch(sw, sh)
                run()
                extra["reference_shader_frame_equals_oracle"] = bool(np.array_equal(frame, oracle_frame, equal_nan=True))
                kind = "reference"
                engine = ("the reference's own GLSL compute shader (%s) compiled for the CPU from its source and dispatched over all host "
                          "threads (oracle/_ref/librefglsl.so)" % ("ebs_ray_bbox_marching.comp" if wl["renderer"] == "ebs" else "ray_marching_1p.comp"))
        except Exception as exc:                      # never lose the bench line over the CPU baseline
            run, kind = oracle_run, "port"
            engine = "OpenMP CPU restatement of the shader (oracle/); the reference-shader library failed: %r" % (exc,)
    for _ in range(warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    dt = (time.perf_counter() - t0) / steps
    samples = int(ns.sum())
    return dict(value=samples / dt / 1e9, ms_per_step=dt * 1e3, samples=samples, cores=bind.orc().orc_num_threads(), kind=kind, engine=engine,
                sample=f"same view subsampled to {sw}x{sh} rays (1/{sub * sub} of the frame), full {n}^3 volume", extra=extra)


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vox = make_volume(wl)
    r = oracle_sample(wl, vox, max(1, args.steps), max(0, min(args.warmup, 1)), with_sat_reference=True, reference_shader=True, sub=4)
    line = {
        "impl": "reference", "metric": "ray samples/sec", "value": r["value"], "unit": "Gsamples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "name": args.workload},
        "cpu_baseline": {"value": r["value"], "unit": "Gsamples/s", "cores": r["cores"], "kind": r["kind"],
                         "sample": r["sample"] + "; engine: " + r["engine"] + " (not llvmpipe: no GL exists in this image)"},
        "e2e": {"value": r["value"], "unit": "Gsamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "samples_per_step": r["samples"],
    }
    line["cpu_baseline"]["llvmpipe"] = probe_llvmpipe()
    line.update(r["extra"])
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------------
def run_vrb(args, wl):
    import torch
    import cpp_volume_rendering_b200 as vrb
    from cpp_volume_rendering_b200 import capi, synth
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, W, H = wl["n"], wl["W"], wl["H"]
    vox = make_volume(wl)
    bpv = vox.dtype.itemsize
    rgbt, rgba, lut = host_tf_arrays(wl["tf"], bpv)
    eye, center, up = synth.camera_state(wl["cam"], n)
    cam = capi.make_camera(eye, center, up, W, H)

    ctx = vrb.Context(local)
    if args.filter:
        ctx.set_filter(args.filter)
    # a real (non-NULL) stream: the kernels, the CUDA events that time them and the NCCL reduce all run on it
    stream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx.set_stream(stream.cuda_stream)
    ev = lambda: torch.cuda.Event(enable_timing=True)

    # ---- Init(): uploads + pre-passes (not part of the per-frame step; reported on the side)
    t0 = time.perf_counter()
    ctx.volume_upload(vox)
    ctx.tf_upload(rgbt, rgba)
    ctx.frame_resize(W, H)
    init = {"volume_upload_s": time.perf_counter() - t0}
    sat_info = None
    if wl["renderer"] == "ebs":
        def time_sat(order):
            ctx.sat_set_order(order)
            ctx.sat_build(lut)                   # warm-up build (allocations)
            e0, e1 = ev(), ev()
            reps = 3
            torch.cuda.synchronize()
            e0.record(stream)
            for _ in range(reps):
                ctx.sat_build(lut)
            e1.record(stream)
            torch.cuda.synchronize()
            # whole call (cudaMalloc/cudaFree of the scratch + atlas copy) and the build kernels alone (CUDA events inside the library)
            return e0.elapsed_time(e1) / reps, float(ctx.lib.vrb_last_prepass_ms(ctx.h))

        # the separable-scan build is the HBM-bound design point (roofline below); the build the frames are rendered with is
        # the default reference-order wavefront build (bit-identical floats to the reference's BuildSAT), timed next to it
        scan_call_ms, sat_ms = time_sat("scan")
        ref_call_ms, ref_ms = time_sat("reference")
        cells = (n + 2) ** 3
        sat_bytes = cells * (bpv + 36)
        peaks, peaks_src = read_peaks()
        sat_info = {"ms": sat_ms, "algorithmic_bytes": sat_bytes, "achieved_gbs": sat_bytes / (sat_ms * 1e-3) / 1e9,
                    "peak_gbs": peaks["hbm_gbs"], "frac": sat_bytes / (sat_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                    "peak_source": peaks_src, "call_ms_incl_alloc_and_atlas": scan_call_ms,
                    "note": "VRB_SAT_ORDER_SCAN: 3 scan kernels (CUDA events inside vrb_sat_build); b_v+36 B per bordered cell",
                    "reference_order_ms": ref_ms, "reference_order_call_ms": ref_call_ms,
                    "reference_order_note": "default build, used for the frames: BuildSAT's fp64 recurrence as %d anti-diagonal "
                                            "wavefront launches (latency-bound), float SAT bit-identical to the reference" % (3 * (n + 2) - 2),
                    "order_used": ctx.sat_get_order()}
        light = capi.default_lighting(light_pos=synth.light_position(n), forward=synth.camera_forward(eye, center))
        prm = capi.default_ebs_params(float(np.sqrt(3.0) * n))
    elif wl["renderer"] == "dos":
        diag = float(np.sqrt(3.0) * n)
        t0 = time.perf_counter()
        ctx.extcoef_build(1.0, (128, 128, 128))
        init["extcoef_pyramid_s"] = time.perf_counter() - t0
        occ, _, _ = capi.host_cone_sampler(20.0, 1, 0.5 * diag, 0.35)
        sdw, _, _ = capi.host_cone_sampler(0.5, 0, 0.75 * diag, 1.0)
        ctx.dos_set_cones(occ, sdw)
        light = capi.default_lighting(light_pos=synth.light_position(n), forward=synth.camera_forward(eye, center))
        prm = capi.default_dos_params(0.5, apply_shadow=True)
    elif wl["renderer"] == "gt":
        occ_r, sdw_r = capi.host_gt_ray_tables(wl["rays"], 90.0, wl["rays"], 1.0)
        ctx.gt_set_rays(occ_r, sdw_r)
        fwd = synth.camera_forward(eye, center)
        light = capi.default_lighting(light_pos=synth.light_position(n), forward=tuple(-f for f in fwd))
        prm = capi.default_gt_params(float(np.sqrt(3.0) * n), wl["rays"], wl["rays"])
    elif wl["renderer"] == "vct":
        t0 = time.perf_counter()
        ctx.vct_build(capi.host_opacity_by_density(synth.TFS[wl["tf"]], bpv))
        init["vct_prepass_s"] = time.perf_counter() - t0
        _, _, ms = ctx.vct_info()
        light = capi.default_lighting(light_pos=synth.light_position(n))
        prm = capi.default_vct_params(255.0 if bpv == 1 else 65535.0, ms)

    if world > 1:
        ctx.set_partition(rank, world, 32, 32)

    def render(count=False):
        if wl["renderer"] == "ebs":
            prm.count_samples = int(count)
            ctx.ebs_render(cam, light, prm)
        elif wl["renderer"] == "dos":
            prm.count_samples = int(count)
            ctx.dos_render(cam, light, prm)
        elif wl["renderer"] == "gt":
            prm.count_samples = int(count)
            ctx.gt_render(cam, light, prm)
        elif wl["renderer"] == "vct":
            prm.count_samples = int(count)
            ctx.vct_render(cam, light, prm)
        else:
            ctx.rc1pass_render(cam, 0.5, count_samples=count)

    # frame as a torch tensor (for the NCCL reduce of the sort-first partial frames)
    fptr, fw, fh = ctx.frame_device_ptr()

    class _Wrap:
        __cuda_array_interface__ = {"shape": (fh, fw, 4), "typestr": "<f2", "data": (fptr, False), "version": 2}
    frame_t = torch.as_tensor(_Wrap(), device=torch.device("cuda", local))

    # Assembling the sort-first frame on rank 0:
    #   p2p     (default) every rank's marcher stores its tiles straight into rank 0's frame buffer through a CUDA-IPC peer
    #           pointer (vrb_frame_set_target): transfer and render are the same kernel, the only collective is a one-int
    #           all-reduce acting as a stream-ordered barrier.  Two target buffers alternate so that frame i+1 never lands in
    #           the buffer rank 0 is still reading frame i from.
    #   reduce  every rank renders into its own zeroed frame, one NCCL reduce(SUM) of the 16.6 MB fp16 frame per step.
    use_p2p = world > 1 and args.assemble == "p2p"
    targets, step_no = None, [0]
    if use_p2p:
        if rank == 0:
            targets = [ctx.frame_extra(0), ctx.frame_extra(1)]
            handles = [ctx.ipc_export(t) for t in targets]
        else:
            handles = None
        box = [handles]
        dist.broadcast_object_list(box, src=0)
        if rank != 0:
            targets = [ctx.ipc_import(h) for h in box[0]]
        flag = torch.zeros(1, dtype=torch.int32, device="cuda")

    def step():
        if use_p2p:
            ctx.frame_set_target(targets[step_no[0] & 1])
            step_no[0] += 1
            render()
            dist.all_reduce(flag)            # barrier on the stream: rank 0's read-back waits for every rank's stores
        else:
            render()
            if world > 1:
                dist.reduce(frame_t, dst=0, op=dist.ReduceOp.SUM)   # tile sets are disjoint: x + 0 is exact in fp16

    # ---- workload size: loop iterations per frame (all ranks), SAT queries per frame
    render(count=True)
    counts = torch.tensor([ctx.last_sample_count, int(ctx.lib.vrb_last_aux_count(ctx.h))], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(counts)
    samples_per_frame, aux_per_frame = int(counts[0]), int(counts[1])

    launches0 = ctx.launches
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(physical_gpu_index(local)) if rank == 0 else None
    if sampler:
        sampler.start()
    l_before = ctx.launches
    e0, e1 = ev(), ev()
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms[0])
    gpu_launches = ctx.launches - l_before
    if sampler:
        sampler.stop_flag = True
        sampler.join()

    # ---- kernel-only duration of the dominant kernel (render call = 16 MB memset + the marcher), this rank
    k0, k1 = ev(), ev()
    torch.cuda.synchronize()
    k0.record(stream)
    for _ in range(args.steps):
        render()
    k1.record(stream)
    torch.cuda.synchronize()
    kern_ms = k0.elapsed_time(k1) / args.steps

    # ---- e2e: frame read back as float RGBA into pinned host memory every step.
    # (a) blocking vrb_frame_read_rgba32f per step (what the reference's glGetTexImage does);
    # (b) pipelined vrb_frame_read_rgba32f_async: the copy of frame i overlaps the render of frame i+1 (two pinned
    #     buffers, one frame of latency); every step's frame still lands inside the timed region.  (b) is `e2e`.
    pinned2 = [torch.empty((H, W, 4), dtype=torch.float32).pin_memory() for _ in range(2)]
    pinned = pinned2[(args.steps - 1) & 1]

    def e2e_loop(pipelined):
        for _ in range(2):
            step()
            if rank == 0:
                ctx.frame_read_into(pinned2[0].data_ptr())
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            step()
            if rank == 0:
                if pipelined:
                    ctx.frame_read_async(pinned2[i & 1].data_ptr())
                    ctx.frame_read_wait(1)                       # frame i-1 has landed
                else:
                    ctx.frame_read_into(pinned2[i & 1].data_ptr())   # synchronises the stream
        if rank == 0 and pipelined:
            ctx.frame_read_wait(0)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([(time.perf_counter() - t0) * 1e3], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms[0])

    e2e_sync_total_ms = e2e_loop(False)
    e2e_total_ms = e2e_loop(True)
    # fp32 differences of 1e7-sized SAT prefix sums make exp(-Stau) overflow the RGBA16F image in places at 512^3:
    # that is the reference's own result, so the checksum skips non-finite pixels
    checksum = float(torch.nan_to_num(pinned, nan=0.0, posinf=0.0, neginf=0.0).sum()) if rank == 0 else 0.0
    nonfinite = int((~torch.isfinite(pinned)).sum()) if rank == 0 else 0

    if rank == 0:
        peaks, peaks_src = read_peaks()
        l1 = C.c_double()
        ctx._ck(ctx.lib.vrb_measure_l1_bandwidth(ctx.h, C.byref(l1)))
        hb = C.c_double()
        ctx._ck(ctx.lib.vrb_measure_hbm_bandwidth(ctx.h, C.byref(hb)))
        gr = C.c_double()
        ctx._ck(ctx.lib.vrb_measure_gather_rate(ctx.h, C.byref(gr)))
        # algorithmic L1 bytes per frame: primary sample = 8 fp16 voxel taps + 2 RGBA16F TF texels = 32 B
        # (our texels are fp16 for u8 data too); SAT box query = 8 corners x 8 fp32 texels = 256 B
        aux_bytes = {"ebs": 256, "dos": 16, "gt": 32, "vct": 72}.get(wl["renderer"], 0)   # SURVEY.md section 8d per-unit figures
        l1_bytes = samples_per_frame * 32 + aux_per_frame * aux_bytes
        # share of this rank's kernel: with sort-first every rank does ~1/N of it
        l1_bytes_rank = l1_bytes / world
        unique_bytes = vox.size * 2 + (W * H * 8) + ((n + 2) ** 3 * 4 if wl["renderer"] == "ebs" else 0)
        sm_clock = (sampler.summary()["sm_mhz"] or 1965.0) if sampler else 1965.0
        l1_theory = 128.0 * 148 * sm_clock * 1e6 / 1e9
        line = {
            "metric": "ray samples/sec", "value": samples_per_frame * args.steps / (total_ms * 1e-3) / 1e9, "unit": "Gsamples/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "name": args.workload, "texture_filter": ctx.get_filter(),
                       "l2": "inputs larger than L2 (fp16 volume %.0f MB + SAT %.0f MB vs 126 MB L2)" %
                             (vox.size * 2 / 1e6, ((n + 2) ** 3 * 4 / 1e6) if wl["renderer"] == "ebs" else 0.0)
                             if unique_bytes > 126e6 else "working set fits L2 (L2-resident by design; no flush)",
                       "parallelism": "sort-first 32x32 tiles round-robin over %d GPU(s), volume replicated%s" % (
                           world, "" if world == 1 else (", frame assembled by peer stores into rank 0 (CUDA IPC) + 1-int all-reduce barrier"
                                                         if use_p2p else ", NCCL reduce(SUM) of the fp16 frame"))},
            "samples_per_frame": samples_per_frame, "sat_queries_per_frame": aux_per_frame if wl["renderer"] == "ebs" else None,
            "secondary_units_per_frame": aux_per_frame,
            "sat_layout": int(ctx.lib.vrb_sat_layout(ctx.h)) if wl["renderer"] == "ebs" else None,
            "ms_per_frame_kernel_only_rank0": kern_ms,
            "e2e": {"value": samples_per_frame * args.steps / (e2e_total_ms * 1e-3) / 1e9, "unit": "Gsamples/s",
                    "ms_per_step": e2e_total_ms / args.steps,
                    "readback": "pipelined (vrb_frame_read_rgba32f_async, 2 pinned buffers, 1 frame latency)",
                    "ms_per_step_blocking_readback": e2e_sync_total_ms / args.steps,
                    "h2d_bytes_per_step": C.sizeof(capi.Camera) + (C.sizeof(capi.Lighting) + C.sizeof(capi.EbsParams) if wl["renderer"] == "ebs" else C.sizeof(capi.Rc1passParams)),
                    "d2h_bytes_per_step": W * H * 16, "checksum": checksum, "nonfinite_values": nonfinite},
            "gpu_launches": int(gpu_launches),
            "roofline": {"bound": "l1tex", "kernel": {"ebs": "k_ebs_coop", "dos": "k_dos", "gt": "k_gt", "vct": "k_vct"}.get(wl["renderer"], "k_rc1pass"),
                         "achieved": l1_bytes_rank / (kern_ms * 1e-3) / 1e9, "peak": l1.value, "unit": "GB/s",
                         "frac": l1_bytes_rank / (kern_ms * 1e-3) / 1e9 / l1.value, "traffic": None,
                         "peak_source": "measured in this run (vrb_measure_l1_bandwidth: L1-resident LDG.128 on all SMs); "
                                        "theoretical 128 B/clk/SM x 148 x %.0f MHz = %.0f GB/s" % (sm_clock, l1_theory),
                         "algorithmic_bytes_per_launch": l1_bytes_rank},
            "roofline_hbm": {"bound": "hbm", "achieved": unique_bytes / (kern_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                             "frac": unique_bytes / (kern_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "peak_source": peaks_src,
                             "copy_bandwidth_this_run_gbs": hb.value,
                             "unique_bytes_per_launch": unique_bytes},
            "clocks": sampler.summary() if sampler else None,
            "init": init,
        }
        if ctx.get_filter() == "hardware" and wl["renderer"] in ("rc1pass", "dos", "gt", "vct"):
            # hardware filter mode: every volume / pyramid tap is one trilinear tex3D fetch, so the ceiling that matters is
            # the texture pipe's trilinear rate, measured in this run (guarded: an extra object, never the bench line itself)
            try:
                tr = C.c_double()
                ctx._ck(ctx.lib.vrb_measure_tex3d_rate(ctx.h, C.byref(tr)))
                # fetches per frame: one per primary sample; per secondary unit: DOS tap 1, GT step 1, VCT cone step 2 (two mip levels)
                per_aux = {"dos": 1.0, "gt": 1.0, "vct": 2.0}.get(wl["renderer"], 0.0)
                fetches = (samples_per_frame + aux_per_frame * per_aux) / world
                line["roofline_tex3d"] = {"bound": "l1tex", "kernel": line["roofline"]["kernel"], "achieved": fetches / (kern_ms * 1e-3) / 1e9,
                                          "peak": tr.value, "unit": "G trilinear fetches/s", "frac": fetches / (kern_ms * 1e-3) / 1e9 / tr.value,
                                          "peak_source": "measured in this run (vrb_measure_tex3d_rate: cache-resident R16F 3-D array, GL_LINEAR, "
                                                         "32 lanes on neighbouring texels)", "fetches_per_launch": fetches}
            except Exception as exc:
                line["roofline_tex3d"] = {"error": repr(exc)}
        if ctx.get_filter() == "exact" and wl["renderer"] in ("rc1pass", "dos", "gt", "vct"):
            # exact filter mode: every trilinear fetch is eight 16-bit loads (RG16F super-voxels: eight 32-bit ones), so the
            # ceiling that matters is the load pipe's rate for such scattered narrow loads, measured in this run (guarded)
            try:
                lr = C.c_double()
                ctx._ck(ctx.lib.vrb_measure_ldg16_rate(ctx.h, C.byref(lr)))
                per_aux = {"dos": 8.0, "gt": 8.0, "vct": 16.0}.get(wl["renderer"], 0.0)
                loads = (samples_per_frame * 8.0 + aux_per_frame * per_aux) / world
                line["roofline_ldg16"] = {"bound": "l1tex", "kernel": line["roofline"]["kernel"], "achieved": loads / (kern_ms * 1e-3) / 1e9,
                                          "peak": lr.value, "unit": "G lane-level loads/s", "frac": loads / (kern_ms * 1e-3) / 1e9 / lr.value,
                                          "peak_source": "measured in this run (vrb_measure_ldg16_rate: L1-resident fp16 brick, eight LDG.U16 per "
                                                         "trilinear footprint, 32 lanes on neighbouring texels)", "loads_per_launch": loads}
            except Exception as exc:
                line["roofline_ldg16"] = {"error": repr(exc)}
        if wl["renderer"] == "ebs":
            # the SAT box queries go through the texture pipe (tex2Dgather, 16 B per lane-level gather, 16 gathers per
            # query): that pipe, not LDG bandwidth, is what ncu shows saturated, so the headline fraction uses ITS
            # measured ceiling; the LDG-byte figure stays next to it
            line["roofline_l1_ldg"] = line["roofline"]
            gathers = aux_per_frame * 16.0 / world
            tex_bytes = gathers * 16.0 + samples_per_frame * 32.0 / world
            line["roofline"] = {
                "bound": "l1tex", "kernel": "k_ebs_coop", "achieved": tex_bytes / (kern_ms * 1e-3) / 1e9, "peak": gr.value * 16.0,
                "unit": "GB/s", "frac": tex_bytes / (kern_ms * 1e-3) / 1e9 / (gr.value * 16.0),
                "traffic": 283659264 if (args.workload == "cfg2" and world == 1) else None,
                "traffic_source": "ncu --set full, profiles/r1_v2_cfg2_k_ebs_coop.txt: dram__bytes_read.sum 274.96 MB + dram__bytes_write.sum 8.70 MB per launch "
                                  "(unique bytes 828 MB: the rays stop before most of the SAT is touched)",
                "peak_source": "measured in this run (vrb_measure_gather_rate: %.1f G lane-gathers/s x 16 B, cache-resident R32F tex2Dgather, "
                               "32 coherent lanes); ncu on the same kernel: l1tex data-pipe wavefronts 80.5 %% of peak, "
                               "3.4 active lanes per texture request" % gr.value,
                "algorithmic_bytes_per_launch": tex_bytes, "gathers_per_launch": gathers}
        if sat_info:
            line["roofline_sat"] = dict(bound="hbm", achieved=sat_info["achieved_gbs"], peak=sat_info["peak_gbs"], unit="GB/s",
                                        frac=sat_info["frac"], traffic=None, **{k: sat_info[k] for k in ("ms", "algorithmic_bytes", "peak_source", "note", "call_ms_incl_alloc_and_atlas",
                                                                                            "reference_order_ms", "reference_order_call_ms", "reference_order_note", "order_used")})
        if world == 1 and not args.no_cpu_baseline and wl["renderer"] in ("ebs", "rc1pass"):
            try:
                r = oracle_sample(wl, vox, 1, 0, with_sat_reference=True, reference_shader=True)
                line["cpu_baseline"] = {"value": r["value"], "unit": "Gsamples/s", "cores": r["cores"], "kind": r["kind"],
                                        "sample": r["sample"] + "; engine: " + r["engine"], "ms_per_sample_frame": r["ms_per_step"],
                                        "llvmpipe": probe_llvmpipe(), **r["extra"]}
            except Exception as exc:                      # the GPU numbers above are the bench line; never lose them over the CPU leg
                line["cpu_baseline"] = {"value": None, "unit": "Gsamples/s", "cores": 0, "kind": "port", "sample": "failed: %r" % (exc,)}
        print(json.dumps(line))
    if use_p2p:
        torch.cuda.synchronize()
        dist.barrier()
        ctx.frame_set_target(None)
        if rank != 0:
            for t in targets:
                ctx.ipc_close(t)
        dist.barrier()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="vrb", choices=["vrb", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--assemble", default="p2p", choices=["p2p", "reduce"],
                    help="N > 1: how the sort-first frame reaches rank 0 (peer stores from the marchers, or an NCCL reduce)")
    ap.add_argument("--filter", default=None, choices=["exact", "hardware"],
                    help="texture filtering of the marchers (default: the library's, see vrb_ctx_set_filter)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    import __graft_entry__ as g
    if int(os.environ.get("LOCAL_RANK", "0")) == 0:
        g.build()
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_vrb(args, wl)


if __name__ == "__main__":
    main()
