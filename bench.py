#!/usr/bin/env python
"""bench.py -- ray samples/s (Gsamples/s) and ms/frame of the B200 volume marcher.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--impl vrb|reference] [--extras auto|none]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one frame of the workload: Update(camera) + Redraw of the renderer through the C ABI, volume / transfer
function / SAT resident in HBM (they are built once in Init(), like the reference does).  The headline workload is
config 2 of BASELINE.json: extinction-based shading (rc1pextbsd) of a 512^3 uint8 volume at 1920x1080.

  value     Gsamples/s = primary-ray loop iterations the single-GPU algorithm executes per frame (counted by the
            kernels' own counter, equal to the oracle's count) x K / device time, whole job over all N GPUs.
  e2e       same metric, each step additionally reading the float RGBA frame back into pinned host memory through
            vrb_frame_read_rgba32f (the reference's glGetTexImage(GL_RGBA, GL_FLOAT)); camera uniforms are the H2D.
  roofline  dominant kernel: algorithmic L1 bytes / that kernel's own CUDA-event duration
            (vrb_ctx_set_kernel_timing), against the ceiling measured in this run; roofline_hbm: unique bytes against
            MEASURED_PEAKS.json; `traffic` from this round's ncu capture (profiles/ncu_traffic.json); roofline_issue: the
            kernel's warp-instruction count (same capture) / the same live duration against the SM issue rate -- the
            ceiling the shade kernels of the lit renderers actually sit on.
  workloads the other BASELINE configs on the same box, same run: cfg1, cfg3, cfg4, cfg5-1gpu at N=1; at N>1 cfg3
            sort-first (the named scaling config) and, for power-of-two N, config 5 itself: 2048^3 u16 at 3840x2160 as
            sort-last bricks composited over NVLink (cpp_volume_rendering_b200/sort_last.py).
  cpu_baseline  the reference's own GLSL marcher compiled for the CPU (oracle/_ref/librefglsl.so, kind "reference"; the
            OpenMP oracle, kind "port", when that library is absent) on a bounded sample of the same workload, all host
            threads, rank 0, N=1.
  --impl reference  the reference arm: that CPU execution of the reference's shader + the reference's own
            SummedAreaTable3D for the SAT, on ALL host cores (the thread count is set explicitly: torchrun exports
            OMP_NUM_THREADS=1), same metric/config.  Loads nothing of the product.

N > 1: sort-first image tiles (volume replicated); every rank's kernels store their tiles straight into rank 0's frame
through CUDA-IPC peer pointers ("scaling": "strong": the frame is fixed, ranks split its tiles).
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
DOMINANT = {"ebs": "k_ebs_shade", "dos": "k_dos_shade", "gt": "k_gt_shade", "vct": "k_vct", "rc1pass": "k_rc1pass"}
# algorithmic L1 bytes per unit.  Primary sample: 8 fp16 voxel taps + 2 RGBA16F TF texels = 32 B
# (our texels are fp16 for u8 data too).  Secondary units: SAT box query 8 corners x 8 fp32 texels = 256 B; DOS cone tap
# 8 fp16 taps = 16 B; GT secondary step = a primary sample; VCT cone step 16 RG16F taps + 4 LUT taps = 72 B.
PRIMARY_BYTES = 32
AUX_BYTES = {"ebs": 256, "dos": 16, "gt": 32, "vct": 72, "rc1pass": 0}


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def make_volume(wl):
    from cpp_volume_rendering_b200 import synth
    return synth.make_volume(wl["volume"], wl["dtype"], wl["n"])


def host_tf_arrays(tfname, bpv):
    """TF textures + per-voxel-value extinction LUT from the C++ host mirror (product code, not the oracle)."""
    from cpp_volume_rendering_b200 import capi, synth
    return capi.host_tf_arrays(synth.TFS[tfname], bpv)


def config_for(args, wl, name, world):
    """The `config` object: built from the command line alone, so that both arms print the same keys and values."""
    n = wl["n"]
    vol_mb = n ** 3 * 2 / 1e6
    sat_mb = (n + 2) ** 3 * 4 / 1e6 if wl["renderer"] == "ebs" else 0.0
    big = vol_mb * 1e6 + sat_mb * 1e6 + wl["W"] * wl["H"] * 8 > 126e6
    par = "sort-first %dx%d tiles round-robin over %d GPU(s), volume replicated" % (args.tile, args.tile, world)
    if world > 1:
        par += (", frame assembled by peer stores into rank 0 (CUDA IPC) + 1-int all-reduce barrier" if args.assemble == "p2p"
                else ", NCCL reduce(SUM) of the fp16 frame")
    return {"workload": wl["desc"], "name": name, "texture_filter": args.filter or os.environ.get("VRB_FILTER", "exact"),
            "l2": ("inputs larger than L2 (fp16 volume %.0f MB + SAT %.0f MB vs 126 MB L2)" % (vol_mb, sat_mb)) if big
                  else "working set fits L2 (L2-resident by design; no flush)",
            "parallelism": par}


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
                time.sleep(0.02)
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


def read_ncu(name, kernel):
    """This round's ncu figures of a workload's dominant kernel (profiles/ncu_traffic.json, written by
    profiles/make_traffic.py from the `ncu --set full` summaries kept next to it); None when there is no capture."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        rec = json.load(open(p)).get(name)
    except Exception:
        return None
    if not rec or kernel not in rec.get("kernel", ""):
        return None
    return rec


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
def oracle_sample(wl, vox, steps, warmup, with_sat_reference, reference_shader=False, sub=8, threads=None):
    """CPU leg: a bounded sample of the workload (the same view, every sub-th ray per axis: 1/64 of the rays in the main arm, 1/16 in the reference arm) on the host cores.
    reference_shader=False: the oracle (OpenMP restatement, kind "port").  reference_shader=True: the REFERENCE'S OWN
    GLSL compute shader compiled for the CPU (oracle/_ref/librefglsl.so, oracle/glsl_cpu; kind "reference"), one
    invocation per ray over all host threads; the sample count comes from one untimed oracle run of the same rays (the
    two produce bit-identical frames, tests/test_refglsl.py).  Falls back to the oracle when that library is absent."""
    from cpp_volume_rendering_b200 import capi, synth
    from oracle import bind
    if threads:
        try:
            bind.orc().orc_set_num_threads(int(threads))      # one libgomp per process: also the reference-shader library's team size
        except AttributeError:
            pass
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
    """The reference arm.  Runs on rank 0 only; uses every host core whatever the launcher exported (torchrun sets
    OMP_NUM_THREADS=1 for N > 1, which made round 1's N>1 ratios meaningless); loads oracle/ only, never the product."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import subprocess
    threads = int(os.environ.get("VRB_REF_THREADS", "0")) or (len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count())
    os.environ["OMP_NUM_THREADS"] = str(threads)              # before libgomp is loaded (the oracle libraries load lazily)
    orc = os.path.join(ROOT, "oracle")
    if not os.path.exists(os.path.join(orc, "liboracle.so")):
        subprocess.check_call(["make", "-C", orc, "-s", "-B"])
    vox = make_volume(wl)
    r = oracle_sample(wl, vox, max(1, args.steps), max(0, min(args.warmup, 1)), with_sat_reference=True, reference_shader=True, sub=4, threads=threads)
    line = {
        "impl": "reference", "metric": "ray samples/sec", "value": r["value"], "unit": "Gsamples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_for(args, wl, args.workload, int(os.environ.get("WORLD_SIZE", "1"))),
        "cpu_baseline": {"value": r["value"], "unit": "Gsamples/s", "cores": r["cores"], "kind": r["kind"],
                         "sample": r["sample"] + "; engine: " + r["engine"] + " (not llvmpipe: no GL exists in this image)",
                         "omp_threads_requested": threads, "host_cpus": os.cpu_count()},
        "e2e": {"value": r["value"], "unit": "Gsamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "samples_per_step": r["samples"],
    }
    line["cpu_baseline"]["llvmpipe"] = probe_llvmpipe()
    line.update(r["extra"])
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------------
class Env:
    """Process-wide state of the product arm: rank, device, the one non-NULL stream everything runs on."""

    def __init__(self):
        import torch
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (there is no CPU fallback for the product path)")
        torch.cuda.set_device(self.local)
        self.dist = None
        if self.world > 1:
            import datetime
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local), timeout=datetime.timedelta(seconds=300))
            self.dist = dist
        # a real (non-NULL) stream: the kernels, the CUDA events that time them and the NCCL calls all run on it
        self.stream = torch.cuda.Stream(device=self.local)
        torch.cuda.set_stream(self.stream)
        assert self.stream.cuda_stream != 0

    def ev(self):
        return self.torch.cuda.Event(enable_timing=True)

    def sync(self):
        self.torch.cuda.synchronize()

    def barrier(self):
        if self.dist:
            self.dist.barrier()

    def max_over_ranks(self, v):
        t = self.torch.tensor([float(v)], dtype=self.torch.float64, device="cuda")
        if self.dist:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t[0])

    def sum_over_ranks(self, vals):
        t = self.torch.tensor([int(v) for v in vals], dtype=self.torch.int64, device="cuda")
        if self.dist:
            self.dist.all_reduce(t)
        return [int(v) for v in t]


def setup_renderer(env, ctx, wl, vox, init):
    """Init() of the workload's renderer through the C ABI -> (render(count), sat_info, h2d bytes per frame)."""
    from cpp_volume_rendering_b200 import capi, synth
    torch = env.torch
    n = wl["n"]
    bpv = vox.dtype.itemsize
    rgbt, rgba, lut = host_tf_arrays(wl["tf"], bpv)
    eye, center, up = synth.camera_state(wl["cam"], n)
    cam = capi.make_camera(eye, center, up, wl["W"], wl["H"])
    t0 = time.perf_counter()
    ctx.volume_upload(vox)
    ctx.tf_upload(rgbt, rgba)
    ctx.frame_resize(wl["W"], wl["H"])
    init["volume_upload_s"] = time.perf_counter() - t0
    sat_info = None
    light = prm = None
    h2d = C.sizeof(capi.Camera)
    if wl["renderer"] == "ebs":
        def time_sat(order):
            ctx.sat_set_order(order)
            ctx.sat_build(lut)                   # warm-up build (allocations)
            e0, e1 = env.ev(), env.ev()
            reps = 3
            env.sync()
            e0.record(env.stream)
            for _ in range(reps):
                ctx.sat_build(lut)
            e1.record(env.stream)
            env.sync()
            # whole call (cudaMalloc/cudaFree of the scratch + atlas copy) and the build kernels alone (CUDA events inside the library)
            return e0.elapsed_time(e1) / reps, float(ctx.lib.vrb_last_prepass_ms(ctx.h))

        # the separable-scan build is the HBM-bound design point; the build the frames are rendered with is the default
        # reference-order build (bit-identical floats to the reference's BuildSAT), timed next to it
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
                    "reference_order_frac": sat_bytes / (ref_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                    "reference_order_note": "default build, used for the frames: BuildSAT's fp64 recurrence association reproduced, float SAT "
                                            "bit-identical to the reference",
                    "order_used": ctx.sat_get_order()}
        light = capi.default_lighting(light_pos=synth.light_position(n), forward=synth.camera_forward(eye, center))
        prm = capi.default_ebs_params(float(np.sqrt(3.0) * n))
        h2d += C.sizeof(capi.Lighting) + C.sizeof(capi.EbsParams)
    elif wl["renderer"] == "dos":
        diag = float(np.sqrt(3.0) * n)
        t0 = time.perf_counter()
        ctx.extcoef_build(1.0, (128, 128, 128))
        ctx.synchronize()
        init["extcoef_pyramid_s"] = time.perf_counter() - t0
        occ, _, _ = capi.host_cone_sampler(20.0, 1, 0.5 * diag, 0.35)
        sdw, _, _ = capi.host_cone_sampler(0.5, 0, 0.75 * diag, 1.0)
        ctx.dos_set_cones(occ, sdw)
        light = capi.default_lighting(light_pos=synth.light_position(n), forward=synth.camera_forward(eye, center))
        prm = capi.default_dos_params(0.5, apply_shadow=True)
        h2d += C.sizeof(capi.Lighting) + C.sizeof(capi.DosParams)
    elif wl["renderer"] == "gt":
        occ_r, sdw_r = capi.host_gt_ray_tables(wl["rays"], 90.0, wl["rays"], 1.0)
        ctx.gt_set_rays(occ_r, sdw_r)
        fwd = synth.camera_forward(eye, center)
        light = capi.default_lighting(light_pos=synth.light_position(n), forward=tuple(-f for f in fwd))
        prm = capi.default_gt_params(float(np.sqrt(3.0) * n), wl["rays"], wl["rays"])
        h2d += C.sizeof(capi.Lighting) + C.sizeof(capi.GtParams)
    elif wl["renderer"] == "vct":
        t0 = time.perf_counter()
        ctx.vct_build(capi.host_opacity_by_density(synth.TFS[wl["tf"]], bpv))
        ctx.synchronize()
        init["vct_prepass_s"] = time.perf_counter() - t0
        _, _, ms = ctx.vct_info()
        light = capi.default_lighting(light_pos=synth.light_position(n))
        prm = capi.default_vct_params(255.0 if bpv == 1 else 65535.0, ms)
        h2d += C.sizeof(capi.Lighting) + C.sizeof(capi.VctParams)
    else:
        h2d += C.sizeof(capi.Rc1passParams)

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
    return render, sat_info, h2d


def run_workload(env, args, name, steps, warmup, full):
    """One workload on the current process group: returns the measurement dict on rank 0 (None elsewhere).
    full=True: the headline line (clock sampling, blocking + pipelined e2e, every roofline, cpu_baseline);
    full=False: an entry of `workloads` (device-timed steps, pipelined e2e, the dominant kernel's roofline)."""
    import cpp_volume_rendering_b200 as vrb
    torch, dist, rank, world, local, stream = env.torch, env.dist, env.rank, env.world, env.local, env.stream
    wl = WORKLOADS[name]
    n, W, H = wl["n"], wl["W"], wl["H"]
    vox = make_volume(wl)
    ctx = vrb.Context(local)
    if args.filter:
        ctx.set_filter(args.filter)
    ctx.set_stream(stream.cuda_stream)
    init = {}
    render, sat_info, h2d_bytes = setup_renderer(env, ctx, wl, vox, init)
    if world > 1:
        ctx.set_partition(rank, world, args.tile, args.tile)

    # frame as a torch tensor (for the NCCL reduce of the sort-first partial frames, --assemble reduce)
    fptr, fw, fh = ctx.frame_device_ptr()

    class _Wrap:
        __cuda_array_interface__ = {"shape": (fh, fw, 4), "typestr": "<f2", "data": (fptr, False), "version": 2}
    frame_t = torch.as_tensor(_Wrap(), device=torch.device("cuda", local))

    # Assembling the sort-first frame on rank 0:
    #   p2p     (default) every rank's kernels store their tiles straight into rank 0's frame buffer through a CUDA-IPC peer
    #           pointer (vrb_frame_set_target): transfer and render are the same kernel.  Completion is a stream-ordered
    #           barrier.  Two target buffers alternate so that frame i+1 never lands in the buffer rank 0 is still reading.
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

    # ---- workload size: loop iterations per frame (all ranks), secondary units per frame
    render(count=True)
    samples_per_frame, aux_per_frame = env.sum_over_ranks([ctx.last_sample_count, int(ctx.lib.vrb_last_aux_count(ctx.h))])

    for _ in range(warmup):
        step()
    env.sync()
    env.barrier()
    sampler = ClockSampler(physical_gpu_index(local)) if (rank == 0 and full) else None
    if sampler:
        sampler.start()
    l_before = ctx.launches
    e0, e1 = env.ev(), env.ev()
    env.sync()
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    env.sync()
    env.barrier()
    total_ms = env.max_over_ranks(e0.elapsed_time(e1))
    gpu_launches = ctx.launches - l_before
    if sampler:
        sampler.stop_flag = True
        sampler.join()

    # ---- the render call alone (no assembling barrier), and its dominant kernel alone (events inside the library)
    ksteps = steps if full else min(steps, 3)
    k0, k1 = env.ev(), env.ev()
    env.sync()
    k0.record(stream)
    for _ in range(ksteps):
        render()
    k1.record(stream)
    env.sync()
    render_ms = k0.elapsed_time(k1) / ksteps
    render_ms_ranks = [render_ms]
    if dist:
        t_all = torch.zeros(world, dtype=torch.float64, device="cuda")
        t_all[rank] = render_ms
        dist.all_reduce(t_all)
        render_ms_ranks = [float(v) for v in t_all]
    ctx.set_kernel_timing(True)
    dom = []
    dom_name = DOMINANT[wl["renderer"]]
    for _ in range(ksteps):
        render()
        ms_k, dom_name = ctx.last_kernel_ms()
        dom.append(ms_k)
    ctx.set_kernel_timing(False)
    kern_ms = float(np.mean(dom))

    # ---- e2e: frame read back as float RGBA into pinned host memory every step.
    # (a) blocking vrb_frame_read_rgba32f per step (what the reference's glGetTexImage does);
    # (b) pipelined vrb_frame_read_rgba32f_async: the copy of frame i overlaps the render of frame i+1 (two pinned
    #     buffers, one frame of latency); every step's frame still lands inside the timed region.  (b) is `e2e`.
    esteps = steps if full else min(steps, 5)
    pinned2 = [torch.empty((H, W, 4), dtype=torch.float32).pin_memory() for _ in range(2)]
    pinned = pinned2[(esteps - 1) & 1]

    def e2e_loop(pipelined):
        for _ in range(2):
            step()
            if rank == 0:
                ctx.frame_read_into(pinned2[0].data_ptr())
        env.sync()
        env.barrier()
        t0 = time.perf_counter()
        for i in range(esteps):
            step()
            if rank == 0:
                if pipelined:
                    ctx.frame_read_async(pinned2[i & 1].data_ptr())
                    ctx.frame_read_wait(1)                       # frame i-1 has landed
                else:
                    ctx.frame_read_into(pinned2[i & 1].data_ptr())   # synchronises the stream
        if rank == 0 and pipelined:
            ctx.frame_read_wait(0)
        env.sync()
        env.barrier()
        return env.max_over_ranks((time.perf_counter() - t0) * 1e3)

    e2e_sync_total_ms = e2e_loop(False) if full else None
    e2e_total_ms = e2e_loop(True)
    # fp32 differences of 1e7-sized SAT prefix sums make exp(-Stau) overflow the RGBA16F image in places at 512^3:
    # that is the reference's own result, so the checksum skips non-finite pixels
    checksum = float(torch.nan_to_num(pinned, nan=0.0, posinf=0.0, neginf=0.0).sum()) if rank == 0 else 0.0
    nonfinite = int((~torch.isfinite(pinned)).sum()) if rank == 0 else 0

    line = None
    if rank == 0:
        peaks, peaks_src = read_peaks()
        l1 = C.c_double()
        ctx._ck(ctx.lib.vrb_measure_l1_bandwidth(ctx.h, C.byref(l1)))
        rdr = wl["renderer"]
        # bytes the DOMINANT kernel moves through L1 per launch on this rank (sort-first: ~1/N of the frame's).  The deferred
        # lit renderers shade in their own kernel: its bytes are the secondary units'; the primary samples belong to the march kernel
        deferred = dom_name.endswith("_shade")
        dom_units_bytes = (aux_per_frame * AUX_BYTES[rdr] + (0 if deferred else samples_per_frame * PRIMARY_BYTES)) / world
        unique_bytes = vox.size * 2 + (W * H * 8) + ((n + 2) ** 3 * 4 if rdr == "ebs" else 0)
        sm_clock = (sampler.summary()["sm_mhz"] or 1965.0) if sampler else 1965.0
        l1_theory = 128.0 * 148 * sm_clock * 1e6 / 1e9
        ncu = read_ncu(name, dom_name) if world == 1 else None
        line = {
            "metric": "ray samples/sec", "value": samples_per_frame * steps / (total_ms * 1e-3) / 1e9, "unit": "Gsamples/s",
            "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": total_ms / steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_for(args, wl, name, world),
            "samples_per_frame": samples_per_frame, "sat_queries_per_frame": aux_per_frame if rdr == "ebs" else None,
            "secondary_units_per_frame": aux_per_frame,
            "secondary_units_per_s_G": aux_per_frame * steps / (total_ms * 1e-3) / 1e9,
            "sat_layout": int(ctx.lib.vrb_sat_layout(ctx.h)) if rdr == "ebs" else None,
            "ms_per_frame_render_call_rank0": render_ms, "ms_per_frame_render_call_by_rank": render_ms_ranks, "ms_dominant_kernel_rank0": kern_ms, "dominant_kernel": dom_name,
            "e2e": {"value": samples_per_frame * esteps / (e2e_total_ms * 1e-3) / 1e9, "unit": "Gsamples/s",
                    "ms_per_step": e2e_total_ms / esteps, "steps": esteps,
                    "readback": "pipelined (vrb_frame_read_rgba32f_async, 2 pinned buffers, 1 frame latency)",
                    "ms_per_step_blocking_readback": (e2e_sync_total_ms / esteps) if e2e_sync_total_ms is not None else None,
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": W * H * 16, "checksum": checksum, "nonfinite_values": nonfinite},
            "gpu_launches": int(gpu_launches),
            "roofline": {"bound": "l1tex", "kernel": dom_name,
                         "achieved": dom_units_bytes / (kern_ms * 1e-3) / 1e9, "peak": l1.value, "unit": "GB/s",
                         "frac": dom_units_bytes / (kern_ms * 1e-3) / 1e9 / l1.value,
                         "traffic": (ncu["dram_bytes_read"] + ncu["dram_bytes_write"]) if ncu else None,
                         "peak_source": "measured in this run (vrb_measure_l1_bandwidth: L1-resident LDG.128 on all SMs); "
                                        "theoretical 128 B/clk/SM x 148 x %.0f MHz = %.0f GB/s" % (sm_clock, l1_theory),
                         "algorithmic_bytes_per_launch": dom_units_bytes,
                         "duration_source": "CUDA events around the kernel on its launching stream (vrb_ctx_set_kernel_timing), mean of %d launches" % ksteps},
            "roofline_hbm": {"bound": "hbm", "achieved": unique_bytes / (render_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                             "frac": unique_bytes / (render_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "peak_source": peaks_src,
                             "unique_bytes_per_launch": unique_bytes},
            "init": init,
        }
        if ncu:
            line["roofline"]["traffic_source"] = ncu.get("source")
            line["ncu"] = {k: ncu[k] for k in ncu if k not in ("dram_bytes_read", "dram_bytes_write")}
            # issue-slot roofline of the dominant kernel: warp instructions per launch (counted by ncu for this kernel on this
            # workload; the count does not depend on timing) / the kernel's live CUDA-event duration, against 4 warp
            # instructions per clock per SM at the SM clock sampled in this run.  This is the ceiling the shade kernels
            # sit on (the L1-byte `roofline` above is a loose bound for them).
            if ncu.get("warp_instructions_per_launch"):
                sms = torch.cuda.get_device_properties(local).multi_processor_count
                peak_issue = 4.0 * sms * sm_clock * 1e6 / 1e9
                ach_issue = ncu["warp_instructions_per_launch"] / (kern_ms * 1e-3) / 1e9
                line["roofline_issue"] = {"bound": "issue", "kernel": dom_name, "achieved": ach_issue, "peak": peak_issue,
                                          "unit": "G warp instructions/s", "frac": ach_issue / peak_issue,
                                          "warp_instructions_per_launch": ncu["warp_instructions_per_launch"],
                                          "peak_source": "4 warp instructions / clk / SM x %d SMs x %.0f MHz" % (sms, sm_clock),
                                          "count_source": ncu.get("source")}
        if sampler:
            line["clocks"] = sampler.summary()
        if full:
            hb = C.c_double()
            ctx._ck(ctx.lib.vrb_measure_hbm_bandwidth(ctx.h, C.byref(hb)))
            line["roofline_hbm"]["copy_bandwidth_this_run_gbs"] = hb.value
        if ctx.get_filter() == "hardware" and rdr in ("rc1pass", "dos", "gt", "vct"):
            # hardware filter mode: every volume / pyramid tap is one trilinear tex3D fetch, so the ceiling that matters is
            # the texture pipe's trilinear rate, measured in this run (guarded: an extra object, never the bench line itself)
            try:
                tr = C.c_double()
                ctx._ck(ctx.lib.vrb_measure_tex3d_rate(ctx.h, C.byref(tr)))
                # fetches per frame: one per primary sample; per secondary unit: DOS tap 1, GT step 1, VCT cone step 2 (two mip levels)
                per_aux = {"dos": 1.0, "gt": 1.0, "vct": 2.0}.get(rdr, 0.0)
                fetches = (samples_per_frame + aux_per_frame * per_aux) / world
                line["roofline_tex3d"] = {"bound": "l1tex", "kernel": dom_name, "achieved": fetches / (render_ms * 1e-3) / 1e9,
                                          "peak": tr.value, "unit": "G trilinear fetches/s", "frac": fetches / (render_ms * 1e-3) / 1e9 / tr.value,
                                          "peak_source": "measured in this run (vrb_measure_tex3d_rate: cache-resident R16F 3-D array, GL_LINEAR, "
                                                         "32 lanes on neighbouring texels)", "fetches_per_launch": fetches}
            except Exception as exc:
                line["roofline_tex3d"] = {"error": repr(exc)}
        if ctx.get_filter() == "exact" and rdr in ("rc1pass", "dos", "gt", "vct"):
            # exact filter mode: every trilinear fetch is eight 16-bit loads (RG16F super-voxels: eight 32-bit ones), so the
            # ceiling that matters is the load pipe's rate for such scattered narrow loads, measured in this run (guarded)
            try:
                lr = C.c_double()
                ctx._ck(ctx.lib.vrb_measure_ldg16_rate(ctx.h, C.byref(lr)))
                per_aux = {"dos": 8.0, "gt": 8.0, "vct": 16.0}.get(rdr, 0.0)
                loads = ((0 if deferred else samples_per_frame * 8.0) + aux_per_frame * per_aux) / world
                line["roofline_ldg16"] = {"bound": "l1tex", "kernel": dom_name, "achieved": loads / (kern_ms * 1e-3) / 1e9,
                                          "peak": lr.value, "unit": "G lane-level loads/s", "frac": loads / (kern_ms * 1e-3) / 1e9 / lr.value,
                                          "peak_source": "measured in this run (vrb_measure_ldg16_rate: L1-resident fp16 brick, eight LDG.U16 per "
                                                         "trilinear footprint, 32 lanes on neighbouring texels)", "loads_per_launch": loads}
            except Exception as exc:
                line["roofline_ldg16"] = {"error": repr(exc)}
        if rdr == "ebs":
            # the SAT box queries go through the texture pipe (tex2Dgather, 16 B per lane-level gather, 16 gathers per
            # query): that pipe, not LDG bandwidth, is what ncu shows saturated, so the headline fraction uses ITS
            # measured ceiling; the LDG-byte figure stays next to it as roofline_l1_ldg
            gr = C.c_double()
            ctx._ck(ctx.lib.vrb_measure_gather_rate(ctx.h, C.byref(gr)))
            line["roofline_l1_ldg"] = line["roofline"]
            gathers = aux_per_frame * 16.0 / world
            tex_bytes = gathers * 16.0 + (0 if deferred else samples_per_frame * 32.0 / world)
            line["roofline"] = dict(line["roofline_l1_ldg"])
            line["roofline"].update({
                "achieved": tex_bytes / (kern_ms * 1e-3) / 1e9, "peak": gr.value * 16.0,
                "frac": tex_bytes / (kern_ms * 1e-3) / 1e9 / (gr.value * 16.0),
                "peak_source": "measured in this run (vrb_measure_gather_rate: %.1f G lane-gathers/s x 16 B, cache-resident R32F tex2Dgather, "
                               "32 coherent lanes)" % gr.value,
                "algorithmic_bytes_per_launch": tex_bytes, "gathers_per_launch": gathers})
        if sat_info:
            line["roofline_sat"] = dict(bound="hbm", achieved=sat_info["achieved_gbs"], peak=sat_info["peak_gbs"], unit="GB/s",
                                        frac=sat_info["frac"], traffic=None,
                                        **{k: sat_info[k] for k in ("ms", "algorithmic_bytes", "peak_source", "note", "call_ms_incl_alloc_and_atlas",
                                                                    "reference_order_ms", "reference_order_call_ms", "reference_order_frac",
                                                                    "reference_order_note", "order_used")})
        if full and world == 1 and not args.no_cpu_baseline and rdr in ("ebs", "rc1pass"):
            try:
                r = oracle_sample(wl, vox, 1, 0, with_sat_reference=True, reference_shader=True)
                line["cpu_baseline"] = {"value": r["value"], "unit": "Gsamples/s", "cores": r["cores"], "kind": r["kind"],
                                        "sample": r["sample"] + "; engine: " + r["engine"], "ms_per_sample_frame": r["ms_per_step"],
                                        "llvmpipe": probe_llvmpipe(), **r["extra"]}
            except Exception as exc:                      # the GPU numbers above are the bench line; never lose them over the CPU leg
                line["cpu_baseline"] = {"value": None, "unit": "Gsamples/s", "cores": 0, "kind": "port", "sample": "failed: %r" % (exc,)}
    if use_p2p:
        env.sync()
        env.barrier()
        ctx.frame_set_target(None)
        if rank != 0:
            for t in targets:
                ctx.ipc_close(t)
        env.barrier()
    del frame_t
    ctx.close()
    return line


EXTRA_KEYS = ("value", "unit", "ms_per_step", "steps", "warmup", "samples_per_frame", "secondary_units_per_frame", "secondary_units_per_s_G",
              "ms_per_frame_render_call_rank0", "ms_per_frame_render_call_by_rank", "ms_dominant_kernel_rank0", "dominant_kernel", "e2e", "gpu_launches", "roofline", "roofline_issue", "roofline_ldg16",
              "roofline_l1_ldg", "roofline_hbm", "ncu", "config", "init")


def run_vrb(args):
    env = Env()
    warmup = max(args.warmup, 3)
    line = run_workload(env, args, args.workload, args.steps, warmup, full=True)
    extras = {}
    if args.extras != "none" and args.workload == "cfg2":
        names = ["cfg1", "cfg3", "cfg4", "cfg5-1gpu"] if env.world == 1 else ["cfg3"]
        for nm in names:
            heavy = WORKLOADS[nm]["renderer"] == "gt"
            t0 = time.perf_counter()
            try:
                r = run_workload(env, args, nm, min(args.steps, 3 if heavy else 10), 3, full=False)
                if r is not None:
                    extras[nm] = {k: r[k] for k in EXTRA_KEYS if k in r}
                    extras[nm]["wall_s"] = time.perf_counter() - t0
            except Exception as exc:              # an extra workload never costs the headline line (single rank; N > 1: the group's timeout)
                extras[nm] = {"error": repr(exc)}
        if env.world >= 2 and (env.world & (env.world - 1)) == 0:
            t0 = time.perf_counter()
            try:
                from cpp_volume_rendering_b200 import sort_last
                r = sort_last.run(env, n=args.cfg5_res, W=3840, H=2160, dtype="u16", renderer="vct", steps=min(args.steps, 5),
                                  filter_mode=args.filter or "exact", gen="device")
                if r is not None:
                    r["wall_s"] = time.perf_counter() - t0
                    extras["cfg5"] = r
            except Exception as exc:
                extras["cfg5"] = {"error": repr(exc)}
    if env.rank == 0:
        if extras:
            line["workloads"] = extras
        print(json.dumps(line))
    if env.dist:
        env.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="vrb", choices=["vrb", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--extras", default="auto", choices=["auto", "none"],
                    help="auto: with the default workload also time the other BASELINE configs into the line's `workloads` object")
    ap.add_argument("--cfg5-res", type=int, default=2048, help="volume edge of the sort-last config-5 extra (N > 1)")
    ap.add_argument("--assemble", default="p2p", choices=["p2p", "reduce"],
                    help="N > 1: how the sort-first frame reaches rank 0 (peer stores from the marchers, or an NCCL reduce)")
    ap.add_argument("--tile", type=int, default=16, help="N > 1: edge of the sort-first image tiles (multiple of 16)")
    ap.add_argument("--filter", default=None, choices=["exact", "hardware"],
                    help="texture filtering of the marchers (default: the library's, see vrb_ctx_set_filter)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args, WORKLOADS[args.workload])
    else:
        # the product arm loads the prebuilt in-tree libraries (capi.load raises when libvrb200.so is missing: no fallback)
        # and touches oracle/ only in the cpu_baseline leg after the timed regions
        run_vrb(args)


if __name__ == "__main__":
    main()
