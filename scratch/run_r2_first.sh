#!/bin/bash
# First GPU call of round 2 (about 6 minutes on one B200):  /usr/local/graft/bin/gpurun --timeout 600 -- 'bash scratch/run_r2_first.sh'
# 1. the whole GPU suite (incl. the golden-frame and reference-shader comparisons), 2. smoke, 3. the default bench line and
# the reference arm, 4. every other workload in both filter modes (they now carry roofline_tex3d / roofline_ldg16 from the new
# vrb_measure_tex3d_rate / vrb_measure_ldg16_rate probes), 5. the ncu launch list of the default bench command.
mkdir -p gpurun_out
S=$(date +%s)
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2_pytest_gpu.txt
cat gpurun_out/r2_pytest_gpu.txt; echo "pytest done at $(( $(date +%s) - S )) s"
timeout 200 python tests/gpu_random_sweep.py --seed 1 --scenes 12 > gpurun_out/r2_random_sweep_exact.jsonl 2>&1; tail -2 gpurun_out/r2_random_sweep_exact.jsonl
timeout 200 python tests/gpu_random_sweep.py --seed 2 --scenes 12 --filter hardware > gpurun_out/r2_random_sweep_hardware.jsonl 2>&1; tail -2 gpurun_out/r2_random_sweep_hardware.jsonl
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r2_smoke.txt
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_cfg2.json 2> gpurun_out/r2_bench_cfg2.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_cfg2_reference.json 2> gpurun_out/r2_bench_cfg2_reference.err
echo "default bench + reference arm done at $(( $(date +%s) - S )) s"
for wl in cfg1 cfg3 cfg5-1gpu cfg4; do
  for f in exact hardware; do
    timeout 200 python bench.py --workload $wl --filter $f --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2_bench_${wl}_$f.json
  done
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r2_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    t3 = d.get("roofline_tex3d") or d.get("roofline_ldg16") or {}
    print(f.split("/")[-1], "ms", round(d.get("ms_per_step", 0), 3), "Gs/s", round(d.get("value", 0), 4), "frac", round((d.get("roofline") or {}).get("frac", 0), 3),
          "tex3d", t3.get("frac"), t3.get("peak"), t3.get("error"))
PY
echo "benches done at $(( $(date +%s) - S )) s"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_cfg2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
wc -l gpurun_out/r2_launches_cfg2.csv; echo "ncu done at $(( $(date +%s) - S )) s"
# memcheck of what was added after round 1's last GPU call: the two new ceiling probes and the Phong variant of k_obj_march
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python - > gpurun_out/r2_memcheck_new_kernels.txt 2>&1 <<'PY'
import ctypes as C
from cpp_volume_rendering_b200 import capi
ctx = capi.Context(0)
for fn in ("vrb_measure_tex3d_rate", "vrb_measure_ldg16_rate"):
    v = C.c_double()
    ctx._ck(getattr(ctx.lib, fn)(ctx.h, C.byref(v)))
    print(fn, v.value)
ctx.close()
PY
tail -4 gpurun_out/r2_memcheck_new_kernels.txt
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_zz_gpu_vs_reference_shader.py -q -m gpu -k "object_space" 2>&1 | tail -4
echo "memcheck done at $(( $(date +%s) - S )) s"
