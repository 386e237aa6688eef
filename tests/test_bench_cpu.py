"""bench.py on the CPU: the reference arm (`--impl reference`) runs end to end on the small workload and prints the JSON
line the driver expects; the product arm refuses to run without a CUDA device (no CPU fallback); ranks other than 0 of the
reference arm exit quietly."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, cwd=ROOT, env=e, timeout=600)


@pytest.mark.parametrize("workload", ["cfg2-small", "cfg1"])
def test_reference_arm_prints_the_contract_line(built, workload):
    r = _run(["--impl", "reference", "--workload", workload, "--steps", "2", "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "Gsamples/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["steps"] == 2 and line["n_gpus"] == 1
    assert line["e2e"] == {"value": line["value"], "unit": "Gsamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["value"] == line["value"] and cb["cores"] >= 1 and cb["kind"] in ("reference", "port") and "sample" in cb and "llvmpipe" in cb
    if cb["kind"] == "reference":                       # the reference's own shader ran: its frame equals the oracle's
        assert line["reference_shader_frame_equals_oracle"] is True


def test_reference_arm_other_ranks_stay_silent(built):
    r = _run(["--impl", "reference", "--workload", "cfg2-small", "--steps", "1", "--warmup", "0", "--gpus", "2"], env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == "", (r.stdout, r.stderr[-500:])


def test_product_arm_fails_loudly_without_a_gpu(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = _run(["--workload", "cfg2-small", "--steps", "1", "--warmup", "1"])
    assert r.returncode != 0
    assert "no CUDA device" in (r.stdout + r.stderr)


def test_both_arms_build_the_same_config_object():
    """`same_config` was false only because the product arm added keys the reference arm did not have.  The
    object is now built by one function from the command line alone."""
    sys.path.insert(0, ROOT)
    import argparse
    import bench
    for world in (1, 8):
        args = argparse.Namespace(filter=None, assemble="p2p", tile=16)
        a = bench.config_for(args, bench.WORKLOADS["cfg2"], "cfg2", world)
        b = bench.config_for(args, bench.WORKLOADS["cfg2"], "cfg2", world)
        assert a == b and set(a) == {"workload", "name", "texture_filter", "l2", "parallelism"}
        assert ("%d GPU(s)" % world) in a["parallelism"]


def test_reference_arm_sets_its_own_thread_count(built):
    """torchrun exports OMP_NUM_THREADS=1 for N > 1; the reference arm must still use every core."""
    r = _run(["--impl", "reference", "--workload", "cfg2-small", "--steps", "1", "--warmup", "0", "--gpus", "2"],
             env={"RANK": "0", "LOCAL_RANK": "0", "WORLD_SIZE": "2", "OMP_NUM_THREADS": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    want = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    assert line["cpu_baseline"]["cores"] == want and line["cpu_baseline"]["omp_threads_requested"] == want


def test_reference_arm_does_not_load_the_product_library(built):
    """The reference arm's process must not map libvrb200.so (the driver lists the loaded .so files per arm)."""
    code = ("import sys, os; sys.argv = ['bench.py', '--impl', 'reference', '--workload', 'cfg2-small', '--steps', '1', '--warmup', '0'];"
            "import runpy; runpy.run_path(os.path.join(%r, 'bench.py'), run_name='__main__');"
            "maps = open('/proc/self/maps').read(); print('LOADED_VRB200' if 'libvrb200.so' in maps else 'CLEAN')" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip().splitlines()[-1] == "CLEAN"


def test_ncu_traffic_file_matches_its_sources():
    """profiles/ncu_traffic.json (read by bench.py for `roofline.traffic`) is generated from the summaries next to it."""
    sys.path.insert(0, os.path.join(ROOT, "profiles"))
    import make_traffic
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    assert os.path.exists(path)
    rec = json.load(open(path))
    assert {"cfg1", "cfg2", "cfg3", "cfg4", "cfg5-1gpu"} <= set(rec)
    for wl, f in make_traffic.SOURCES.items():
        got = make_traffic.parse(os.path.join(ROOT, "profiles", f))
        assert rec[wl]["dram_bytes_read"] == got["dram_bytes_read"] and rec[wl]["kernel"] == got["kernel"], wl
    sys.path.insert(0, ROOT)
    import bench
    assert bench.read_ncu("cfg3", "k_dos_shade")["dram_bytes_read"] > 0
    assert bench.read_ncu("cfg3", "k_some_other_kernel") is None      # a stale capture of another kernel is not reported
