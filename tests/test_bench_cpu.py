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
