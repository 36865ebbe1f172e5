"""CPU: what can be checked of bench.py without a GPU -- the reference arm (`--impl reference`: the oracle port of the
reference's deflate on the host cores, a bounded sample per step) prints ONE JSON line with the contract's keys for
the N = 1 and the N > 1 workload, a non-zero rank of a torchrun launch prints nothing, and the product arm refuses to
run without a CUDA device (there is no CPU fallback)."""
import json
import os
import subprocess
import sys

from conftest import ROOT

BENCH = os.path.join(ROOT, "bench.py")


def _run(args, env=None):
    e = dict(os.environ)
    e.pop("RANK", None)
    e.update(env or {})
    return subprocess.run([sys.executable, BENCH, *args], capture_output=True, text=True, timeout=300, env=e)


def test_reference_arm_lines():
    for args, workload, scaling in ((["--size-mib", "16"], "configs[1]", "weak"), (["--gpus", "2", "--total-mib", "32"], "configs[2]", "strong")):
        p = _run(["--impl", "reference", "--steps", "1", "--warmup", "1", *args])
        assert p.returncode == 0, p.stderr[-2000:]
        lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
        assert len(lines) == 1
        d = json.loads(lines[0])
        assert d["impl"] == "reference" and d["metric"] == "deflate_input_GBps" and d["unit"] == "GB/s"
        assert d["higher_is_better"] is True and d["scaling"] == scaling and d["steps"] == 1 and d["warmup"] == 1
        assert d["config"]["workload"].startswith(workload) and d["value"] > 0 and d["ms_per_step"] > 0
        assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
        assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
        assert d["gpu_launches"] == 0 and d["vs_baseline"] is None and d["dtype"] == "u8"
        assert 0.2 < d["compressed_ratio"] < 0.8
    # under torchrun only rank 0 runs the arm: the other ranks exit 0 without a line
    p = _run(["--impl", "reference", "--gpus", "2", "--total-mib", "32", "--steps", "1", "--warmup", "0"], env={"RANK": "1"})
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    p = _run(["--steps", "1", "--warmup", "1", "--size-mib", "16"])
    assert p.returncode != 0 and "no CUDA device" in (p.stderr + p.stdout)
