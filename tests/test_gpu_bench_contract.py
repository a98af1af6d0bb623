"""bench.py prints ONE JSON line with the round contract's keys for every `--config` (tiny step counts; the numbers
themselves are not asserted here — only that each arm runs on the GPU through the product path and reports completely)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"}


def _run(*args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=900,
                       cwd=ROOT)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and len(lines) == 1, r.stdout[-1500:] + r.stderr[-1500:]
    return json.loads(lines[0])


@pytest.mark.parametrize("cfg,extra", [("vae_train", ["--batch", "8"]), ("unet_train", ["--batch", "2"]),
                                       ("inference", ["--batch", "8"]), ("voxeliser", ["--batch", "64"])])
def test_bench_line_contract(cfg, extra):
    line = _run("--config", cfg, "--steps", "3", "--warmup", "3", "--no-cpu-baseline", *extra)
    assert KEYS <= set(line), KEYS - set(line)
    assert line["value"] > 0 and line["gpu_launches"] > 0 and line["n_gpus"] == 1 and line["higher_is_better"] is True
    assert "workload" in line["config"] and "model" not in line["config"]
    e = line["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    ro = line["roofline"]
    assert ro["bound"] in ("hbm", "tensor") and 0 < ro["frac"] < 1 and ro["peak"] > 0 and ro["achieved"] > 0
    assert line["clocks"]["sm_max_mhz"] and "reasons" in line["clocks"]
