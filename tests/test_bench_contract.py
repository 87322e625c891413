"""bench.py's reference arm on the CPU (no GPU needed): the line the driver reads must describe what actually ran."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_reports_what_it_ran():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ref_binding as rb
    env = dict(os.environ, UCGB200_NCELL_PER_GPU="8")     # 2048 sites instead of 1 000 188: seconds instead of a minute
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "1", "--serial"],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "UCG-LD Matom-steps/s" and line["higher_is_better"] is True
    assert line["config"]["sites"] == 4 * 8 ** 3                       # the configuration named is the one that ran
    assert line["steps"] == 3 and line["steps_requested"] == 3
    cb = line["cpu_baseline"]
    assert cb["kind"] == ("reference" if rb.available() else "port") and cb["cores"] == 1
    assert cb["serial_1M"]["sites"] == line["config"]["sites"] and cb["serial_1M"]["value"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["value"] > 0 and abs(line["value"] - cb["value"]) < 1e-12


def test_gpu_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode != 0 and "no CUDA device" in (out.stderr + out.stdout)
