"""bench.py on a box without a GPU: the product arm fails loudly (no CPU fallback, nothing printed that could be read as
a measurement), the reference arm times the CPU port of the path (the one place bench.py may execute oracle/) and prints
the contract's JSON line."""
import json
import os
import subprocess
import sys

import pytest

from tests import conftest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(conftest._has_gpu(), reason="GPU present: the driver runs the real bench")


def test_product_arm_refuses_to_run_without_a_gpu():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stdout + r.stderr)
    assert not any(ln.startswith("{") for ln in r.stdout.splitlines())


def test_reference_arm_prints_the_contract_line_from_the_cpu_port():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1", "--cpu-seconds", "2"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    j = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert j["impl"] == "reference" and j["unit"] == "rt_channels" and j["higher_is_better"] is True and j["value"] > 0
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["sample"]
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0 and j["e2e"]["value"] == j["value"]
    assert "workload" in j["config"]
