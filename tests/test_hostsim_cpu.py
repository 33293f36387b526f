"""Threading of the host mirror (cuda-audio_b200/host: Convolution, SharedEngine) without a GPU: the mirror's sources
are compiled with g++ against a TEST DOUBLE of the C ABI and of the CUDA runtime (tests/hostsim/fake_engine.cpp) and
driven by K threads through the in-process JACK stand-in, like K JACK clients (main.cu:31-39).  The double checks the
caller contract of include/cuda_audio_b200.h (no overlapping ca_process / ca_load_ir / ca_destroy on one engine, no
call on a destroyed engine, indices in range); the scenarios check that every period's output is exact or accounted
silence.  Built plain, with ThreadSanitizer and with AddressSanitizer + UBSan.

Scenarios (tests/hostsim/sim_main.cpp): steady `engine.shared` batch; prepare() and buildNow() racing a running
batch (rebuild at the rendezvous); a member that stops calling and comes back; a member destroyed mid-run; engine
creation failing and recovering; one Convolution of its own with prepare() against its running callback."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIM = os.path.join(ROOT, "tests", "hostsim")

pytestmark = pytest.mark.skipif(shutil.which("g++") is None or shutil.which("make") is None or not os.path.exists("/usr/local/cuda/include/cuda_runtime.h"),
                                reason="needs g++, make and the CUDA headers")


def _build(target):
    r = subprocess.run(["make", "-C", SIM, target], capture_output=True, text=True)
    return r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def _run(exe, periods, members, env=None):
    for attempt in range(2):
        r = subprocess.run([os.path.join(SIM, exe), str(periods), str(members)], capture_output=True, text=True, timeout=600, env=dict(os.environ, **(env or {})))
        lines = [ln for ln in (r.stdout + r.stderr).splitlines() if not ln.startswith("[")]   # the mirror's log lines
        tail = "\n".join(lines[-60:])
        # sanitizer reports and contract violations are never retried; the scenarios' own bounds (elapsed time, periods
        # lost around a rebuild) get one second chance on a machine that is busy with something else
        assert "ThreadSanitizer" not in r.stderr and "AddressSanitizer" not in r.stderr and "runtime error" not in r.stderr and "VIOLATION" not in r.stderr, tail
        if r.returncode == 0 and "HOSTSIM OK" in r.stdout:
            return
    raise AssertionError(tail)


def test_host_mirror_threading_plain():
    ok, log = _build("hostsim_plain")
    assert ok, log
    _run("hostsim_plain", 3000, 8)
    _run("hostsim_plain", 2000, 3)


def test_host_mirror_threading_under_thread_sanitizer():
    ok, log = _build("hostsim_tsan")
    if not ok:
        pytest.skip("no ThreadSanitizer runtime for this g++: " + log[-300:])
    _run("hostsim_tsan", 1500, 6, env={"TSAN_OPTIONS": "halt_on_error=0 second_deadlock_stack=1"})


def test_host_mirror_threading_under_address_sanitizer():
    ok, log = _build("hostsim_asan")
    if not ok:
        pytest.skip("no AddressSanitizer runtime for this g++: " + log[-300:])
    _run("hostsim_asan", 1500, 6)


def test_parameter_ring_of_the_c_abi_under_thread_sanitizer():
    """csrc/param_queue.h (behind ca_set_params / ca_set_glide, "lock-free from any thread"): six producers
    against the draining thread: nothing lost, duplicated, reordered per producer or torn; a full ring refuses."""
    ok, log = _build("paramq_tsan")
    assert ok, log
    r = subprocess.run([os.path.join(SIM, "paramq_tsan"), "6", "100000"], capture_output=True, text=True, timeout=600)
    assert "ThreadSanitizer" not in r.stderr, r.stderr[-3000:]
    assert r.returncode == 0 and "PARAMQ OK" in r.stdout, r.stdout + r.stderr[-2000:]
