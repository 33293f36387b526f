"""The reference's own executable source (src/main.cu, unmodified) compiles and links against this
engine's host mirror: compat headers named like the reference's + libca_host.a + libcuda_audio_b200.so
(INTEGRATION.md section A).  Needs /root/reference (this container only); the source is copied to a
temporary directory because its quoted includes would otherwise find the reference's own headers."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "cuda-audio_b200")
REF_MAIN = "/root/reference/src/main.cu"


@pytest.mark.skipif(not os.path.exists(REF_MAIN) or shutil.which("nvcc") is None, reason="needs the reference checkout and nvcc")
def test_reference_main_cu_builds_and_starts_against_host_mirror(tmp_path):
    subprocess.check_call(["make", "-C", PKG], stdout=subprocess.DEVNULL)
    src = tmp_path / "src"
    stub = tmp_path / "stub"
    src.mkdir()
    stub.mkdir()
    shutil.copy(REF_MAIN, src / "main.cu")
    (stub / "ncurses.h").write_text("/* main.cu includes ncurses.h and uses nothing from it */\n")
    exe = tmp_path / "cuda-audio"
    cmd = ["nvcc", "-std=c++17", "-Wno-deprecated-gpu-targets", "-gencode", "arch=compute_100a,code=sm_100a",
           "-I", os.path.join(PKG, "host", "compat"), "-I", str(stub), "-o", str(exe), str(src / "main.cu"),
           os.path.join(PKG, "host", "jack_dl.cpp"), os.path.join(PKG, "host", "libca_host.a"),
           "-L", PKG, "-lcuda_audio_b200", "-ldl", "-Xlinker", "-rpath," + PKG]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    # main.cu: selectGpu(); settings.open("settings.txt"); conv.count/2 instances; waits for Enter
    (tmp_path / "settings.txt").write_text("# no instances\nconv.count 0\n")
    r = subprocess.run([str(exe)], cwd=tmp_path, input="\n", capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, (r.stdout, r.stderr)
