"""Drop-in proof on the GPU: the SAME headless driver source (oracle/ref_harness/harness.cu) is
compiled once against the unmodified reference (oracle/_ref/libref_conv.so) and once against the
B200 engine's host mirror (tests/dropin/libdropin_conv.so).  One Python driver runs both through
the reference's own public API -- Convolution(name, fftSize), prepare(), onProcess(), cc[].value,
onMidiMessage(), WavFile -- and the outputs must agree to <= 1e-5 relative L2."""
import importlib.util
import json
import os
import struct
import subprocess
import sys

import numpy as np
import pytest

from oracle import oracle as O
from oracle import refgpu

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "tests", "dropin", "libdropin_conv.so")
RENDER = os.path.join(ROOT, "cuda-audio_b200", "host", "ca_render")
TOL_REF = 1e-5

spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
mg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mg)

needs_libs = pytest.mark.skipif(not (refgpu.available() and os.path.exists(DROPIN)), reason="harness libraries not built")


def drive(lib_path, name):
    """tests/golden/make_golden.py's driver, parametrised by the library it talks to."""
    N, B, irs, x, cc, events = mg.case_inputs(name)
    c = refgpu.RefGpu(N, 0, lib_path)
    for s, pair in enumerate(irs):
        c.prepare(s, pair[0], pair[1], B)
    for i in range(2):
        c.set_cc(i, **cc[i])
    ev = {}
    for e in events:
        ev.setdefault(e[0], []).append(e)
    periods = x.shape[1] // B
    L = np.empty(periods * B, np.float32)
    R = np.empty(periods * B, np.float32)
    trace = []
    for t in range(periods):
        for (_, inp, field, val) in ev.get(t, []):
            c.midi_cc(inp, field, val)
        l, r = c.process(x[0, t * B:(t + 1) * B], x[1, t * B:(t + 1) * B])
        L[t * B:(t + 1) * B] = l
        R[t * B:(t + 1) * B] = r
        g0, g1 = c.get_cc(0), c.get_cc(1)
        trace.append([g0["select"], g0["vsteps"], g1["select"], g1["vsteps"]])
    return L, R, np.array(trace, np.int64), c


@needs_libs
@pytest.mark.parametrize("name", ["A", "B"])
def test_same_driver_same_output(name):
    rl, rr, rt, _ = drive(None, name)
    ml, mr, mt, _ = drive(DROPIN, name)
    assert np.array_equal(rt, mt)
    z = np.load(os.path.join(ROOT, "tests", "golden", f"ref_{name}.npz"))
    assert O.rel_l2(rl, z["L"]) < 1e-6          # the live reference reproduces its committed golden
    eL, eR = O.rel_l2(ml, rl), O.rel_l2(mr, rr)
    assert eL < TOL_REF and eR < TOL_REF, (name, eL, eR)


@needs_libs
def test_class_api_with_ref_quirks_matches_reference_on_unconstrained_irs():
    """`engine.ref_quirks` (CA_ENGINE_REF_QUIRKS): through the reference's own class API the mirror reproduces conv.cu's
    DC / Nyquist bins for the object's fftSize, so raw noise IRs match too.  Own process: the default engine options are
    read from the environment once."""
    code = (
        "import sys, json, numpy as np\n"
        "sys.path.insert(0, %r)\n"
        "import importlib.util\n"
        "spec = importlib.util.spec_from_file_location('t', %r); t = importlib.util.module_from_spec(spec); spec.loader.exec_module(t)\n"
        "from oracle import oracle as O\n"
        "rl, rr, _, _ = t.drive(None, 'E')\n"
        "ml, mr, _, _ = t.drive(t.DROPIN, 'E')\n"
        "print(json.dumps([O.rel_l2(ml, rl), O.rel_l2(mr, rr), float(max(abs(rl).max(), abs(rr).max()))]))\n"
    ) % (ROOT, os.path.abspath(__file__))
    res = {}
    for quirks in ("0", "1"):
        env = dict(os.environ, CA_ENGINE_REF_QUIRKS=quirks)
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        res[quirks] = json.loads(out.stdout.strip().split("\n")[-1])
    assert res["1"][2] < 0.99                                   # below the clamp
    assert max(res["0"][:2]) > 1e-4, res                        # exact engine: differs by the reference's DC / Nyquist terms
    assert max(res["1"][:2]) < TOL_REF, res


@needs_libs
def test_midi_cc_state_machine_matches():
    """handleCC (conv.cu:255-276): select/vsteps/speed bookkeeping and the per-period countdown
    (conv.cu:345,353) observed through cc[].value, identical on both implementations."""
    _, _, rt, rc = drive(None, "D")
    _, _, mt, mc = drive(DROPIN, "D")
    assert np.array_equal(rt, mt)
    for i in range(2):
        a, b = rc.get_cc(i), mc.get_cc(i)
        assert a == b, (a, b)
    for field, val in ((2, 77), (3, 32), (4, 100), (5, 9), (6, 0), (7, 127), (8, 64)):
        rc.midi_cc(0, field, val)
        mc.midi_cc(0, field, val)
    assert rc.get_cc(0) == mc.get_cc(0)


@needs_libs
def test_ir_switch_crossfade_matches_reference():
    """Case D: IR `select` changes mid-run through the MIDI handler; the reference glides its live
    IR spectrum toward the new IR over ~speed periods (conv.cu:15-32, 339-353)."""
    rl, rr, _, _ = drive(None, "D")
    ml, mr, _, _ = drive(DROPIN, "D")
    eL, eR = O.rel_l2(ml, rl), O.rel_l2(mr, rr)
    assert eL < TOL_REF and eR < TOL_REF, (eL, eR)


@needs_libs
def test_wavfile_decode_matches_reference(tmp_path):
    pcm16, raw24, p16, p24 = mg.wav_cases(str(tmp_path))
    for p in (p16, p24):
        a = refgpu.wav_decode(p)
        b = refgpu.wav_decode(p, lib_path=DROPIN)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def write_f32_wav(path, planar, rate):
    planar = np.asarray(planar, np.float32)
    ch, n = planar.shape
    data = planar.T.astype("<f4").tobytes()
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVE")
        f.write(b"fmt " + struct.pack("<IHHIIHH", 16, 3, ch, rate, rate * 4 * ch, 4 * ch, 32))
        f.write(b"data" + struct.pack("<I", len(data)) + data)


def read_f32_wav(path):
    raw = open(path, "rb").read()
    ch = struct.unpack_from("<H", raw, 22)[0]
    n = struct.unpack_from("<I", raw, 40)[0]
    return np.frombuffer(raw, "<f4", n // 4, 44).reshape(-1, ch).T


def test_cfg1_headless_render_vs_fp64(tmp_path):
    """BASELINE configs[0]: mono 44.1 kHz, 256-frame period, 1 s synthetic IR, offline headless
    wav render (C++ Convolution mirror + in-process JACK) checked against the FP64 oracle."""
    fs, B, L = 44100, 256, 44100
    h = O.synth_ir(L, fs, 1000)
    x = O.synth_audio(fs * 3 + 123, 2000)      # ragged length: last period partly filled
    write_f32_wav(tmp_path / "ir.wav", np.stack([h, h]), fs)
    write_f32_wav(tmp_path / "in.wav", x[None], fs)
    out = subprocess.check_output([RENDER, "--ir", str(tmp_path / "ir.wav"), "--in", str(tmp_path / "in.wav"), "--out", str(tmp_path / "out.wav"),
                                   "--mono", "--wet", "1", "--dry", "0", "--period", str(B), "--warmup", "100"], text=True)
    st = json.loads(out.strip().splitlines()[-1])
    assert st["period"] == B and st["rate"] == fs and st["frames"] == len(x)
    y = read_f32_wav(tmp_path / "out.wav")[0]
    truth = O.fft_conv(x, 0.5 * h.astype(np.float64))   # IR wavs load at the reference's half scale
    assert len(y) == len(x)
    assert O.rel_l2(y, truth) < 1e-4
    assert O.rel_l2(y, truth) < 5e-6
    idx = np.random.default_rng(0).integers(0, len(x), 400)
    assert O.rel_l2(y[idx], O.direct_conv_at(x, 0.5 * h.astype(np.float64), idx)) < 1e-4


def test_settings_driven_render_two_instances(tmp_path):
    """main.cu's flow: conv.count 4 = two true-stereo instances, IRs from index files, initial
    values from settings.txt keys, 4 input channels -> 4 output channels."""
    fs, B, L = 48000, 128, 3000
    irs = [O.synth_ir(L, fs, 300 + i) for i in range(4)]
    for j in range(2):
        write_f32_wav(tmp_path / f"ir{j}.wav", np.stack([irs[2 * j], irs[2 * j + 1]]), fs)
    (tmp_path / "all.index").write_text(f"{tmp_path}/ir0.wav\n{tmp_path}/ir1.wav\n")
    x = np.stack([O.synth_audio(B * 100, 400 + c) for c in range(4)])
    write_f32_wav(tmp_path / "in.wav", x, fs)
    lines = ["conv.count 4"]
    vals = [dict(select=0, wet=1.0, dry=0.25, panWet=0.5, panDry=0.0, level=1.0), dict(select=1, wet=0.5, dry=0.0, panWet=0.0, panDry=-0.5, level=0.5),
            dict(select=1, wet=1.0, dry=0.0, panWet=0.0, panDry=0.0, level=1.0), dict(select=0, wet=0.25, dry=1.0, panWet=-1.0, panDry=1.0, level=1.0)]
    for i, v in enumerate(vals):
        lines += [f"conv[{i}].fftSize 8192", f"conv[{i}].maxPredelay 8192", f"conv[{i}].index {tmp_path}/all.index",
                  f"conv[{i}].input system:capture_{i + 1}", f"conv[{i}].output system:playback_{i + 1}", f"conv[{i}].cc.device hw:9,9",
                  f"conv[{i}].cc.message 176"]
        lines += [f"conv[{i}].cc.{k} {20 + n}" for n, k in enumerate(["select", "predelay", "dry", "wet", "speed", "panDry", "panWet", "level"])]
        lines += [f"conv[{i}].value.select {v['select']}", f"conv[{i}].value.predelay 0", f"conv[{i}].value.dry {v['dry']}", f"conv[{i}].value.wet {v['wet']}",
                  f"conv[{i}].value.speed 100", f"conv[{i}].value.panDry {v['panDry']}", f"conv[{i}].value.panWet {v['panWet']}", f"conv[{i}].value.level {v['level']}"]
    (tmp_path / "settings.txt").write_text("# test settings\n" + "\n".join(lines) + "\n")
    subprocess.check_call([RENDER, "--settings", str(tmp_path / "settings.txt"), "--in", str(tmp_path / "in.wav"), "--out", str(tmp_path / "out.wav"),
                           "--period", str(B), "--warmup", "100"], stdout=subprocess.DEVNULL)
    y = read_f32_wav(tmp_path / "out.wav")
    assert y.shape == (4, B * 100)
    stereo = [[0.5 * irs[0], 0.5 * irs[1]], [0.5 * irs[2], 0.5 * irs[3]]]  # bank: idx 0, 1 (half scale)
    for n in range(2):
        pr = [vals[2 * n], vals[2 * n + 1]]
        truth = O.engine_truth(x[2 * n:2 * n + 2], [stereo[pr[0]["select"]], stereo[pr[1]["select"]]], pr)
        for o in range(2):
            assert O.rel_l2(y[2 * n + o], truth[o]) < 5e-6, (n, o)


@needs_libs
def test_shared_batched_engine_through_the_class_api():
    """`engine.shared` / CA_ENGINE_SHARED: three mirror Convolution objects on three host threads are the three
    instances of ONE batched, tiered engine (one set of launches per cycle); same driver code as the
    reference arm (harness.cu), results against fp64."""
    import sys
    env = dict(os.environ, CA_ENGINE_SHARED="3", CA_ENGINE_TIERS="auto")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dropin_shared_worker.py")], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    res = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("SHARED_RESULT ")][-1].split(" ", 1)[1])
    assert res["K"] == 3 and max(res["rel_l2"]) < 5e-6, res


@needs_libs
def test_shared_batched_engine_one_period_late_without_rendezvous():
    """`engine.shared_latency 1` / CA_ENGINE_SHARED_LATENCY: the same three objects on three threads hand their block in
    and take the previous period's out -- nobody waits inside a cycle (hosts that call their clients one after the other);
    every object's output is the fp64 result delayed by exactly one period."""
    import sys
    env = dict(os.environ, CA_ENGINE_SHARED="3", CA_ENGINE_TIERS="auto", CA_ENGINE_SHARED_LATENCY="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dropin_shared_worker.py")], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    res = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("SHARED_RESULT ")][-1].split(" ", 1)[1])
    assert res["K"] == 3 and max(res["rel_l2"]) < 5e-6, res


@needs_libs
@pytest.mark.parametrize("G", [1, 2])
def test_ir_split_group_through_the_class_api(G):
    """`engine.ir_split` / CA_ENGINE_IR_SPLIT: the engine behind ONE mirror Convolution object is a ca_group over G GPUs
    (BASELINE configs[4]; G = 1 runs the same code path on one GPU), same prepare / onProcess surface, results
    against fp64, including a prepare() that rebuilds the live group."""
    import sys
    import torch
    if torch.cuda.device_count() < G:
        pytest.skip("needs %d GPUs" % G)
    env = dict(os.environ, CA_ENGINE_IR_SPLIT=str(G))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dropin_irsplit_worker.py")], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    res = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("IRSPLIT_RESULT ")][-1].split(" ", 1)[1])
    assert res["G"] == G and res["rel_l2"] < 5e-6 and res["rel_l2_after_prepare"] < 5e-6, res


def test_render_with_engine_keys_shared_and_tiers(tmp_path):
    """settings.txt `engine.*` keys reach the engine: two instances as ONE shared, tiered, batched engine give
    the same audio as two private uniform engines."""
    fs, B, L = 48000, 64, 64 * 8 + 512 * 5 + 11
    irs = [O.synth_ir(L, fs, 330 + i) for i in range(4)]
    for j in range(2):
        write_f32_wav(tmp_path / f"ir{j}.wav", np.stack([irs[2 * j], irs[2 * j + 1]]), fs)
    (tmp_path / "all.index").write_text(f"{tmp_path}/ir0.wav\n{tmp_path}/ir1.wav\n")
    x = np.stack([O.synth_audio(B * 200, 430 + c) for c in range(4)])
    write_f32_wav(tmp_path / "in.wav", x, fs)

    def render(extra, name):
        lines = ["conv.count 4"] + extra
        for i in range(4):
            lines += [f"conv[{i}].fftSize 8192", f"conv[{i}].index {tmp_path}/all.index", f"conv[{i}].input system:capture_{i + 1}",
                      f"conv[{i}].output system:playback_{i + 1}", f"conv[{i}].cc.message 176"]
            lines += [f"conv[{i}].cc.{k} {20 + n}" for n, k in enumerate(["select", "predelay", "dry", "wet", "speed", "panDry", "panWet", "level"])]
            lines += [f"conv[{i}].value.select {i % 2}", f"conv[{i}].value.predelay {7 * i}", f"conv[{i}].value.dry 0.25", f"conv[{i}].value.wet 0.75",
                      f"conv[{i}].value.speed 100", f"conv[{i}].value.panDry 0.0", f"conv[{i}].value.panWet {0.25 * i - 0.25}", f"conv[{i}].value.level 1.0"]
        (tmp_path / f"{name}.txt").write_text("\n".join(lines) + "\n")
        subprocess.check_call([RENDER, "--settings", str(tmp_path / f"{name}.txt"), "--in", str(tmp_path / "in.wav"), "--out", str(tmp_path / f"{name}.wav"),
                               "--period", str(B), "--warmup", "100"], stdout=subprocess.DEVNULL)
        return read_f32_wav(tmp_path / f"{name}.wav")

    ya = render([], "private")
    yb = render(["engine.shared 2", "engine.tiers auto", "engine.period 64"], "shared")
    assert ya.shape == yb.shape == (4, B * 200)
    for c in range(4):
        assert O.rel_l2(yb[c], ya[c]) < 2e-6, (c, O.rel_l2(yb[c], ya[c]))
