"""SURVEY 8(f) rank 2, the real JACK adapter: `ca_live` (the reference's main.cu flow on this engine) binds libjack at
run time (host/jack_dl.cpp).  Here it runs against tests/fakejack/libjack.so.0 -- a libjack + jackd stand-in whose
"server" thread calls the registered process callback once per period like jackd's real-time thread -- with a
settings file, IR index files and 16-bit IR wavs on disk, exactly the inputs of the reference's executable
(main.cu:18-116, settings.txt).  What comes out of the playback ports is compared with the engine driven through the
C ABI on the same data and with the fp64 convolution."""
import json
import os
import struct
import subprocess
import time

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIVE = os.path.join(ROOT, "cuda-audio_b200", "host", "ca_live")
FAKE = os.path.join(ROOT, "tests", "fakejack")


def ca():
    import cuda_audio_b200 as m
    return m


def write_wav16(path, left, right, rate=48000):
    pcm = np.stack([left, right], axis=1)
    data = np.clip(np.round(pcm * 32768.0), -32768, 32767).astype("<i2").tobytes()
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVE")
        f.write(b"fmt " + struct.pack("<IHHIIHH", 16, 1, 2, rate, rate * 4, 4, 16))
        f.write(b"data" + struct.pack("<I", len(data)) + data)
    return np.frombuffer(data, "<i2").reshape(-1, 2).astype(np.float32) / 65536.0   # the reference's half-scale decode, wav.cu:13-14


def run_live(tmp_path, settings_lines, x, B, pace_us=0, timeout=120):
    (tmp_path / "settings.txt").write_text("\n".join(settings_lines) + "\n")
    x.astype(np.float32).tofile(tmp_path / "in.f32")
    outp = tmp_path / "out.f32"
    env = dict(os.environ, LD_LIBRARY_PATH=FAKE + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""), FAKEJACK_IN=str(tmp_path / "in.f32"),
               FAKEJACK_OUT=str(outp), FAKEJACK_NFRAMES=str(B), FAKEJACK_RATE="48000", FAKEJACK_PERIOD_US=str(pace_us))
    p = subprocess.Popen([LIVE, str(tmp_path / "settings.txt")], env=env, stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    done = str(outp) + ".done"
    t0 = time.time()
    while not os.path.exists(done) and p.poll() is None and time.time() - t0 < timeout:
        time.sleep(0.05)
    try:
        so, se = p.communicate("\n", timeout=30)      # Enter: main.cu:95
    except subprocess.TimeoutExpired:
        p.kill()
        so, se = p.communicate()
    assert os.path.exists(done), (p.returncode, so[-500:], se[-2000:])
    assert p.returncode == 0, (so[-500:], se[-2000:])
    stats = json.loads(open(done).read())
    y = np.fromfile(outp, np.float32).reshape(2, -1)
    return y, stats, so


def settings_for(tmp_path, fft, values, extra=()):
    lines = ["conv.count 2"] + list(extra)
    for i in range(2):
        lines += [f"conv[{i}].fftSize {fft}", f"conv[{i}].index {tmp_path}/ir{i}.index", f"conv[{i}].input system:capture_{i + 1}",
                  f"conv[{i}].output system:playback_{i + 1}"]
        lines += [f"conv[{i}].value.{k} {v}" for k, v in values[i].items()]
    return lines


@pytest.mark.skipif(not (os.path.exists(LIVE) and os.path.exists(os.path.join(FAKE, "libjack.so.0"))), reason="ca_live / fake libjack not built")
@pytest.mark.parametrize("tiers", [None, "auto"])
def test_live_executable_through_libjack_binding(tmp_path, tiers):
    m = ca()
    fs, B, L, fft = 48000, 256, 6000, 8192
    irs = [[O.synth_ir(L, fs, 900 + 2 * i + o) for o in range(2)] for i in range(2)]
    dec = []
    for i in range(2):
        d = write_wav16(tmp_path / f"ir{i}.wav", irs[i][0] * 0.9, irs[i][1] * 0.9)   # what every decoder must see: v / 65536
        dec.append(d)
        # input i selects line `select` of ITS index file (main.cu:72-80): both files list both IRs in the same order
    for i in range(2):
        (tmp_path / f"ir{i}.index").write_text(f"{tmp_path}/ir0.wav\n{tmp_path}/ir1.wav\n")
    periods = 300
    x = np.stack([O.synth_audio(B * periods, 2500 + i) for i in range(2)])
    vals = [dict(select=0, predelay=120, speed=100, dry=0.25, wet=0.75, panDry=-0.25, panWet=0.5, level=0.75),
            dict(select=1, predelay=0, speed=100, dry=0.5, wet=0.5, panDry=0.5, panWet=-0.25, level=1.0)]
    extra = ["engine.tiers auto", "engine.tier_growth 4"] if tiers else []
    y, stats, so = run_live(tmp_path, settings_for(tmp_path, fft, vals, extra), x, B)
    assert stats["periods"] == periods and "average convolution runtime" in so
    # the same through the C ABI (ctypes): IRs as decoded, parameters as in the settings file, glide from silence
    kw = dict(tiers="auto", tier_growth=4) if tiers else {}
    with m.Engine(period=B, max_ir_frames=L, max_voices=3, flags=m.FLAG_GRAPH, **kw) as e:
        for i in range(2):
            e.load_ir(i, dec[i][:, 0].copy(), dec[i][:, 1].copy())
        for i in range(2):
            e.set_params(0, i, **vals[i])
        yy = e.render(x[None])[0]
    for o in range(2):
        assert O.rel_l2(y[o], yy[o]) < 1e-6, (o, O.rel_l2(y[o], yy[o]))
    # and against fp64 once the fade-in glide has converged
    pr = [dict(wet=v["wet"], dry=v["dry"], level=v["level"], panWet=v["panWet"], panDry=v["panDry"]) for v in vals]
    truth = O.engine_truth(x, [[dec[i][:, o] for o in range(2)] for i in range(2)], pr, predelay=120)
    sl = slice(150 * B, None)
    for o in range(2):
        assert O.rel_l2(y[o][sl], truth[o][sl]) < 1e-4, (o, O.rel_l2(y[o][sl], truth[o][sl]))


@pytest.mark.skipif(not (os.path.exists(LIVE) and os.path.exists(os.path.join(FAKE, "libjack.so.0"))), reason="ca_live / fake libjack not built")
def test_live_callback_meets_the_deadline_when_paced(tmp_path):
    """the callback as jackd would call it: one period every 5.33 ms; no callback may take longer than the period"""
    fs, B, L = 48000, 256, 48000
    irs = [[O.synth_ir(L, fs, 950 + 2 * i + o) for o in range(2)] for i in range(2)]
    for i in range(2):
        write_wav16(tmp_path / f"ir{i}.wav", irs[i][0], irs[i][1])
        (tmp_path / f"ir{i}.index").write_text(f"{tmp_path}/ir{i}.wav\n")
    x = np.stack([O.synth_audio(B * 150, 2600 + i) for i in range(2)])
    vals = [dict(select=0, predelay=0, speed=100, dry=0.5, wet=0.5, panDry=0.0, panWet=0.0, level=1.0)] * 2
    y, stats, _ = run_live(tmp_path, settings_for(tmp_path, 65536, vals, ["engine.tiers auto"]), x, B, pace_us=5333)
    assert stats["periods"] == 150
    assert stats["max_us"] < 5333, stats          # includes the very first callback (the engine is built in onStart, not here)
    assert np.abs(y).max() > 0.01
