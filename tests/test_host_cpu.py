"""Host-side logic of the C++ mirror that needs no GPU: settings.txt parser and WAV decoder
(cuda-audio_b200/host/), checked through ca_render's dump modes against Python restatements of
the reference's behaviour (settings.cu:4-24, wav.cu:46-118)."""
import json
import os
import struct
import subprocess

import numpy as np
import pytest

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RENDER = os.path.join(ROOT, "cuda-audio_b200", "host", "ca_render")


@pytest.fixture(scope="module", autouse=True)
def built():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "cuda-audio_b200")], stdout=subprocess.DEVNULL)
    assert os.path.exists(RENDER)


def ref_settings_parse(text):
    """settings.cu:4-24 restated: token based; a key starting with '#' drops the rest of its line."""
    out, i, n = {}, 0, len(text)

    def token():
        nonlocal i
        while i < n and text[i].isspace():
            i += 1
        j = i
        while i < n and not text[i].isspace():
            i += 1
        return text[j:i]

    while True:
        key = token()
        if not key:
            break
        if key[0] == "#":
            while i < n and text[i] != "\n":
                i += 1
            continue
        out[key] = token()
    return out


SETTINGS = """# MY CONVOLUTION SETTINGS
#-------------------------
conv.count 2

# left
conv[0].fftSize \t131072\t
conv[0].maxPredelay\t8192
conv[0].index\t\t./ir/all.index
conv[0].input\t\tsystem:capture_1   # trailing comment
conv[0].cc.device\thw:2,0
conv[0].cc.message\t176
conv[0].value.wet\t0.5
conv[0].value.panDry \t-0.25
conv[1].fftSize
   131072
#conv[1].disabled yes
conv[1].flag yes
"""


def test_settings_parser_matches_reference_semantics(tmp_path):
    p = tmp_path / "settings.txt"
    p.write_text(SETTINGS)
    out = subprocess.check_output([RENDER, "--dump-settings", str(p)], text=True)
    got = dict(line.split("=", 1) for line in out.strip().splitlines())
    want = ref_settings_parse(SETTINGS)
    assert got == want
    assert got["conv[1].fftSize"] == "131072" and got["conv[0].input"] == "system:capture_1"
    assert "#conv[1].disabled" not in got and got["conv[1].flag"] == "yes"


def wav_bytes(fmt_tag, channels, rate, bits, data, extra_before=b"", extra_after=b"", fmt_extra=b""):
    block = channels * bits // 8
    fmt = struct.pack("<HHIIHH", fmt_tag, channels, rate, rate * block, block, bits) + fmt_extra
    body = b"WAVE" + extra_before + b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"data" + struct.pack("<I", len(data)) + data
    if len(data) & 1:
        body += b"\0"
    body += extra_after
    return b"RIFF" + struct.pack("<I", len(body)) + body


def dump(path, scale, tmp_path):
    outp = tmp_path / "dump.f32"
    meta = json.loads(subprocess.check_output([RENDER, "--dump-wav", str(path), str(scale), str(outp)], text=True))
    data = np.fromfile(outp, np.float32).reshape(meta["channels"], meta["frames"])
    return meta, data


def test_wav_pcm16_pcm24_half_scale_bit_exact(tmp_path):
    """IR convention of the reference: int16 / 65536, int24 / 2^24 (wav.cu:13-14, 24-41)."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "ref_W.npz"))
    p16, p24 = tmp_path / "a.wav", tmp_path / "b.wav"
    p16.write_bytes(wav_bytes(1, 2, 44100, 16, z["pcm16"].tobytes()))
    # trailing LIST + cue chunks after data, like 3 of the reference's IR files
    p24.write_bytes(wav_bytes(1, 2, 44100, 24, z["raw24"].tobytes(), extra_after=b"cue " + struct.pack("<I", 4) + b"\0\0\0\0" + b"LIST" + struct.pack("<I", 4) + b"INFO"))
    m16, d16 = dump(p16, 0.5, tmp_path)
    assert (m16["channels"], m16["rate"], m16["bits"], m16["frames"]) == (2, 44100, 16, 777)
    assert np.array_equal(d16[0], z["l16"]) and np.array_equal(d16[1], z["r16"])   # == the real reference's WavFile output
    m24, d24 = dump(p24, 0.5, tmp_path)
    assert m24["frames"] == 555
    assert np.array_equal(d24[0], z["l24"]) and np.array_equal(d24[1], z["r24"])
    assert np.array_equal(d16.T.reshape(-1), O.pcm16_to_float(z["pcm16"]))            # and the C restatement


def test_wav_robust_chunk_walk_mono_float_and_extensible(tmp_path):
    rng = np.random.default_rng(0)
    x = rng.standard_normal(101).astype(np.float32)
    # LIST chunk BEFORE fmt (the reference would misparse this: wav.cu:71 assumes fmt comes first), odd-sized chunk padding
    p = tmp_path / "f.wav"
    p.write_bytes(wav_bytes(3, 1, 48000, 32, x.tobytes(), extra_before=b"LIST" + struct.pack("<I", 5) + b"INFOx\0"))
    m, d = dump(p, 1.0, tmp_path)
    assert (m["channels"], m["format"], m["frames"]) == (1, 3, 101) and np.array_equal(d[0], x)
    # WAVE_FORMAT_EXTENSIBLE PCM16
    pcm = rng.integers(-32768, 32768, 2 * 50).astype(np.int16)
    ext = struct.pack("<HHI", 22, 16, 3) + struct.pack("<H", 1) + b"\x00\x00\x00\x00\x10\x00\x80\x00\x00\xaa\x00\x38\x9b\x71"
    p2 = tmp_path / "e.wav"
    p2.write_bytes(wav_bytes(0xFFFE, 2, 96000, 16, pcm.tobytes(), fmt_extra=ext))
    m, d = dump(p2, 1.0, tmp_path)
    assert (m["channels"], m["rate"], m["frames"], m["format"]) == (2, 96000, 50, 1)
    assert np.array_equal(d[0], pcm[0::2] / np.float32(32768))
    # truncated data chunk (header claims more than the file holds) and garbage files
    raw = wav_bytes(1, 2, 44100, 16, pcm.tobytes())
    p3 = tmp_path / "t.wav"
    p3.write_bytes(raw[:-40])
    m, _ = dump(p3, 1.0, tmp_path)
    assert m["frames"] == 40
    bad = tmp_path / "bad.wav"
    bad.write_bytes(b"not a wav file at all")
    r = subprocess.run([RENDER, "--dump-wav", str(bad), "1", str(tmp_path / "x")], capture_output=True, text=True)
    assert r.returncode == 1 and "not a RIFF/WAVE" in r.stderr


def test_render_fails_loudly_without_gpu(tmp_path):
    from tests import conftest
    if conftest._has_gpu():
        pytest.skip("GPU present")
    r = subprocess.run([RENDER, "--synthetic-ir", "0.05", "--synthetic-in", "0.1", "--out", str(tmp_path / "o.wav")], capture_output=True, text=True)
    assert r.returncode != 0 and not (tmp_path / "o.wav").exists()


def test_live_executable_fails_cleanly_without_libjack(tmp_path):
    """ca_live = the reference's main.cu flow with libjack resolved at run time: without libjack /
    jackd it must report the problem and exit non-zero (the reference asserts)."""
    live = os.path.join(ROOT, "cuda-audio_b200", "host", "ca_live")
    assert os.path.exists(live)
    lines = ["conv.count 2"]
    for i in range(2):
        lines += [f"conv[{i}].fftSize 8192", f"conv[{i}].index {tmp_path}/none.index", f"conv[{i}].input system:capture_{i + 1}",
                  f"conv[{i}].output system:playback_{i + 1}"]
        lines += [f"conv[{i}].value.{k} 0" for k in ("select", "predelay", "speed")]
        lines += [f"conv[{i}].value.{k} 0.5" for k in ("dry", "wet", "panDry", "panWet", "level")]
    (tmp_path / "settings.txt").write_text("\n".join(lines) + "\n")
    r = subprocess.run([live, str(tmp_path / "settings.txt")], capture_output=True, text=True, stdin=subprocess.DEVNULL)
    assert r.returncode == 1
    assert "cannot start JACK client" in r.stderr or "cannot open JACK client" in r.stderr
    r = subprocess.run([live, str(tmp_path / "missing.txt")], capture_output=True, text=True, stdin=subprocess.DEVNULL)
    assert r.returncode == 1 and "cannot open settings file" in r.stderr
