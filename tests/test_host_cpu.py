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


def test_resampler_44k1_to_48k_matches_analytic_sines(tmp_path):
    """wav_resample (opt-in IR sample-rate conversion): two sines below the input
    Nyquist come out as the same sines at the new rate, DC gain is exactly 1, length = ceil(n*160/147),
    and a tone above the output band of a DOWN-conversion is removed."""
    fs0, fs1, n = 44100, 48000, 44100
    t0 = np.arange(n) / fs0
    x = 0.4 * np.sin(2 * np.pi * 1000 * t0) + 0.3 * np.sin(2 * np.pi * 15000 * t0 + 0.5) + 0.1
    p = tmp_path / "a.wav"
    p.write_bytes(wav_bytes(3, 2, fs0, 32, np.stack([x, -x], axis=1).astype(np.float32).tobytes()))
    outp = tmp_path / "r.f32"
    meta = json.loads(subprocess.check_output([RENDER, "--resample-to", str(fs1), "--dump-wav", str(p), "1.0", str(outp)], text=True))
    assert meta["rate"] == fs1 and meta["frames"] == int(np.ceil(n * fs1 / fs0))
    y = np.fromfile(outp, np.float32).reshape(2, -1)
    t1 = np.arange(y.shape[1]) / fs1
    want = 0.4 * np.sin(2 * np.pi * 1000 * t1) + 0.3 * np.sin(2 * np.pi * 15000 * t1 + 0.5) + 0.1
    mid = slice(200, y.shape[1] - 200)  # away from the ends, where the kernel sees the file boundary
    assert np.max(np.abs(y[0, mid] - want[mid])) < 2e-5
    assert np.max(np.abs(y[1, mid] + want[mid])) < 2e-5
    # 48 k -> 32 k: a 20 kHz tone is above the new Nyquist and must vanish (> 90 dB down), 1 kHz stays
    t = np.arange(48000) / 48000
    z = 0.5 * np.sin(2 * np.pi * 20000 * t) + 0.25 * np.sin(2 * np.pi * 1000 * t)
    p2 = tmp_path / "b.wav"
    p2.write_bytes(wav_bytes(3, 1, 48000, 32, z.astype(np.float32).tobytes()))
    subprocess.check_output([RENDER, "--resample-to", "32000", "--dump-wav", str(p2), "1.0", str(outp)], text=True)
    d = np.fromfile(outp, np.float32)
    td = np.arange(d.size) / 32000
    resid = d[300:-300] - 0.25 * np.sin(2 * np.pi * 1000 * td[300:-300])
    assert np.max(np.abs(resid)) < 0.5 * 10 ** (-90 / 20)


def test_midi_stream_parser(tmp_path):
    """MidiParser (rawmidi.cpp): running status, real-time bytes inside a message, 2-byte messages,
    system common, SysEx kept whole, stray data ignored -- midi.cu:3-20,49-51 handles only the
    3-byte subset and asserts on the rest."""
    stream = bytes([0x40,                                  # data byte before any status: ignored
                    0xB0, 20, 64,                          # CC
                    21, 0xF8, 100,                         # running status, clock byte in the middle
                    0xC1, 5,                               # program change (2 bytes)
                    6,                                     # running status program change
                    0x90, 60, 0xFE, 127,                   # note on with active sensing inside
                    0xF0, 1, 2, 3, 0xF7,                   # SysEx
                    0xF2, 0x10, 0x20,                      # song position
                    7,                                     # running status was cancelled by system common: ignored
                    0xF6,                                  # tune request
                    0xE0, 0, 64,                           # pitch bend
                    0xF0, 9, 9, 0xB1, 22, 33])             # SysEx aborted by a status byte, CC follows
    p = tmp_path / "m.bin"
    p.write_bytes(stream)
    out = subprocess.check_output([RENDER, "--midi-parse", str(p)], text=True).split("\n")
    assert [l for l in out if l] == ["b0 14 40", "f8", "b0 15 64", "c1 05", "c1 06", "fe", "90 3c 7f", "f0 01 02 03 f7",
                                     "f2 10 20", "f6", "e0 00 40", "b1 16 21"]


def test_midi_device_thread_drives_cc_mapping(tmp_path):
    """RawMidi::Device on a FIFO standing in for /dev/snd/midiC*D*: the reader thread parses the byte
    stream and Convolution::onMidiMessage applies the reference's CC -> parameter mapping
    (conv.cu:255-276): dry/wet/level = v/128, pan = v/64 - 1, predelay = v*8192/128, speed = v*1024/128."""
    assert subprocess.check_output([RENDER, "--midi-parse", os.devnull], text=True) == ""
    fifo = tmp_path / "midi.fifo"
    os.mkfifo(fifo)
    proc = subprocess.Popen([RENDER, "--midi-listen", str(fifo), "1.0"], stdout=subprocess.PIPE, text=True)
    assert proc.stdout.readline().strip() == "listening"
    with open(fifo, "wb", buffering=0) as f:
        f.write(bytes([0xB0, 22, 32, 23, 96]))         # dry = 0.25, wet = 0.75 (running status)
        f.write(bytes([0xB0, 25, 0, 0xB0, 26, 127]))   # panDry = -1, panWet = 127/64 - 1
        f.write(bytes([0xB0, 21, 64, 24, 16, 27, 64])) # predelay 4096, speed 128, level 0.5
        f.write(bytes([0xB1, 22, 127]))                # other channel: ignored (cc.message is 0xB0)
    out = json.loads(proc.stdout.read().strip().split("\n")[-1])
    proc.wait(timeout=10)
    assert out == {"predelay": 4096, "dry": 0.25, "wet": 0.75, "speed": 128, "panDry": -1.0, "panWet": 127 / 64 - 1, "level": 0.5}
    assert RENDER and subprocess.run([RENDER, "--midi-listen", str(tmp_path / "nope"), "0.1"], capture_output=True).returncode == 1


def test_wav_with_lying_block_align_is_rejected_or_clamped(tmp_path):
    """A header whose blockAlign is smaller than channels x bytes per sample would walk the sample loop past the
    data chunk: rejected.  A larger blockAlign (padding per frame) is legal: the last frame must still lie inside."""
    pcm = (np.arange(40, dtype=np.int16) * 100).tobytes()              # 20 stereo 16-bit frames = 80 bytes
    good = wav_bytes(1, 2, 48000, 16, pcm)
    off = good.index(b"fmt ") + 8 + 12                                  # blockAlign field
    lying = bytearray(good)
    lying[off:off + 2] = struct.pack("<H", 2)                           # claims 2 bytes per frame, needs 4
    p = tmp_path / "lying.wav"
    p.write_bytes(bytes(lying))
    r = subprocess.run([RENDER, "--dump-wav", str(p), "1", str(tmp_path / "x")], capture_output=True, text=True)
    assert r.returncode == 1 and "blockAlign" in r.stderr
    padded = bytearray(wav_bytes(1, 2, 48000, 16, pcm + b"\0\0"))      # 82 bytes of data, blockAlign 6: 13 frames ((13-1)*6 + 4 = 76 <= 82)
    padded[off:off + 2] = struct.pack("<H", 6)
    q = tmp_path / "padded.wav"
    q.write_bytes(bytes(padded))
    m, d = dump(q, 1.0, tmp_path)
    assert m["frames"] == 13 and d.shape == (2, 13)
    assert d[0, 1] == np.float32(300 / 32768.0)                          # frame 1 starts at byte 6 = sample 3


REF_IR = "/root/reference/ir"


def ref_wav_decode(raw: bytes):
    """wav.cu:46-118 restated on bytes: RIFF header, the NEXT chunk is taken as `fmt ` (id not checked, wav.cu:71-74), the
    one after it as `data` (wav.cu:85-89), frames = dataBytes / (channels * bytes) (wav.cu:95), stereo 16- or 24-bit only
    (wav.cu:103-114); samples scaled by the C restatement of f_wavConvert / f_wavConvert24 (half scale)."""
    assert raw[:4] == b"RIFF" and raw[8:12] == b"WAVE"
    fmt_size = struct.unpack_from("<I", raw, 16)[0]
    tag, ch, rate, _, align, bits = struct.unpack_from("<HHIIHH", raw, 20)
    p = 20 + fmt_size
    data_id, data_len = raw[p:p + 4], struct.unpack_from("<I", raw, p + 4)[0]
    body = raw[p + 8:p + 8 + data_len]
    frames = data_len // (ch * bits // 8)
    body = np.frombuffer(body[:frames * ch * bits // 8], np.uint8)
    x = O.pcm16_to_float(body.view(np.int16)) if bits == 16 else O.pcm24_to_float(body)
    return dict(tag=tag, channels=ch, rate=rate, align=align, bits=bits, frames=frames, fmt_id=raw[12:16], data_id=data_id), x.reshape(frames, ch).T


@pytest.mark.skipif(not os.path.isdir(REF_IR), reason="the reference's IR library exists only in the build container")
def test_every_ir_of_the_reference_library_decodes_bit_exact(tmp_path):
    """On-disk format step before the path: all 153 wavs shipped under ir/ (16- and 24-bit stereo, some
    with LIST / cue chunks after the data) go through the product decoder and give the reference's samples bit for bit;
    every line of every index file (main.cu:72-80) names one of them."""
    wavs = sorted(os.path.join(d, f) for d, _, fs in os.walk(REF_IR) for f in fs if f.endswith(".wav"))
    assert len(wavs) >= 150
    seen_bits = set()
    for w in wavs:
        raw = open(w, "rb").read()
        want_meta, want = ref_wav_decode(raw)
        assert want_meta["fmt_id"] == b"fmt " and want_meta["data_id"] == b"data", w   # the reference's layout assumption holds for its own files
        meta, got = dump(w, 0.5, tmp_path)
        assert (meta["channels"], meta["rate"], meta["bits"], meta["frames"]) == (2, want_meta["rate"], want_meta["bits"], want_meta["frames"]), w
        assert np.array_equal(got, want), w
        seen_bits.add(meta["bits"])
    assert seen_bits == {16, 24}
    root = os.path.dirname(REF_IR)
    n_lines = 0
    for idx in sorted(f for f in os.listdir(REF_IR) if f.endswith(".index")):
        for line in open(os.path.join(REF_IR, idx)).read().splitlines():
            if line.strip():
                n_lines += 1
                assert os.path.normpath(os.path.join(root, line.strip())) in wavs, (idx, line)
    assert n_lines >= len(wavs)
