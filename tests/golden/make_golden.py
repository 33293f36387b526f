"""Generate golden vectors from the REAL reference (oracle/_ref = unmodified limitz/cuda-audio
conv.cu + wav.cu, cuFFT path) on a GPU box:

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'
    cp gpurun_out/golden/*.npz tests/golden/

Inputs are regenerated from seeds by tests (oracle.synth_ir / synth_audio); only the
reference's outputs (and the case description) are stored.  Cases:
  A  protocol run (DC/Nyquist-free IRs, silent warm-up, wet only)             N=4096  B=64
  B  reference defaults + pan/level/predelay from the first period (fade-in)  N=16384 B=256
  C  quirk-exposing: IRs WITHOUT DC/Nyquist correction, loud input (clamp)    N=8192  B=128
  D  IR switch mid-run through the reference's MIDI handler (select + glide)  N=8192  B=128
  W  PCM16 / PCM24 wav decode through the reference's WavFile
"""
import json
import os
import struct
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle as O  # noqa: E402
from oracle import refgpu  # noqa: E402

FS = 48000


def case_inputs(name):
    """(N, B, irs[slot][ch], x[2][n], cc[2], events) -- shared with tests/test_oracle_golden.py"""
    if name == "A":
        N, B = 4096, 64
        L = N - B
        irs = [[O.synth_ir(L, FS, 1000 + 2 * i + o) for o in range(2)] for i in range(2)]
        x = np.stack([np.concatenate([np.zeros(100 * B, np.float32), O.synth_audio(B * 150, 2000 + i)]) for i in range(2)])
        cc = [dict(select=0, wet=1.0, dry=0.0), dict(select=1, wet=1.0, dry=0.0)]
        return N, B, irs, x, cc, []
    if name == "B":
        N, B, pd = 16384, 256, 300
        L = N - 2 * B - pd
        irs = [[O.synth_ir(L, FS, 1100 + 2 * i + o) for o in range(2)] for i in range(2)]
        x = np.stack([O.synth_audio(B * 200, 2100 + i) for i in range(2)])
        cc = [dict(select=0, predelay=pd, panWet=0.3, panDry=-0.2, level=0.8), dict(select=1, panWet=-0.5, panDry=0.4)]
        return N, B, irs, x, cc, []
    if name == "C":
        N, B = 8192, 128
        L = N - B
        irs = [[O.synth_ir(L, FS, 1200 + 2 * i + o, parity_safe=False) for o in range(2)] for i in range(2)]
        x = np.stack([O.synth_audio(B * 120, 2200 + i, rms=0.4) for i in range(2)])
        cc = [dict(select=0, wet=0.9, dry=0.3, panWet=0.2), dict(select=1, wet=0.9, dry=0.3, panWet=-0.3)]
        return N, B, irs, x, cc, []
    if name == "D":
        N, B = 8192, 128
        L = 3000
        irs = [[O.synth_ir(L, FS, 1300 + 2 * i + o) for o in range(2)] for i in range(3)]
        x = np.stack([O.synth_audio(B * 260, 2300 + i) for i in range(2)])
        cc = [dict(select=0, wet=1.0, dry=0.0, speed=40), dict(select=1, wet=1.0, dry=0.0, speed=40)]
        # (period, input, cc-field 1..8 (select, predelay, dry, wet, speed, panDry, panWet, level), 7-bit value)
        events = [(100, 0, 1, 127 * 2 // 3 + 1), (160, 1, 4, 64), (200, 1, 1, 0)]
        return N, B, irs, x, cc, events
    if name == "E":   # not a golden: unconstrained IRs below the clamp, for the live drop-in comparison with engine.ref_quirks
        N, B = 8192, 128
        L = N - B - 500
        irs = [[O.synth_ir(L, FS, 1400 + 2 * i + o, parity_safe=False) for o in range(2)] for i in range(2)]
        x = np.stack([O.synth_audio(B * 150, 2400 + i) for i in range(2)])
        cc = [dict(select=0, wet=0.9, dry=0.3, panWet=0.2, predelay=333), dict(select=1, wet=0.9, dry=0.3, panWet=-0.3)]
        return N, B, irs, x, cc, []
    raise KeyError(name)


def write_wav(path, data_bytes, bits, extra_chunks=b""):
    block = 2 * bits // 8
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(data_bytes) + len(extra_chunks)) + b"WAVE")
        f.write(b"fmt " + struct.pack("<IHHIIHH", 16, 1, 2, 44100, 44100 * block, block, bits))
        f.write(b"data" + struct.pack("<I", len(data_bytes)) + data_bytes)
        f.write(extra_chunks)


def wav_cases(tmpdir):
    rng = np.random.default_rng(7)
    pcm16 = rng.integers(-32768, 32768, size=2 * 777, dtype=np.int64).astype(np.int16)
    pcm16[:4] = [32767, -32768, 0, -1]
    raw24 = rng.integers(0, 256, size=6 * 555, dtype=np.int64).astype(np.uint8)
    raw24[:6] = [0xFF, 0xFF, 0x7F, 0x00, 0x00, 0x80]  # +max, -max
    p16 = os.path.join(tmpdir, "g16.wav")
    p24 = os.path.join(tmpdir, "g24.wav")
    write_wav(p16, pcm16.tobytes(), 16)
    write_wav(p24, raw24.tobytes(), 24, extra_chunks=b"LIST" + struct.pack("<I", 4) + b"INFO")
    return pcm16, raw24, p16, p24


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    for name in "ABCD":
        N, B, irs, x, cc, events = case_inputs(name)
        ref = refgpu.RefGpu(N)
        for s, pair in enumerate(irs):
            ref.prepare(s, pair[0], pair[1], B)
        for i in range(2):
            ref.set_cc(i, **cc[i])
        periods = x.shape[1] // B
        L = np.empty(periods * B, np.float32)
        R = np.empty(periods * B, np.float32)
        ev = {}
        for e in events:
            ev.setdefault(e[0], []).append(e)
        cc_trace = []
        for t in range(periods):
            for (_, inp, field, val) in ev.get(t, []):
                ref.midi_cc(inp, field, val)
            l, r = ref.process(x[0, t * B:(t + 1) * B], x[1, t * B:(t + 1) * B])
            L[t * B:(t + 1) * B] = l
            R[t * B:(t + 1) * B] = r
            if events:
                cc_trace.append([ref.get_cc(0)["select"], ref.get_cc(0)["vsteps"], ref.get_cc(1)["select"], ref.get_cc(1)["vsteps"]])
        np.savez_compressed(os.path.join(outdir, f"ref_{name}.npz"), L=L, R=R, cc_trace=np.array(cc_trace, np.int64),
                            meta=json.dumps(dict(N=N, B=B, cc=cc, events=events, periods=periods)))
        print("case", name, "rms", float(np.sqrt((L ** 2).mean())), float(np.sqrt((R ** 2).mean())), flush=True)
    pcm16, raw24, p16, p24 = wav_cases(outdir)
    l16, r16 = refgpu.wav_decode(p16)
    l24, r24 = refgpu.wav_decode(p24)
    np.savez_compressed(os.path.join(outdir, "ref_W.npz"), pcm16=pcm16, raw24=raw24, l16=l16, r16=r16, l24=l24, r24=r24)
    os.remove(p16)
    os.remove(p24)
    print("wav", len(l16), len(l24))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
