"""Worker for tests/test_dropin_gpu.py::test_ir_split_group_through_the_class_api (a subprocess with CA_ENGINE_IR_SPLIT=G in
the environment: the options are read when the first Convolution object of a process is constructed).  ONE mirror
`Convolution` object whose engine is a ca_group over G GPUs (engine.ir_split, BASELINE configs[4]) driven by the same
harness code as the reference arm; output against the fp64 oracle, with an IR prepared while the object is live."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cuda-audio_b200", "python"))

import numpy as np  # noqa: E402

from oracle import oracle as O  # noqa: E402
from oracle import refgpu  # noqa: E402

DROPIN = os.path.join(ROOT, "tests", "dropin", "libdropin_conv.so")


def main():
    fs, B, L, nper = 48000, 256, 256 * 90 + 5, 160
    N = 32768
    c = refgpu.RefGpu(N, 0, DROPIN)
    irs = [[O.synth_ir(L, fs, 7000 + 2 * i + o) for o in range(2)] for i in range(2)]
    for i in range(2):
        c.prepare(i, irs[i][0], irs[i][1], B)
        c.set_cc(i, select=i, wet=0.8, dry=0.25, panWet=0.3 - 0.4 * i)
    x = np.stack([np.concatenate([np.zeros(100 * B, np.float32), O.synth_audio(B * nper, 7500 + i)]) for i in range(2)])
    y = c.render(x[0], x[1], B)                                   # one onProcess per period
    truth = O.engine_truth(x, irs, [dict(wet=0.8, dry=0.25, panWet=0.3), dict(wet=0.8, dry=0.25, panWet=-0.1)])
    sl = slice(100 * B, None)                                     # after the wet fade-in (conv.cu:27)
    err = max(O.rel_l2(y[o][sl], truth[o][sl]) for o in range(2))
    # a longer IR for slot 1 on the live object: the group is rebuilt behind the same surface
    longer = [O.synth_ir(L + 3000, fs, 7100 + o) for o in range(2)]
    c.prepare(1, longer[0], longer[1], B)
    y2 = c.render(x[0], x[1], B)
    truth2 = O.engine_truth(x, [irs[0], longer], [dict(wet=0.8, dry=0.25, panWet=0.3), dict(wet=0.8, dry=0.25, panWet=-0.1)])
    err2 = max(O.rel_l2(y2[o][sl], truth2[o][sl]) for o in range(2))
    print("IRSPLIT_RESULT " + json.dumps({"G": int(os.environ["CA_ENGINE_IR_SPLIT"]), "rel_l2": err, "rel_l2_after_prepare": err2}), flush=True)


if __name__ == "__main__":
    main()
