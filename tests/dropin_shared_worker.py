"""Worker for tests/test_dropin_gpu.py::test_shared_batched_engine_through_the_class_api (run as a subprocess
with CA_ENGINE_SHARED=K CA_ENGINE_TIERS=auto in the environment: the options are read when the first
Convolution object of a process is constructed).  K mirror `Convolution` objects, one host thread each like
one JACK client each (main.cu:31-39), share ONE batched engine; every object's output is compared with the
fp64 oracle.  Prints one JSON line."""
import json
import os
import sys
import threading

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cuda-audio_b200", "python"))

import numpy as np  # noqa: E402

from oracle import oracle as O  # noqa: E402
from oracle import refgpu  # noqa: E402

DROPIN = os.path.join(ROOT, "tests", "dropin", "libdropin_conv.so")


def main():
    K = int(os.environ["CA_ENGINE_SHARED"])
    fs, B, L, nper = 48000, 256, 256 * 70 + 5, 150
    N = 32768
    objs, irs, xs = [], [], []
    for k in range(K):
        c = refgpu.RefGpu(N, 0, DROPIN)
        pair = [[O.synth_ir(L, fs, 6000 + 8 * k + 2 * i + o) for o in range(2)] for i in range(2)]
        for i in range(2):
            c.prepare(i, pair[i][0], pair[i][1], B)
            c.set_cc(i, select=i, wet=0.9, dry=0.2 + 0.1 * k, panWet=0.1 * k)
        objs.append(c)
        irs.append(pair)
        xs.append(np.stack([np.concatenate([np.zeros(100 * B, np.float32), O.synth_audio(B * nper, 6500 + 2 * k + i)]) for i in range(2)]))
    outs = [None] * K

    def run(k):
        outs[k] = objs[k].render(xs[k][0], xs[k][1], B)      # ref_render: one onProcess per period, rendezvous inside

    th = [threading.Thread(target=run, args=(k,)) for k in range(K)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    errs = []
    sl = slice(100 * B, None)                                 # after the wet fade-in (conv.cu:27)
    late = B if int(os.environ.get("CA_ENGINE_SHARED_LATENCY", "0")) else 0   # engine.shared_latency 1: every output one period late
    for k in range(K):
        truth = O.engine_truth(xs[k], irs[k], [dict(wet=0.9, dry=0.2 + 0.1 * k, panWet=0.1 * k)] * 2)
        if late:
            truth = [np.concatenate([np.zeros(late, t.dtype), t[:-late]]) for t in truth]
        errs.append(max(O.rel_l2(outs[k][o][sl], truth[o][sl]) for o in range(2)))
    print("SHARED_RESULT " + json.dumps({"K": K, "rel_l2": errs}), flush=True)


if __name__ == "__main__":
    main()
