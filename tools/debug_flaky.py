import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "cuda-audio_b200", "python"))
import numpy as np
from oracle import oracle as O
import cuda_audio_b200 as m
fs = 48000
def case(N, B, tag):
    L = N - B
    irs = [[O.synth_ir(L, fs, 1000 + 2 * i + o) for o in range(2)] for i in range(2)]
    warm = 100
    x = np.stack([np.concatenate([np.zeros(warm * B, np.float32), O.synth_audio(B * 150, 2000 + i)]) for i in range(2)])
    truth = O.engine_truth(x, irs, [dict(wet=1.0)] * 2)
    with m.Engine(period=B, max_ir_frames=L) as e:
        for i in range(2):
            e.load_ir(i, irs[i][0], irs[i][1]); e.set_params(0, i, select=i, wet=1.0, dry=0.0)
        y = e.render(x[None])[0]
        st = e.stats()
    for o in range(2):
        d = np.abs(y[o] - truth[o]); bad = np.nonzero(d > 1e-4)[0]
        print(tag, N, B, "split", st.mac_split, "out", o, "rel %.2e" % O.rel_l2(y[o], truth[o]), "bad samples", len(bad), ("first period %d, periods hit %s" % (bad[0] // B, sorted(set(bad // B))[:12])) if len(bad) else "", flush=True)
for rep in range(3):
    case(4096, 64, "rep%d" % rep)
    case(65536, 256, "rep%d" % rep)
