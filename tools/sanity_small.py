"""Small end-to-end exercise of every kernel (uniform, tiers, fused tier 0, cross-fade, predelay,
chunked host path) (written for compute-sanitizer runs; the tool is closed on this GPU pool, so it is a plain smoke run)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "cuda-audio_b200", "python"))
import numpy as np
import cuda_audio_b200 as ca
from oracle import oracle as O

def run(K, B, L, tiers, fuse, periods=40, **kw):
    os.environ["CA_FUSE"] = fuse
    irs = [[O.synth_ir(L, 48000, 10 + 2 * i + o) for o in range(2)] for i in range(3)]
    x = np.stack([np.stack([O.synth_audio(B * periods, 20 + 2 * s + i) for i in range(2)]) for s in range(K)])
    with ca.Engine(period=B, max_ir_frames=L, n_instances=K, n_ir_slots=3, tiers=tiers, **kw) as e:
        for j in range(3):
            e.load_ir(j, irs[j][0], irs[j][1])
        for s in range(K):
            for i in range(2):
                e.set_params(s, i, select=i, wet=1.0, dry=0.2, predelay=17 * (s % 3))
                e.set_glide(s, i, 1.0)
        out = np.empty((K, 2, B * periods), np.float32)
        for t in range(periods):
            if t == periods // 2 and K > 1:   # instance 0 cross-fades; instance 1 is the one checked
                e.set_params(0, 0, select=2, wet=1.0, dry=0.2, vsteps=5)
            out[:, :, t * B:(t + 1) * B] = e.process(x[:, :, t * B:(t + 1) * B])
    truth = O.engine_truth(x[1 % K], [irs[0], irs[1]], [dict(wet=1.0, dry=0.2)] * 2, predelay=17 * ((1 % K) % 3))
    err = O.rel_l2(out[1 % K, 0], truth[0])
    print(f"K={K} B={B} tiers={tiers} fuse={fuse}: rel-L2 {err:.2e}", flush=True)
    assert err < 5e-6

run(2, 64, 64 * 9, None, "0")
run(2, 64, 64 * 8 + 512 * 3, [(64, 8), (512, 0)], "0")
run(2, 64, 64 * 8 + 512 * 3, [(64, 8), (512, 0)], "1")
run(3, 256, 256 * 8 + 2048 * 2, [(256, 8), (2048, 0)], "0", periods=24)
run(520, 64, 64 * 6, None, "0", periods=6)          # chunked host pipeline (>= 512 instances)
run(1, 32, 32 * 8 + 256 * 9 + 2048 * 2, [(32, 8), (256, 8), (2048, 0)], "1", periods=140, flags=ca.FLAG_GRAPH)
print("sanity ok")
