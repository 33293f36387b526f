"""debug: the persistent-MAC parity test body with progress lines"""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "cuda-audio_b200", "python"))
if len(sys.argv) > 1:
    import numpy as np
    import cuda_audio_b200 as m
    xfade, pdel, usetiers = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    B, K = 64, 11
    tiers = [(64, 8), (512, 3), (2048, 0)] if usetiers else None
    L = 64 * 8 + 512 * 3 + 2048 * 2 - 100
    rng = np.random.default_rng(0)
    n = B * 200
    x = (rng.standard_normal((K, 2, n)) * 0.1).astype(np.float32)
    with m.Engine(period=B, max_ir_frames=L, n_instances=K, n_ir_slots=2 * K, tiers=tiers, max_voices=2, mac_split=1) as e:
        for s in range(K):
            for i in range(2):
                h = (rng.standard_normal((2, L)) * 0.02).astype(np.float32)
                e.load_ir(2 * s + i, h[0], h[1])
                e.set_params(s, i, select=2 * s + i, wet=0.8, dry=0.1, predelay=7 * s * pdel, panWet=0.05 * s - 0.2)
                e.set_glide(s, i, 0.8)
        print("engine up", flush=True)
        for t in range(n // B):
            if t == 90 and xfade:
                e.set_params(4, 0, select=3, wet=0.8, dry=0.1, predelay=28 * pdel, panWet=0.0, vsteps=30)
            y = e.process(x[:, :, t * B:(t + 1) * B])
            if t % 20 == 0 or 88 <= t <= 140: print("period", t, float(np.abs(y).max()), flush=True)
    print("done", flush=True)
    sys.exit(0)
base = {"CA_MAC_PERSIST": "1", "CA_MAC_SLOTS": "1", "CA_FUSE": "0"}
for env, args in [(base, "1 1 1"), (base, "0 1 1"), (base, "1 0 1"), (base, "1 0 0"), (dict(base, CA_MAC_PERSIST="0"), "1 1 1")]:
    print("====", env, args, flush=True)
    try:
        r = subprocess.run([sys.executable, "-u", __file__] + args.split(), env=dict(os.environ, **env), timeout=40, capture_output=True, text=True)
        print(r.stdout[-700:], r.stderr[-800:], "rc", r.returncode, flush=True)
    except subprocess.TimeoutExpired as ex:
        print("TIMEOUT", (ex.stdout or b"")[-700:], (ex.stderr or b"")[-500:], flush=True)
