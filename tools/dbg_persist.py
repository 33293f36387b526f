"""debug: tiny persistent-MAC run with progress lines (each configuration in its own process)"""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "cuda-audio_b200", "python"))
if len(sys.argv) > 1:
    import numpy as np
    import cuda_audio_b200 as ca
    B, K, P = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    L = B * P
    tiers = None
    if len(sys.argv) > 4:
        tiers = [(64, 8), (512, 3), (2048, 0)]
        L = 64 * 8 + 512 * 3 + 2048 * 2 - 100
    nper = int(sys.argv[5]) if len(sys.argv) > 5 else 6
    rng = np.random.default_rng(0)
    with ca.Engine(period=B, max_ir_frames=L, n_instances=K, n_ir_slots=K, mac_split=1, tiers=tiers, max_voices=2) as e:
        for s in range(K):
            h = (rng.standard_normal((2, L)) * 0.1).astype(np.float32)
            e.load_ir(s, h[0], h[1])
            for i in range(2):
                e.set_params(s, i, select=s, wet=1.0, dry=0.0); e.set_glide(s, i, 1.0)
        print("engine up", flush=True)
        for t in range(nper):
            y = e.process((rng.standard_normal((K, 2, B)) * 0.1).astype(np.float32))
            if t < 12 or t % 16 == 0: print("period", t, float(np.abs(y).max()), flush=True)
    print("done", flush=True)
    sys.exit(0)
for env, args in [({"CA_MAC_PERSIST": "1", "CA_MAC_SLOTS": "1000", "CA_FUSE": "0"}, "64 3 0 tiers 100"), ({"CA_MAC_PERSIST": "1", "CA_MAC_SLOTS": "1", "CA_FUSE": "0"}, "64 3 0 tiers 100"),
                  ({"CA_MAC_PERSIST": "1", "CA_MAC_SLOTS": "1", "CA_FUSE": "0"}, "64 11 0 tiers 200"), ({"CA_MAC_PERSIST": "1", "CA_MAC_SLOTS": "3", "CA_FUSE": "0"}, "64 11 0 tiers 200")]:
    print("====", env, args, flush=True)
    try:
        r = subprocess.run([sys.executable, "-u", __file__] + args.split(), env=dict(os.environ, **env), timeout=40, capture_output=True, text=True)
        print(r.stdout[-1500:], r.stderr[-800:], "rc", r.returncode, flush=True)
    except subprocess.TimeoutExpired as ex:
        print("TIMEOUT", (ex.stdout or b"")[-1500:], (ex.stderr or b"")[-500:], flush=True)
