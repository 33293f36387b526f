"""L2 / HBM read bandwidth of this GPU and the single-instance (L2-resident) FDL MAC against it."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "cuda-audio_b200", "python"))
import numpy as np
import cuda_audio_b200 as ca
for mb in (8, 16, 32, 64, 96, 256, 4096):
    print(f"read sweep {mb:5d} MB: {ca.measure_read_gbs(mb << 20, 50 if mb <= 256 else 10):8.0f} GB/s", flush=True)
rng = np.random.default_rng(0)
for B, L, name in ((256, 192000, "cfg2 (9.2 MB)"), (64, 480000, "cfg3 uniform (23 MB)"), (256, 2880000, "cfg5 60 s (138 MB)")):
    for split in (0, 64, 148, 256):
        e = ca.Engine(period=B, max_ir_frames=L, flags=ca.FLAG_PROFILE, mac_split=split)
        h = rng.standard_normal((4, L)).astype(np.float32) * 1e-2
        e.load_ir(0, h[0], h[1]); e.load_ir(1, h[2], h[3])
        for i in range(2):
            e.set_params(0, i, select=i); e.set_glide(0, i, 0.5)
        x = np.full((1, 2, B), 0.05, np.float32)
        P = (L + B - 1) // B
        for _ in range(P + 20): e.process(x)
        e.reset_stats()
        for _ in range(300): e.process(x)
        s = e.stats()
        print(f"{name:22s} split={s.mac_split:3d} fwd={s.fwd_us:5.1f} mac={s.mac_us:6.1f} inv={s.inv_us:5.1f} us  MAC {s.mac_bytes / s.mac_us / 1e3:7.0f} GB/s", flush=True)
        e.close()
