"""Single-instance per-period latency through ca_process (pinned buffers, CUDA graph) for several
partitionings.  usage: python tools/latency.py B IR_FRAMES [periods]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "cuda-audio_b200", "python"))
import numpy as np
import cuda_audio_b200 as ca
B, L = int(sys.argv[1]), int(sys.argv[2])
periods = int(sys.argv[3]) if len(sys.argv) > 3 else 3000
rng = np.random.default_rng(0)
env = np.exp(-6.91 * np.arange(L) / (0.8 * L)).astype(np.float32)
h = rng.standard_normal((4, L)).astype(np.float32) * env
h /= np.sqrt((h ** 2).sum(axis=1, keepdims=True))
P = (L + B - 1) // B
for name, kw in [("uniform", {}), ("uniform+L2persist", dict(extra=ca.FLAG_L2_PERSIST)), ("tiers g8 max16384", dict(tiers="auto")), ("tiers g8 max16384 async", dict(tiers="auto", extra=ca.FLAG_ASYNC_TIERS)), ("tiers g8 max4096", dict(tiers="auto", tier_max_block=4096)),
                 ("tiers g4 max4096", dict(tiers="auto", tier_growth=4, tier_max_block=4096)), ("tiers g8 max2048", dict(tiers="auto", tier_max_block=2048)),
                 ("tiers g16 max16384", dict(tiers="auto", tier_growth=16))]:
    extra = kw.pop("extra", 0)
    try:
        e = ca.Engine(period=B, max_ir_frames=L, flags=ca.FLAG_GRAPH | extra, sample_rate=48000, **kw)
    except ca.CaError as ex:
        print(name, "->", ex); continue
    e.load_ir(0, h[0], h[1]); e.load_ir(1, h[2], h[3])
    for i in range(2):
        e.set_params(0, i, select=i); e.set_glide(0, i, 0.5)
    a, b = ca.PinnedArray((1, 2, B)), ca.PinnedArray((1, 2, B))
    a.array[...] = 0.05
    st = e.stats()
    warm = min(P, 12000) + 300
    for _ in range(warm): e.process_raw(a.ptr, b.ptr)
    e.reset_stats()
    for _ in range(periods): e.process_raw(a.ptr, b.ptr)
    s = e.stats()
    e.reset_stats()
    import time
    pace = float(os.environ.get("PACE_US", 0.1e6 * B / 48000))  # 10 x faster than real time
    tick = time.perf_counter()
    for _ in range(periods):
        tick += pace * 1e-6
        while time.perf_counter() < tick: pass
        e.process_raw(a.ptr, b.ptr)
    sp = e.stats()
    print(f"   paced every {pace:.0f} us: p50={sp.p50_us:.1f} p99={sp.p99_us:.1f} max={sp.max_us:.1f}", end="  |  ")
    print(f"B={B} L={L} {name:22s} tiers={list(st.tier_block[:st.n_tiers])}x{list(st.tier_parts[:st.n_tiers])} split={s.mac_split} "
          f"p50={s.p50_us:.1f} p99={s.p99_us:.1f} max={s.max_us:.1f} us (deadline {1e6*B/48000:.0f}) bytes/period={s.mac_bytes_amortized/1e6:.2f} MB", flush=True)
    e.close(); a.free(); b.free()
