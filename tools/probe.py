"""Quick device-resident throughput probe of the batched engine (dev tool, not the bench).
usage: python tools/probe.py [K instances] [periods] [mac_split]   (env CA_MAC_VARIANT=0..3)
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cuda-audio_b200", "python"))
import torch  # noqa: E402  (device memory + RNG plumbing only)
import cuda_audio_b200 as ca  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 512
periods = int(sys.argv[2]) if len(sys.argv) > 2 else 50
split = int(sys.argv[3]) if len(sys.argv) > 3 else 0
tiers = "auto" if os.environ.get("CA_TIERS") else None
nvoices = int(os.environ.get("CA_VOICES", "0"))
fs, B, L = 48000, 256, 192000
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
n = torch.arange(L, device=dev, dtype=torch.float32)
env = torch.exp(-6.91 * n / (0.8 * L))
flags = (0 if os.environ.get('CA_NOPROFILE') else ca.FLAG_PROFILE) | (ca.FLAG_STREAMING if K >= 16 else 0) | (ca.FLAG_LEGACY_FFT if os.environ.get('CA_LEGACY_FFT') else 0)
t0 = time.time()
e = ca.Engine(period=B, max_ir_frames=L, n_instances=K, n_ir_slots=2 * K, flags=flags, mac_split=split, tiers=tiers, max_voices=nvoices,
              tier_growth=int(os.environ.get('CA_TIER_GROWTH', '0')), tier_max_block=int(os.environ.get('CA_TIER_MAXBLOCK', '0')))
g = torch.Generator(device=dev)
for s in range(2 * K):
    g.manual_seed(1000 + s)
    h = torch.randn(2, L, device=dev, generator=g) * env
    h = h / h.pow(2).sum(dim=1, keepdim=True).sqrt()
    torch.cuda.synchronize()
    e.load_ir_device(s, h[0].data_ptr(), h[1].data_ptr(), L)
for s in range(K):
    for i in range(2):
        e.set_params(s, i, select=2 * s + i, wet=1.0, dry=0.0)
        e.set_glide(s, i, 1.0)
print(f"setup {time.time() - t0:.1f}s", flush=True)
x = torch.randn(K, 2, B, device=dev) * (0.0 if os.environ.get("CA_ZERO_INPUT") else 0.1)
y = torch.empty(K, 2, B, device=dev)
torch.cuda.synchronize()
for _ in range(760):  # steady state: every FDL slot filled (the engine skips slots older than a voice's start)
    e.process_device(x.data_ptr(), y.data_ptr())
e.sync()
e.reset_stats()
t0 = time.time()
for _ in range(periods):
    e.process_device(x.data_ptr(), y.data_ptr())
e.sync()
dt = (time.time() - t0) / periods
st = e.stats()
if os.environ.get('CA_NOPROFILE'):
    pin, pout = ca.PinnedArray((K, 2, B)), ca.PinnedArray((K, 2, B))
    pin.array[...] = x.cpu().numpy()
    for _ in range(20):
        e.process_raw(pin.ptr, pout.ptr)
    t0 = time.time()
    for _ in range(periods):
        e.process_raw(pin.ptr, pout.ptr)
    e.sync()
    dte = (time.time() - t0) / periods
    print(f"K={K} device wall/period={dt * 1e6:.1f}us rt_channels={K * (B / fs) / dt:.0f} | e2e {dte * 1e6:.1f}us rt_channels={K * (B / fs) / dte:.0f} "
          f"persist={os.environ.get('CA_MAC_PERSIST', 'auto')} ctas={os.environ.get('CA_MAC_CTAS', '-')} pipeline={os.environ.get('CA_PIPELINE', 'auto')} "
          f"y_rms={float(y.pow(2).mean().sqrt()):.4f}/{float((pout.array.astype('f8') ** 2).mean() ** 0.5):.4f}", flush=True)
    sys.exit(0)
gbs = st.mac_bytes / (st.mac_us * 1e-6) / 1e9
print(f"tiers={list(st.tier_block[:st.n_tiers])}x{list(st.tier_parts[:st.n_tiers])} dev_bytes={st.device_bytes/1e9:.1f}GB total={st.total_us:.1f} fwd0={st.fwd_us:.1f} mac0={st.mac_us:.1f} inv0={st.inv_us:.1f} tfwd={st.tier_fwd_us:.1f} tmac={st.tier_mac_us:.1f} tinv={st.tier_inv_us:.1f} bytes/period={st.mac_bytes_amortized/1e9:.2f}GB")
print(f"variant={os.environ.get('CA_MAC_VARIANT', '1')} K={K} split={st.mac_split} fwd={st.fwd_us:.1f}us mac={st.mac_us:.1f}us inv={st.inv_us:.1f}us "
      f"wall/period={dt * 1e6:.1f}us MAC {gbs:.0f} GB/s ({gbs / 6460.2:.3f} of measured HBM) "
      f"rt_channels={K * (B / fs) / (st.total_us * 1e-6):.0f} y_rms={float(y.pow(2).mean().sqrt()):.4f}", flush=True)
