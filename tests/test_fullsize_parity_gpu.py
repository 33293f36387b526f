"""Parity of the configurations that produce the headline, at their real sizes.

* BASELINE configs[1] at the reference's own size (fftSize 262144, B = 256, 4 s IR) against the LIVE
  reference (oracle/_ref = unmodified conv.cu + cuFFT, conv.cu:287-466), through the uniform engine AND
  through the non-uniform tiers with the flags the bench uses (streaming hints, persistent MAC).
* A batch of 2048 tiered instances through the chunked ca_process host pipeline, long enough for the
  16 K tier's delay line to wrap, three sampled instances against the fp64 oracle.
* BASELINE configs[4]: one 60 s IR (P = 11250) on one GPU, >= 2000 fp64 dot-product samples plus the
  fp64 FFT convolution of the whole signal.
"""
import numpy as np
import pytest

from oracle import oracle as O
from oracle import refgpu

pytestmark = pytest.mark.gpu
FS = 48000
TOL_REF = 1e-5    # north star: relative L2 vs the reference's conv.cu on the same input and IR
TOL_FP64 = 1e-4   # north star: vs the fp64 direct-convolution oracle


def ca():
    import cuda_audio_b200 as m
    return m


needs_ref = pytest.mark.skipif(not refgpu.available(), reason="oracle/_ref/libref_conv.so not built")


@needs_ref
def test_cfg2_reference_size_uniform_and_tiered(monkeypatch):
    """SURVEY 8(c) protocol: DC/Nyquist-free IRs, 100 silent warm-up periods, unclipped levels,
    predelay 0; 1000 periods of noise so every partition of every tier (16 K tier: 11 blocks of 64
    periods) carries signal."""
    m = ca()
    N, B, L = 262144, 256, 192000
    irs = [[O.synth_ir(L, FS, 1000 + 2 * i + o) for o in range(2)] for i in range(2)]
    warm, nper = 100, 1000
    x = np.stack([np.concatenate([np.zeros(warm * B, np.float32), O.synth_audio(B * nper, 2000 + i)]) for i in range(2)])
    ref = refgpu.RefGpu(N)
    for i in range(2):
        ref.prepare(i, irs[i][0], irs[i][1], B)
        ref.set_cc(i, select=i, wet=1.0, dry=0.0)
    rl, rr = ref.render(x[0], x[1], B)
    sl = slice(warm * B, None)
    truth = O.engine_truth(x, irs, [dict(wet=1.0)] * 2)
    ref_err = [O.rel_l2(r[sl], truth[o][sl]) for o, r in enumerate((rl, rr))]

    def ours(**kw):
        with m.Engine(period=B, max_ir_frames=L, **kw) as e:
            for i in range(2):
                e.load_ir(i, irs[i][0], irs[i][1])
                e.set_params(0, i, select=i, wet=1.0, dry=0.0)
            y = e.render(x[None])[0]
            return y, e.stats()

    yu, su = ours()
    assert su.partitions == 750 and su.n_tiers == 1
    monkeypatch.setenv("CA_MAC_PERSIST", "1")      # the schedule the batched bench runs, forced for one instance
    yt, st = ours(tiers="auto", flags=m.FLAG_STREAMING)
    monkeypatch.delenv("CA_MAC_PERSIST")
    assert [int(st.tier_block[j]) for j in range(st.n_tiers)] == [256, 2048, 16384]
    yg, _ = ours(tiers="auto", flags=m.FLAG_GRAPH)  # the latency schedule (graph, split MAC, fused tier 0)
    for name, y in (("uniform", yu), ("tiered+streaming+persistent", yt), ("tiered+graph", yg)):
        for o, r in enumerate((rl, rr)):
            err = O.rel_l2(y[o][sl], r[sl])
            assert err < TOL_REF, (name, o, err, ref_err)
            assert O.rel_l2(y[o][sl], truth[o][sl]) < 5e-6, (name, o)


def test_batch_2048_tiered_chunked_host_pipeline_vs_fp64():
    """K = 2048 distinct-IR instances, cfg2 geometry, tiers auto + streaming hints, host buffers through
    ca_process (3-chunk H2D | kernels | D2H pipeline, persistent MAC, side-stream tiers, PDL): 900
    periods, instances 0 / 1023 / 2047 against the fp64 oracle."""
    import torch
    m = ca()
    B, L, K, nper = 256, 192000, 2048, 900
    dev = torch.device("cuda", 0)
    picks = [0, 1023, 2047]
    n = torch.arange(L, device=dev, dtype=torch.float32)
    env = torch.exp(-6.91 * n / (0.8 * L))
    g = torch.Generator(device=dev)
    kept = {}
    with m.Engine(period=B, max_ir_frames=L, n_instances=K, n_ir_slots=2 * K, flags=m.FLAG_STREAMING, tiers="auto") as e:
        for s in range(2 * K):
            g.manual_seed(1000 + s)
            h = torch.randn(2, L, device=dev, generator=g) * env
            h = h / h.pow(2).sum(dim=1, keepdim=True).sqrt()
            e.load_ir_device(s, h[0].data_ptr(), h[1].data_ptr(), L)
            if s // 2 in picks:
                kept[s] = h.cpu().numpy().astype(np.float64)
        for s in range(K):
            for i in range(2):
                e.set_params(s, i, select=2 * s + i, wet=0.8, dry=0.3, panWet=0.25 if i == 0 else -0.5, level=0.9)
                e.set_glide(s, i, 0.8)
        pin, pout = m.PinnedArray((K, 2, B)), m.PinnedArray((K, 2, B))
        g.manual_seed(77)
        xs = np.zeros((len(picks), 2, nper * B), np.float32)
        ys = np.zeros((len(picks), 2, nper * B), np.float32)
        for t in range(nper):
            xb = (torch.randn(K, 2, B, device=dev, generator=g) * 0.1).clamp_(-0.9, 0.9).cpu().numpy()
            pin.array[...] = xb
            e.process_raw(pin.ptr, pout.ptr)
            xs[:, :, t * B:(t + 1) * B] = xb[picks]
            ys[:, :, t * B:(t + 1) * B] = pout.array[picks]
        st = e.stats()
        assert [int(st.tier_block[j]) for j in range(st.n_tiers)] == [256, 2048, 16384] and st.tier0_fused == 0
        pin.free()
        pout.free()
    pr = [dict(wet=0.8, dry=0.3, panWet=0.25, level=0.9), dict(wet=0.8, dry=0.3, panWet=-0.5, level=0.9)]
    for j, s in enumerate(picks):
        irs = [[kept[2 * s + i][o] for o in range(2)] for i in range(2)]
        truth = O.engine_truth(xs[j], irs, pr)
        for o in range(2):
            err = O.rel_l2(ys[j, o], truth[o])
            assert err < 5e-6, (s, o, err)


def test_batch_growth4_four_tiers_chunked_host_pipeline_vs_fp64():
    """The four-tier shape (tier_growth = 4: 256 x 4 | 1024 x 3 | 4096 x 3 | 16384 x ..) as a batch of 768 through
    ca_process (chunked pipeline, persistent single-stage MAC on every tier): instances 0 / 383 / 767 vs fp64."""
    import torch
    m = ca()
    B, L, K, nper = 256, 40000, 768, 260
    dev = torch.device("cuda", 0)
    picks = [0, 383, 767]
    n = torch.arange(L, device=dev, dtype=torch.float32)
    env = torch.exp(-6.91 * n / (0.8 * L))
    g = torch.Generator(device=dev)
    kept = {}
    with m.Engine(period=B, max_ir_frames=L, n_instances=K, n_ir_slots=2 * K, flags=m.FLAG_STREAMING, tiers="auto", tier_growth=4) as e:
        for s in range(2 * K):
            g.manual_seed(5000 + s)
            h = torch.randn(2, L, device=dev, generator=g) * env
            h = h / h.pow(2).sum(dim=1, keepdim=True).sqrt()
            e.load_ir_device(s, h[0].data_ptr(), h[1].data_ptr(), L)
            if s // 2 in picks:
                kept[s] = h.cpu().numpy().astype(np.float64)
        for s in range(K):
            for i in range(2):
                e.set_params(s, i, select=2 * s + i, wet=0.8, dry=0.3, panWet=0.25 if i == 0 else -0.5, level=0.9, predelay=(7 * s) % 300)
                e.set_glide(s, i, 0.8)
        pin, pout = m.PinnedArray((K, 2, B)), m.PinnedArray((K, 2, B))
        g.manual_seed(78)
        xs = np.zeros((len(picks), 2, nper * B), np.float32)
        ys = np.zeros((len(picks), 2, nper * B), np.float32)
        for t in range(nper):
            xb = (torch.randn(K, 2, B, device=dev, generator=g) * 0.1).clamp_(-0.9, 0.9).cpu().numpy()
            pin.array[...] = xb
            e.process_raw(pin.ptr, pout.ptr)
            xs[:, :, t * B:(t + 1) * B] = xb[picks]
            ys[:, :, t * B:(t + 1) * B] = pout.array[picks]
        st = e.stats()
        assert [int(st.tier_block[j]) for j in range(st.n_tiers)] == [256, 1024, 4096, 16384] and st.tier0_fused == 0
        pin.free()
        pout.free()
    pr = [dict(wet=0.8, dry=0.3, panWet=0.25, level=0.9), dict(wet=0.8, dry=0.3, panWet=-0.5, level=0.9)]
    for j, s in enumerate(picks):
        irs = [[kept[2 * s + i][o] for o in range(2)] for i in range(2)]
        truth = O.engine_truth(xs[j], irs, pr, predelay=(7 * s) % 300)
        for o in range(2):
            err = O.rel_l2(ys[j, o], truth[o])
            assert err < 5e-6, (s, o, err)


def test_cfg5_60s_ir_sampled_fp64_and_fftconvolve():
    """60 s IR at 48 kHz (2 880 000 frames, P = 11250), one GPU, uniform partitioning with the split MAC:
    the full output against the fp64 FFT convolution, and 2000 random output samples against fp64 dot
    products (direct convolution is infeasible at this size)."""
    m = ca()
    B, L = 256, 60 * FS
    nper = L // B + 300
    irs = [[O.synth_ir(L, FS, 500 + 2 * i + o) for o in range(2)] for i in range(2)]
    x = np.stack([O.synth_audio(B * nper, 600 + i) for i in range(2)])
    pr = [dict(wet=1.0, dry=0.0)] * 2
    with m.Engine(period=B, max_ir_frames=L, flags=m.FLAG_GRAPH) as e:
        assert e.stats().partitions == 11250
        for i in range(2):
            e.load_ir(i, irs[i][0], irs[i][1])
            e.set_params(0, i, select=i, **pr[i])
            e.set_glide(0, i, 1.0)
        y = e.render(x[None])[0]
    truth = O.engine_truth(x, irs, pr)
    for o in range(2):
        assert O.rel_l2(y[o], truth[o]) < 5e-6, (o, O.rel_l2(y[o], truth[o]))
    idx = np.sort(np.random.default_rng(3).integers(L, B * nper, 2000))
    d = O.direct_conv_at(x[0], irs[0][0], idx) + O.direct_conv_at(x[1], irs[1][0], idx)
    assert O.rel_l2(y[0][idx], d) < TOL_FP64, O.rel_l2(y[0][idx], d)
