"""Parity tests proper: the CUDA engine, called through the C ABI, against the oracles.

Tolerances (BASELINE.json north_star): relative L2 <= 1e-5 vs the reference's own conv.cu
(oracle/_ref, cuFFT path, run live on the same GPU under the parity protocol of SURVEY 8c),
<= 1e-4 vs the FP64 convolution oracle.  The engine is fp32; observed errors are ~3e-7.
"""
import os

import numpy as np
import pytest

from oracle import oracle as O
from oracle import refgpu

pytestmark = pytest.mark.gpu

TOL_FP64 = 1e-4
TOL_REF = 1e-5


def ca():
    import cuda_audio_b200 as m
    return m


def make_irs(L, fs, n_in=2, n_out=2, seed0=1000):
    return [[O.synth_ir(L, fs, seed0 + i * n_out + o) for o in range(n_out)] for i in range(n_in)]


def load_true_stereo(eng, irs):
    """input i uses bank slot i = stereo IR (irs[i][0], irs[i][1])"""
    for i, pair in enumerate(irs):
        eng.load_ir(i, pair[0], pair[1] if len(pair) > 1 else None)


def test_cfg1_mono_vs_fp64():
    """BASELINE configs[0]: mono 44.1 kHz, 256-frame period, 1 s synthetic IR, offline render."""
    m = ca()
    fs, B, L = 44100, 256, 44100
    h = O.synth_ir(L, fs, 1000)
    x = O.synth_audio(B * 400, 2000)
    with m.Engine(period=B, max_ir_frames=L, n_in=1, n_out=1, n_ir_slots=1, sample_rate=fs) as e:
        e.load_ir(0, h)
        e.set_params(0, 0, select=0, wet=1.0, dry=0.0)
        e.set_glide(0, 0, 1.0)
        y = e.render(x[None, None, :])[0, 0]
        st = e.stats()
    truth = O.fft_conv(x, h)
    err = O.rel_l2(y, truth)
    assert st.partitions == 173
    assert err < TOL_FP64, err
    assert err < 5e-6, err  # fp32 partitioned convolution sits ~3e-7 from fp64
    # spot-check with the direct (time-domain) fp64 convolution on sampled outputs
    idx = np.random.default_rng(0).integers(0, len(x), 512)
    d = O.direct_conv_at(x, h, idx)
    assert O.rel_l2(y[idx], d) < TOL_FP64


@pytest.mark.parametrize("B", [32, 64, 128, 256, 512, 1024])
def test_every_period_size_true_stereo(B):
    m = ca()
    fs = 48000
    L = 9 * B + 17  # ragged: last partition partly filled
    irs = make_irs(L, fs)
    x = np.stack([O.synth_audio(B * 40, 2000 + i) for i in range(2)])
    pr = [dict(wet=0.8, dry=0.3, level=0.9, panWet=0.25, panDry=-0.5), dict(wet=0.6, dry=0.2, level=1.0, panWet=-0.4, panDry=0.3)]
    with m.Engine(period=B, max_ir_frames=L, n_ir_slots=2) as e:
        load_true_stereo(e, irs)
        for i in range(2):
            e.set_params(0, i, select=i, **pr[i])
            e.set_glide(0, i, pr[i]["wet"])
        y = e.render(x[None])[0]
    truth = O.engine_truth(x, irs, pr)
    for o in range(2):
        assert O.rel_l2(y[o], truth[o]) < 5e-6, (B, o, O.rel_l2(y[o], truth[o]))


def test_predelay():
    m = ca()
    fs, B, L = 48000, 128, 1000
    irs = make_irs(L, fs)
    x = np.stack([O.synth_audio(B * 60, 2000 + i, rms=0.5) for i in range(2)])  # loud: clamp fires
    pr = [dict(wet=1.0, dry=0.5, level=1.0, panWet=0.0, panDry=0.0)] * 2
    for pd in (0, 1, 77, 128, 1000, 8191):
        with m.Engine(period=B, max_ir_frames=L) as e:
            load_true_stereo(e, irs)
            for i in range(2):
                e.set_params(0, i, select=i, predelay=pd, **pr[i])
                e.set_glide(0, i, 1.0)
            y = e.render(x[None])[0]
        truth = O.engine_truth(x, irs, pr, predelay=pd)
        assert np.abs(truth).max() > 0.5
        for o in range(2):
            assert O.rel_l2(y[o], truth[o]) < 5e-6, (pd, o)


def test_clamp_fires():
    m = ca()
    fs, B, L = 48000, 64, 300
    h = 4.0 * O.synth_ir(L, fs, 5)
    x = O.synth_audio(B * 50, 7, rms=0.6)
    with m.Engine(period=B, max_ir_frames=L, n_in=1, n_out=1, n_ir_slots=1) as e:
        e.load_ir(0, h)
        e.set_params(0, 0, wet=1.0, dry=0.25)
        e.set_glide(0, 0, 1.0)
        y = e.render(x[None, None])[0, 0]
    truth = O.engine_truth(x[None], [[h]], [dict(wet=1.0, dry=0.25)])[0]
    wet_only = O.fft_conv(x, h)
    assert (np.abs(wet_only) > 1.0).sum() > 100  # the clamp really is exercised
    # clamping is discontinuous only in derivative: errors stay at fp32 level
    assert O.rel_l2(y, truth) < 5e-6


def test_impulse_returns_ir_full_cfg2_size():
    """Size-independent property at BASELINE's full size (48 kHz, B=256, 4 s IR, P=750):
    a unit impulse on input i must return the IR of path (i, o) on output o."""
    m = ca()
    fs, B, L = 48000, 256, 192000
    irs = make_irs(L, fs)
    n = L + 4 * B
    n = (n // B) * B
    with m.Engine(period=B, max_ir_frames=L) as e:
        load_true_stereo(e, irs)
        assert e.stats().partitions == 750
        for i in range(2):
            e.set_params(0, i, select=i, wet=1.0, dry=0.0)
            e.set_glide(0, i, 1.0)
        x = np.zeros((1, 2, n), np.float32)
        x[0, 0, 3] = 1.0
        x[0, 1, B + 5] = 0.5
        y = e.render(x)[0]
    for o in range(2):
        want = np.zeros(n)
        want[3:3 + L] += irs[0][o][: n - 3]
        want[B + 5:B + 5 + L] += 0.5 * irs[1][o][: n - B - 5]
        assert O.rel_l2(y[o], want) < 2e-6, (o, O.rel_l2(y[o], want))


def test_cfg2_vs_fp64_sampled():
    m = ca()
    fs, B, L = 48000, 256, 192000
    irs = make_irs(L, fs)
    n = B * 1200
    x = np.stack([O.synth_audio(n, 2000 + i) for i in range(2)])
    pr = [dict(wet=1.0, dry=0.0)] * 2
    with m.Engine(period=B, max_ir_frames=L) as e:
        load_true_stereo(e, irs)
        for i in range(2):
            e.set_params(0, i, select=i, **pr[i])
            e.set_glide(0, i, 1.0)
        y = e.render(x[None])[0]
    truth = O.engine_truth(x, irs, pr)
    for o in range(2):
        assert O.rel_l2(y[o], truth[o]) < 5e-6
    idx = np.random.default_rng(1).integers(L, n, 256)
    d = O.direct_conv_at(x[0], irs[0][0], idx) + O.direct_conv_at(x[1], irs[1][0], idx)
    assert O.rel_l2(y[0][idx], d) < TOL_FP64


def test_batch_instances_distinct_irs_and_active_subset():
    m = ca()
    fs, B, L, K = 48000, 64, 700, 5
    irs = [make_irs(L, fs, seed0=1000 + 8 * s) for s in range(K)]
    x = np.stack([np.stack([O.synth_audio(B * 30, 2000 + 2 * s + i) for i in range(2)]) for s in range(K)])
    pr = [dict(wet=1.0, dry=0.1)] * 2
    with m.Engine(period=B, max_ir_frames=L, n_instances=K, n_ir_slots=2 * K) as e:
        for s in range(K):
            for i in range(2):
                e.load_ir(2 * s + i, irs[s][i][0], irs[s][i][1])
                e.set_params(s, i, select=2 * s + i, **pr[i])
                e.set_glide(s, i, 1.0)
        y = e.render(x)
        for s in range(K):
            truth = O.engine_truth(x[s], irs[s], pr)
            for o in range(2):
                assert O.rel_l2(y[s, o], truth[o]) < 5e-6, (s, o)
    # subset: only the first 2 instances are processed
    with m.Engine(period=B, max_ir_frames=L, n_instances=K, n_ir_slots=2 * K) as e:
        for s in range(K):
            for i in range(2):
                e.load_ir(2 * s + i, irs[s][i][0], irs[s][i][1])
                e.set_params(s, i, select=2 * s + i, **pr[i])
                e.set_glide(s, i, 1.0)
        e.set_active(2)
        y2 = e.render(x[:2])
        assert np.array_equal(y2, y[:2])


def test_shared_ir_bank_and_select():
    """Two inputs selecting the same / swapped bank slots (reference: cc[i].value.select)."""
    m = ca()
    fs, B, L = 48000, 64, 500
    irs = make_irs(L, fs)
    x = np.stack([O.synth_audio(B * 30, 2000 + i) for i in range(2)])
    with m.Engine(period=B, max_ir_frames=L) as e:
        load_true_stereo(e, irs)
        e.set_params(0, 0, select=1, wet=1.0, dry=0.0)
        e.set_params(0, 1, select=0, wet=1.0, dry=0.0)
        e.set_glide(0, 0, 1.0)
        e.set_glide(0, 1, 1.0)
        y = e.render(x[None])[0]
    truth = O.engine_truth(x, [irs[1], irs[0]], [dict(wet=1.0)] * 2)
    for o in range(2):
        assert O.rel_l2(y[o], truth[o]) < 5e-6


def test_split_graph_and_variants_agree():
    m = ca()
    fs, B, L = 48000, 256, 256 * 37 + 5
    irs = make_irs(L, fs)
    x = np.stack([O.synth_audio(B * 60, 2000 + i) for i in range(2)])

    def run(**kw):
        with m.Engine(period=B, max_ir_frames=L, **kw) as e:
            load_true_stereo(e, irs)
            for i in range(2):
                e.set_params(0, i, select=i, wet=1.0, dry=0.2)
                e.set_glide(0, i, 1.0)
            y = e.render(x[None])[0]
            return y, e.stats()

    y1, s1 = run(mac_split=1)
    y8, s8 = run(mac_split=8)
    yg, _ = run(mac_split=8, flags=m.FLAG_GRAPH)
    yp, sp = run(mac_split=8, flags=m.FLAG_PROFILE)
    ys, _ = run(mac_split=3, flags=m.FLAG_STREAMING)
    assert s1.mac_split == 1 and s8.mac_split > 1
    assert np.array_equal(y8, yg)          # graph replay == plain launches, bitwise
    assert np.array_equal(y8, yp)
    assert sp.mac_us > 0 and sp.fwd_us > 0 and sp.inv_us > 0
    assert O.rel_l2(y8, y1) < 1e-6         # different summation order only
    assert O.rel_l2(ys, y1) < 1e-6
    truth = O.engine_truth(x, irs, [dict(wet=1.0, dry=0.2)] * 2)
    assert O.rel_l2(y1[0], truth[0]) < 5e-6


def test_partition_range_shards_sum_to_whole():
    """a long IR split by partition range; partial outputs sum to the 1-engine
    result (this is the collective-free check of the multi-GPU sharding math)."""
    m = ca()
    fs, B, L = 48000, 128, 128 * 40
    irs = make_irs(L, fs)
    x = np.stack([O.synth_audio(B * 90, 2000 + i) for i in range(2)])

    def run(pb, pc):
        with m.Engine(period=B, max_ir_frames=L, part_begin=pb, part_count=pc) as e:
            load_true_stereo(e, irs)
            for i in range(2):
                e.set_params(0, i, select=i, wet=1.0, dry=0.0)
                e.set_glide(0, i, 1.0)
            return e.render(x[None])[0]

    whole = run(0, 0)
    parts = run(0, 10).astype(np.float64) + run(10, 17) + run(27, 13)
    assert O.rel_l2(parts, whole) < 1e-6
    truth = O.engine_truth(x, irs, [dict(wet=1.0)] * 2)
    assert O.rel_l2(parts[0], truth[0]) < 5e-6


def test_wet_glide_fade_in_matches_formula():
    """Freshly started engine fades the wet path in as 1 - 0.8^k (conv.cu:27 with vsteps = 0)."""
    m = ca()
    fs, B, L = 48000, 64, 64
    h = np.zeros(L, np.float32)
    h[0] = 1.0  # identity IR: output = g_t * x
    x = np.full(B * 30, 0.25, np.float32)
    with m.Engine(period=B, max_ir_frames=L, n_in=1, n_out=1, n_ir_slots=1) as e:
        e.load_ir(0, h)
        e.set_params(0, 0, wet=1.0, dry=0.0)
        y = e.render(x[None, None])[0, 0]
    g = 0.0
    for t in range(30):
        g = g + (1.0 - g) / 5.0
        assert np.allclose(y[t * B:(t + 1) * B], 0.25 * g, rtol=2e-6, atol=1e-7), t


def test_error_codes():
    m = ca()
    with m.Engine(period=64, max_ir_frames=100, n_ir_slots=2) as e:
        with pytest.raises(m.CaError) as ei:
            e.set_params(0, 0, select=1)  # slot never loaded: reference would fault (conv.cu:340)
        assert ei.value.code == -4
        e.load_ir(0, np.ones(10, np.float32), np.ones(10, np.float32))
        e.set_params(0, 0, select=0)
        with pytest.raises(m.CaError) as ei:
            e.set_params(0, 0, select=7)
        assert ei.value.code == -1
        x = np.zeros((1, 2, 32), np.float32)
        out = np.zeros((1, 2, 32), np.float32)
        rc = m.lib().ca_process(e._h, x.ctypes.data, out.ctypes.data, 32)
        assert rc == -6
    with pytest.raises(m.CaError):
        m.Engine(period=100, max_ir_frames=100)  # not a power of two
    with pytest.raises(m.CaError):
        m.Engine(period=64, max_ir_frames=100, n_in=3)


def test_pinned_buffers_and_stats():
    m = ca()
    fs, B, L, K = 48000, 256, 2048, 3
    h = O.synth_ir(L, fs, 3)
    with m.Engine(period=B, max_ir_frames=L, n_instances=K, n_ir_slots=1) as e:
        e.load_ir(0, h, h)
        for s in range(K):
            for i in range(2):
                e.set_params(s, i, wet=1.0, dry=0.0)
                e.set_glide(s, i, 1.0)
        pin = m.PinnedArray((K, 2, B))
        pout = m.PinnedArray((K, 2, B))
        x = O.synth_audio(K * 2 * B * 20, 9).reshape(20, K, 2, B)
        ys = []
        for t in range(20):
            pin.array[...] = x[t]
            e.process_raw(pin.ptr, pout.ptr)
            ys.append(pout.array.copy())
        st = e.stats()
        assert st.periods == 20 and st.gpu_launches >= 61 and st.p50_us > 0
        y = np.concatenate(ys, axis=-1)
        xin = np.concatenate(list(x), axis=-1)
        truth = O.engine_truth(xin[1], [[h, h], [h, h]], [dict(wet=1.0)] * 2)
        assert O.rel_l2(y[1, 0], truth[0]) < 5e-6
        pin.free()
        pout.free()


# ------------------------------------------------------------------------------------------
# against the real reference (oracle/_ref = unmodified conv.cu + cuFFT), live on this GPU
# ------------------------------------------------------------------------------------------
needs_ref = pytest.mark.skipif(not refgpu.available(), reason="oracle/_ref/libref_conv.so not built")


@needs_ref
@pytest.mark.parametrize("N,B", [(4096, 64), (65536, 256)])
def test_vs_reference_conv_cu(N, B):
    """Parity protocol of SURVEY 8(c): DC/Nyquist-free IRs, >= 80 silent warm-up periods,
    unclipped levels, predelay 0.  Reference = its own conv.cu through prepare()/onProcess()."""
    m = ca()
    fs = 48000
    L = N - B
    irs = make_irs(L, fs)
    warm = 100
    x = np.stack([np.concatenate([np.zeros(warm * B, np.float32), O.synth_audio(B * 150, 2000 + i)]) for i in range(2)])
    ref = refgpu.RefGpu(N)
    ref.prepare(0, irs[0][0], irs[0][1], B)
    ref.prepare(1, irs[1][0], irs[1][1], B)
    ref.set_cc(0, select=0, wet=1.0, dry=0.0)
    ref.set_cc(1, select=1, wet=1.0, dry=0.0)
    rl, rr = ref.render(x[0], x[1], B)
    with m.Engine(period=B, max_ir_frames=L) as e:
        load_true_stereo(e, irs)
        for i in range(2):
            e.set_params(0, i, select=i, wet=1.0, dry=0.0)
        y = e.render(x[None])[0]
    sl = slice(warm * B, None)
    truth = O.engine_truth(x, irs, [dict(wet=1.0)] * 2)
    diag = dict(ours_vs_truth=[O.rel_l2(y[o][sl], truth[o][sl]) for o in range(2)],
                ref_vs_truth=[O.rel_l2(r[sl], truth[o][sl]) for o, r in enumerate((rl, rr))])
    bad = np.nonzero(np.abs(y[1] - truth[1]) > 1e-4)[0]
    if len(bad):
        diag["ours_R_bad"] = dict(n=len(bad), first_period=int(bad[0] // B), periods=sorted(set((bad // B).tolist()))[:10],
                                  in_block=sorted(set((bad % B).tolist()))[:10])
    assert O.rel_l2(y[0][sl], rl[sl]) < TOL_REF, diag
    assert O.rel_l2(y[1][sl], rr[sl]) < TOL_REF, diag


@needs_ref
def test_vs_reference_defaults_pan_predelay_fade_in():
    """Reference defaults (wet = dry = 0.5) plus pan / level / predelay, compared from the very
    first period: exercises the wet fade-in glide, the pan law, predelay and the dry mix."""
    m = ca()
    fs, N, B, pd = 48000, 16384, 256, 300
    L = N - B - pd - B
    irs = make_irs(L, fs)
    x = np.stack([O.synth_audio(B * 200, 2000 + i) for i in range(2)])
    cc = [dict(select=0, predelay=pd, panWet=0.3, panDry=-0.2, level=0.8), dict(select=1, panWet=-0.5, panDry=0.4)]
    ref = refgpu.RefGpu(N)
    ref.prepare(0, irs[0][0], irs[0][1], B)
    ref.prepare(1, irs[1][0], irs[1][1], B)
    ref.set_cc(0, **cc[0])
    ref.set_cc(1, **cc[1])
    rl, rr = ref.render(x[0], x[1], B)
    with m.Engine(period=B, max_ir_frames=L) as e:
        load_true_stereo(e, irs)
        e.set_params(0, 0, **cc[0])
        e.set_params(0, 1, **cc[1])
        y = e.render(x[None])[0]
    assert O.rel_l2(y[0], rl) < TOL_REF, O.rel_l2(y[0], rl)
    assert O.rel_l2(y[1], rr) < TOL_REF, O.rel_l2(y[1], rr)


@needs_ref
def test_restatement_pinned_by_live_reference():
    """oracle/refconv.c (the CPU restatement) against the compiled reference on quirk-exposing
    input: white-noise IRs WITHOUT the DC/Nyquist correction, fade-in, clamp active."""
    fs, N, B = 48000, 8192, 128
    L = N - B
    irs = [[O.synth_ir(L, fs, 50 + 2 * i + o, parity_safe=False) for o in range(2)] for i in range(2)]
    x = np.stack([O.synth_audio(B * 120, 2000 + i, rms=0.4) for i in range(2)])
    ref = refgpu.RefGpu(N)
    cpu = O.RefConv(N)
    for i in range(2):
        ref.prepare(i, irs[i][0], irs[i][1], B)
        cpu.prepare(i, irs[i][0], irs[i][1], B)
        ref.set_cc(i, select=i, wet=0.9, dry=0.3, panWet=0.2 - 0.5 * i)
        cpu.set_cc(i, select=i, wet=0.9, dry=0.3, panWet=0.2 - 0.5 * i)
    rl, rr = ref.render(x[0], x[1], B)
    cl, cr = cpu.render(x[0], x[1], B)
    assert O.rel_l2(cl, rl) < TOL_REF, O.rel_l2(cl, rl)
    assert O.rel_l2(cr, rr) < TOL_REF, O.rel_l2(cr, rr)


@needs_ref
@pytest.mark.parametrize("N,B,pd,tiers,graph", [
    (8192, 128, 0, None, False),          # uniform, no predelay
    (16384, 256, 301, None, True),        # odd predelay (the Nyquist term changes sign), CUDA graph per period
    (65536, 256, 8191, "auto", False),    # non-uniform partitioning, maximum predelay
])
def test_ref_quirks_match_reference_on_unconstrained_irs(N, B, pd, tiers, graph):
    """CA_FLAG_REF_QUIRKS: raw white-noise IRs WITHOUT the DC/Nyquist correction, reference defaults
    plus pan / level / predelay, compared from the first period (fade-in glide included).  The exact engine differs
    from conv.cu by the DC / Nyquist terms (> 1e-4); with the flag it matches to the reference tolerance."""
    m = ca()
    fs = 48000
    L = N - 2 * B - pd if pd < 4096 else 20000
    irs = [[O.synth_ir(L, fs, 60 + 2 * i + o, parity_safe=False) for o in range(2)] for i in range(2)]
    x = np.stack([O.synth_audio(B * 300, 2000 + i) for i in range(2)])
    cc = [dict(select=0, predelay=pd, wet=0.8, dry=0.4, panWet=0.3, panDry=-0.2, level=0.8), dict(select=1, wet=0.6, dry=0.2, panWet=-0.5, panDry=0.4)]
    ref = refgpu.RefGpu(N)
    for i in range(2):
        ref.prepare(i, irs[i][0], irs[i][1], B)
        ref.set_cc(i, **cc[i])
    rl, rr = ref.render(x[0], x[1], B)
    assert max(np.abs(rl).max(), np.abs(rr).max()) < 0.99      # the clamp (which the reference applies to partial sums) stays out of it

    def go(flags):
        with m.Engine(period=B, max_ir_frames=L, tiers=tiers, flags=flags, ref_fft_size=N) as e:
            load_true_stereo(e, irs)
            for i in range(2):
                e.set_params(0, i, **cc[i])
            return e.render(x[None])[0]

    y = go(m.FLAG_REF_QUIRKS | (m.FLAG_GRAPH if graph else 0))
    y_exact = go(0)
    eq = [O.rel_l2(y[0], rl), O.rel_l2(y[1], rr)]
    ee = [O.rel_l2(y_exact[0], rl), O.rel_l2(y_exact[1], rr)]
    assert max(ee) > 1e-4, ee          # the quirks are audible in the difference ...
    assert max(eq) < TOL_REF, (eq, ee)  # ... and reproduced with the flag


@needs_ref
def test_ref_quirks_follow_an_ir_switch_and_a_batch():
    """The DC / Nyquist terms glide with the live IR (conv.cu:15-32): mid-run `select` change on unconstrained IRs;
    second instance of the same engine with other parameters (chunk arguments of the quirk kernel)."""
    m = ca()
    fs, N, B = 48000, 8192, 128
    L = 3000
    irs = [[O.synth_ir(L, fs, 80 + 2 * i + o, parity_safe=False) for o in range(2)] for i in range(3)]
    x = np.stack([O.synth_audio(B * 260, 2300 + i) for i in range(2)])
    ccs = [[dict(select=0, wet=1.0, dry=0.0, speed=40), dict(select=1, wet=1.0, dry=0.0, speed=40)],
           [dict(select=2, wet=0.5, dry=0.5, predelay=77, panWet=0.4), dict(select=0, wet=0.7, dry=0.1, level=0.6)]]
    refs = []
    for cc in ccs:
        ref = refgpu.RefGpu(N)
        for s in range(3):
            ref.prepare(s, irs[s][0], irs[s][1], B)
        for i in range(2):
            ref.set_cc(i, **cc[i])
        refs.append(ref)
    periods = x.shape[1] // B
    out_ref = np.zeros((2, 2, periods * B), np.float32)
    out = np.zeros((2, 2, periods * B), np.float32)
    with m.Engine(period=B, max_ir_frames=L, n_instances=2, n_ir_slots=3, max_voices=3, flags=m.FLAG_REF_QUIRKS, ref_fft_size=N) as e:
        for s in range(3):
            e.load_ir(s, irs[s][0], irs[s][1])
        for k in range(2):
            for i in range(2):
                e.set_params(k, i, **ccs[k][i])
        for t in range(periods):
            if t == 100:   # instance 0 cross-fades input 0 to IR 2 over 40 periods (handleCC: vsteps = speed, conv.cu:261)
                refs[0].set_cc(0, select=2, vsteps=40)
                e.set_params(0, 0, **dict(ccs[0][0], select=2, vsteps=40))
            blk = x[:, t * B:(t + 1) * B]
            for k in range(2):
                l, r = refs[k].process(blk[0], blk[1])
                out_ref[k, 0, t * B:(t + 1) * B] = l
                out_ref[k, 1, t * B:(t + 1) * B] = r
            out[:, :, t * B:(t + 1) * B] = e.process(np.stack([blk, blk]))
    for k in range(2):
        for o in range(2):
            assert O.rel_l2(out[k, o], out_ref[k, o]) < TOL_REF, (k, o, O.rel_l2(out[k, o], out_ref[k, o]))


def test_chunked_host_pipeline_matches_single_stream():
    """>= 512 instances: ca_process pipelines H2D | kernels | D2H in instance chunks on three streams;
    results must equal the device-resident single-stream path bit for bit."""
    m = ca()
    fs, B, L, K = 48000, 64, 300, 1100
    h = [O.synth_ir(L, fs, 40 + j) for j in range(4)]
    rng = np.random.default_rng(5)
    x = (rng.standard_normal((12, K, 2, B)) * 0.1).astype(np.float32)

    def engine():
        e = m.Engine(period=B, max_ir_frames=L, n_instances=K, n_ir_slots=2)
        e.load_ir(0, h[0], h[1])
        e.load_ir(1, h[2], h[3])
        for s in range(K):
            for i in range(2):
                e.set_params(s, i, select=(s + i) % 2, wet=1.0, dry=0.3, panWet=0.01 * (s % 50))
                e.set_glide(s, i, 1.0)
        return e

    with engine() as e:
        ya = np.stack([e.process(x[t]) for t in range(12)])          # chunked host path
    import torch
    with engine() as e:
        xd = torch.from_numpy(x).cuda()
        yd = torch.empty(12, K, 2, B, device="cuda")
        for t in range(12):
            e.process_device(xd[t].data_ptr(), yd[t].data_ptr())      # one stream, device buffers
        e.sync()
        yb = yd.cpu().numpy()
    assert np.array_equal(ya, yb)
    s = 777
    truth = O.engine_truth(np.concatenate(list(x[:, s]), axis=-1), [[h[2 * ((s + i) % 2)], h[2 * ((s + i) % 2) + 1]] for i in range(2)],
                           [dict(wet=1.0, dry=0.3, panWet=0.01 * (s % 50))] * 2)
    assert O.rel_l2(np.concatenate(list(ya[:, s, 0]), axis=-1), truth[0]) < 5e-6


@pytest.mark.parametrize("n_in,n_out", [(1, 2), (2, 1), (1, 1)])
@pytest.mark.parametrize("tiers", [None, [(64, 8), (512, 0)]])
def test_channel_layouts(n_in, n_out, tiers):
    """mono-in / stereo-out, stereo-in / mono-out and mono: uniform and tiered."""
    m = ca()
    fs, B, L = 48000, 64, 64 * 8 + 512 * 2 + 31
    irs = [[O.synth_ir(L, fs, 500 + 2 * i + o) for o in range(n_out)] for i in range(n_in)]
    x = np.stack([O.synth_audio(B * 70, 600 + i) for i in range(n_in)])
    pr = [dict(wet=0.9, dry=0.25, level=0.7, panWet=0.3, panDry=-0.6), dict(wet=0.4, dry=0.5, level=1.0, panWet=-0.2, panDry=0.1)][:n_in]
    with m.Engine(period=B, max_ir_frames=L, n_in=n_in, n_out=n_out, n_ir_slots=n_in, tiers=tiers) as e:
        for i in range(n_in):
            e.load_ir(i, irs[i][0], irs[i][1] if n_out == 2 else None)
            e.set_params(0, i, select=i, predelay=9, **pr[i])
            e.set_glide(0, i, pr[i]["wet"])
        y = e.render(x[None])[0]
    truth = O.engine_truth(x, irs, pr, predelay=9)
    for o in range(n_out):
        assert O.rel_l2(y[o], truth[o]) < 5e-6, (n_in, n_out, o, O.rel_l2(y[o], truth[o]))


@pytest.mark.parametrize("tiers", [None, "auto"])
def test_edge_cases_tiny_truncated_and_zero_irs_silence(tiers):
    """Ragged / degenerate inputs: a 1-frame IR (a pure gain), an IR one frame longer than a partition,
    an IR longer than the engine's capacity (truncated like conv.cu:239), an all-zero IR, silent input,
    and an input block of denormal-sized values."""
    m = ca()
    B, cap = 64, 64 * 40
    rng = np.random.default_rng(3)
    x = (rng.standard_normal((2, B * 70)) * 0.2).astype(np.float32)
    x[:, B * 30:B * 40] = 0.0                       # ten periods of silence in the middle
    x[:, B * 50:B * 51] = 1e-39                     # denormals
    long_ir = (rng.standard_normal((2, cap + 500)) * 0.05).astype(np.float32)
    cases = {
        "one_frame": np.array([[0.5], [-0.25]], np.float32),
        "partition_plus_one": (rng.standard_normal((2, B + 1)) * 0.1).astype(np.float32),
        "truncated": long_ir,
        "zeros": np.zeros((2, 300), np.float32),
    }
    pr = [dict(wet=1.0, dry=0.0, level=1.0, panWet=0.0, panDry=0.0)] * 2
    for name, h in cases.items():
        with m.Engine(period=B, max_ir_frames=cap, tiers=tiers) as e:
            for i in range(2):
                e.load_ir(i, h[0], h[1])
                e.set_params(0, i, select=i, **pr[i])
                e.set_glide(0, i, 1.0)
            y = e.render(x[None])[0]
        hh = h[:, :cap]
        truth = O.engine_truth(x, [[hh[0], hh[1]], [hh[0], hh[1]]], pr)
        assert np.isfinite(y).all(), name
        if name == "zeros":
            assert np.abs(y).max() == 0.0
        else:
            for o in range(2):
                assert O.rel_l2(y[o], truth[o]) < 5e-6, (name, o, O.rel_l2(y[o], truth[o]))


def test_l2_persist_window_flag_is_bit_identical():
    """CA_FLAG_L2_PERSIST only changes cache policy (access-policy window over [IR spectra | delay lines])."""
    m = ca()
    fs, B, L = 48000, 256, 256 * 120 + 9
    irs = make_irs(L, fs)
    x = np.stack([O.synth_audio(B * 150, 2100 + i) for i in range(2)])

    def go(flags):
        with m.Engine(period=B, max_ir_frames=L, flags=flags) as e:
            load_true_stereo(e, irs)
            for i in range(2):
                e.set_params(0, i, select=i, wet=0.8, dry=0.1)
                e.set_glide(0, i, 0.8)
            return e.render(x[None])[0]

    assert np.array_equal(go(m.FLAG_GRAPH), go(m.FLAG_GRAPH | m.FLAG_L2_PERSIST))


@pytest.mark.parametrize("B,n_in,n_out", [(256, 2, 2), (64, 2, 1), (128, 1, 2)])
def test_persistent_kernel_mode_matches_graph_mode_and_fp64(B, n_in, n_out):
    """CA_FLAG_PERSISTENT: one resident cooperative kernel + a mailbox in mapped host memory instead of launches.
    Same audio as the launched pipeline (different MAC split => fp32 rounding only) and as fp64, through a
    predelay, pans, an IR cross-fade, an IR load (stops and relaunches the kernel) and an idle gap long enough
    for the kernel to leave by itself."""
    import time
    m = ca()
    fs, L = 48000, B * 40 + 9
    irs = make_irs(L, fs, n_in=2, n_out=2, seed0=3100)
    extra = make_irs(L, fs, n_in=1, n_out=2, seed0=3200)[0]
    nper = 160
    x = np.stack([O.synth_audio(B * nper, 3300 + i) for i in range(n_in)])
    pr = dict(wet=0.8, dry=0.3, level=0.9, panWet=0.25, panDry=-0.5)

    def go(flags):
        with m.Engine(period=B, max_ir_frames=L, n_in=n_in, n_out=n_out, n_ir_slots=3, flags=flags, max_voices=2) as e:
            for i in range(2):
                e.load_ir(i, irs[i][0], irs[i][1] if n_out == 2 else None)
            for i in range(n_in):
                e.set_params(0, i, select=i, predelay=77, **pr)
                e.set_glide(0, i, pr["wet"])
            out = np.zeros((n_out, B * nper), np.float32)
            for t in range(nper):
                if t == 50:
                    e.load_ir(2, extra[0], extra[1] if n_out == 2 else None)
                if t == 60:
                    e.set_params(0, 0, select=2, predelay=77, vsteps=15, **pr)      # cross-fade input 0 to the new IR
                if t == 100 and (flags & m.FLAG_PERSISTENT):
                    time.sleep(1.6)                                                  # the kernel leaves after ~1.1 s of silence
                out[:, t * B:(t + 1) * B] = e.process(x[None, :, t * B:(t + 1) * B])[0]
            st = e.stats()
            return out, st

    yp, sp = go(m.FLAG_PERSISTENT)
    yg, sg = go(m.FLAG_GRAPH)
    assert sp.gpu_launches < 40 < sg.gpu_launches          # a handful of (re)launches instead of 3 per period
    for o in range(n_out):
        assert O.rel_l2(yp[o], yg[o]) < 2e-6, (o, O.rel_l2(yp[o], yg[o]))
    # fp64 truth for the part before the switch (static parameters)
    hs = [[irs[i][o] for o in range(n_out)] for i in range(n_in)]
    truth = O.engine_truth(x[:, :60 * B], hs, [pr] * n_in, predelay=77)
    for o in range(n_out):
        assert O.rel_l2(yp[o, :60 * B], truth[o]) < 5e-6


def test_persistent_kernel_mode_rejects_what_it_cannot_do():
    m = ca()
    for kw in (dict(n_instances=2), dict(tiers=[(64, 8), (512, 0)]), dict(period=512, max_ir_frames=4096), dict(flags_extra=m.FLAG_GRAPH)):
        fl = m.FLAG_PERSISTENT | kw.pop("flags_extra", 0)
        args = dict(period=64, max_ir_frames=64 * 8 + 512 * 3, flags=fl)
        args.update(kw)
        with pytest.raises(m.CaError) as ei:
            m.Engine(**args)
        assert ei.value.code == -5, kw


@pytest.mark.parametrize("seed", range(10))
def test_random_configurations_vs_fp64(seed):
    """Seeded random sweep over the configuration space (period, IR length, channel layout, tier plan, flags,
    predelay, pans, instance count) against the fp64 oracle: the parity gate should not depend on the handful of
    shapes the other tests happen to use."""
    m = ca()
    rng = np.random.default_rng(1234 + seed)
    B = int(rng.choice([32, 64, 128, 256, 256, 512]))
    n_in, n_out = int(rng.integers(1, 3)), int(rng.integers(1, 3))
    K = int(rng.choice([1, 1, 2, 5, 19]))
    mode = rng.choice(["uniform", "auto", "auto", "explicit"])
    if mode == "explicit" and B <= 256:
        S1 = int(max(256, B * int(rng.choice([4, 8]))))
        p0 = S1 // B + int(rng.integers(0, 3))
        tiers = [(B, p0), (S1, 0)]
        L = B * p0 + S1 * int(rng.integers(2, 6)) - int(rng.integers(0, B))
    else:
        tiers = None if mode != "auto" else "auto"
        L = int(rng.integers(3 * B, 60 * B)) if mode != "auto" else int(rng.integers(20 * B, 200 * B))
    flags = 0
    if K == 1 and rng.random() < 0.5:
        flags |= m.FLAG_GRAPH
    if tiers is not None and rng.random() < 0.3 and mode == "auto":
        flags |= m.FLAG_ASYNC_TIERS
    if rng.random() < 0.3:
        flags |= m.FLAG_LEGACY_FFT
    nper = int(max(40, (L // B) + 30))
    pd = int(rng.choice([0, 0, 1, 63, 64, 777, 8191]))
    irs = [[[O.synth_ir(L, 48000, 40000 + 100 * seed + 8 * s + 2 * i + o) for o in range(n_out)] for i in range(n_in)] for s in range(min(K, 3))]
    x = np.stack([np.stack([O.synth_audio(B * nper, 41000 + 100 * seed + 2 * s + i) for i in range(n_in)]) for s in range(K)])
    prs = [[dict(wet=float(rng.uniform(0.2, 1.0)), dry=float(rng.uniform(0, 0.5)), level=float(rng.uniform(0.5, 1.0)),
                 panWet=float(rng.uniform(-1, 1)), panDry=float(rng.uniform(-1, 1))) for i in range(n_in)] for s in range(K)]
    with m.Engine(period=B, max_ir_frames=L, n_instances=K, n_in=n_in, n_out=n_out, n_ir_slots=n_in * min(K, 3), tiers=tiers, flags=flags) as e:
        for s in range(min(K, 3)):
            for i in range(n_in):
                e.load_ir(n_in * s + i, irs[s][i][0], irs[s][i][1] if n_out == 2 else None)
        for s in range(K):
            for i in range(n_in):
                e.set_params(s, i, select=n_in * (s % 3 if K >= 3 else s) + i, predelay=pd, **prs[s][i])
                e.set_glide(s, i, prs[s][i]["wet"])
        y = e.render(x)
        st = e.stats()
    for s in sorted(set([0, K // 2, K - 1])):
        truth = O.engine_truth(x[s], irs[s % 3 if K >= 3 else s], prs[s], predelay=pd)
        for o in range(n_out):
            err = O.rel_l2(y[s, o], truth[o])
            assert err < 5e-6, dict(seed=seed, B=B, L=L, K=K, n_in=n_in, n_out=n_out, mode=str(mode), flags=flags, pd=pd, tiers=[int(st.tier_block[j]) for j in range(st.n_tiers)], s=s, o=o, err=err)


def test_reset_restarts_like_a_new_engine():
    """ca_reset: history dropped, fade-in glide from silence again -- the second render equals the first bit for bit
    (uniform and tiered; the host mirror relies on it after its silent warm-up period)."""
    m = ca()
    fs, B, L = 48000, 256, 256 * 4 + 1024 * 3 + 4096 * 2 - 7
    irs = make_irs(L, fs, seed0=3100)
    x = np.stack([O.synth_audio(B * 70, 3200 + i) for i in range(2)])
    for tiers in (None, [(256, 4), (1024, 3), (4096, 0)]):
        with m.Engine(period=B, max_ir_frames=L, tiers=tiers) as e:
            load_true_stereo(e, irs)
            for i in range(2):
                e.set_params(0, i, select=i, wet=0.7, dry=0.2, predelay=50 * i)
            y1 = e.render(x[None])[0]
            e.reset()
            y2 = e.render(x[None])[0]
        assert np.abs(y1).max() > 0.01
        if tiers is None:
            assert np.array_equal(y1, y2)
        else:   # the period counter runs on, so the long tiers' blocks close at other offsets of the input: equal to fp32 rounding
            assert max(O.rel_l2(y2[o], y1[o]) for o in range(2)) < 1e-6, [O.rel_l2(y2[o], y1[o]) for o in range(2)]
