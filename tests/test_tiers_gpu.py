"""Non-uniform partitioning (BASELINE configs[2]: low-latency 64-frame period, 10 s IR, small head
partitions + large tail partitions) against the FP64 oracle and against the uniform engine."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu
FS = 48000


def ca():
    import cuda_audio_b200 as m
    return m


def irs2x2(L, seed0=1000):
    return [[O.synth_ir(L, FS, seed0 + 2 * i + o) for o in range(2)] for i in range(2)]


def run(m, B, L, irs, x, pr, predelay=0, **kw):
    with m.Engine(period=B, max_ir_frames=L, **kw) as e:
        for i in range(2):
            e.load_ir(i, irs[i][0], irs[i][1])
            e.set_params(0, i, select=i, predelay=predelay, **pr[i])
            e.set_glide(0, i, pr[i]["wet"])
        y = e.render(x[None])[0]
        return y, e.stats()


@pytest.mark.parametrize("B,L,tiers", [
    (64, 64 * 8 + 512 * 3 + 100, [(64, 8), (512, 0)]),
    (64, 20000, [(64, 8), (512, 7), (4096, 0)]),
    (32, 9000, [(32, 8), (256, 8), (2048, 0)]),
    (128, 30000, [(128, 4), (512, 3), (2048, 0)]),
    (256, 70000, "auto"),
])
def test_tiers_match_fp64_and_uniform(B, L, tiers):
    m = ca()
    irs = irs2x2(L)
    n = ((L + 6 * 16384) // B) * B
    x = np.stack([O.synth_audio(n, 2000 + i) for i in range(2)])
    pr = [dict(wet=0.9, dry=0.2, level=0.8, panWet=0.2, panDry=-0.3), dict(wet=0.7, dry=0.1, level=1.0, panWet=-0.4, panDry=0.1)]
    yt, st = run(m, B, L, irs, x, pr, predelay=37, tiers=tiers)
    assert st.n_tiers >= 2
    yu, su = run(m, B, L, irs, x, pr, predelay=37)
    truth = O.engine_truth(x, irs, pr, predelay=37)
    for o in range(2):
        assert O.rel_l2(yt[o], truth[o]) < 5e-6, (o, O.rel_l2(yt[o], truth[o]))
        assert O.rel_l2(yt[o], yu[o]) < 2e-6
    assert st.mac_bytes_amortized < su.mac_bytes_amortized


def test_cfg3_full_size_impulse_and_noise():
    """configs[2] at full size: 64-frame period, 10 s IR (480 000 frames), tiers 64/512/4096/16384.
    Size-independent property: an impulse returns the IR; plus noise vs the fp64 FFT convolution."""
    m = ca()
    B, L = 64, 480000
    irs = irs2x2(L)
    n = ((L + 40000) // B) * B
    x = np.zeros((2, n), np.float32)
    x[0, 5] = 1.0
    x[1, 70] = -0.5
    pr = [dict(wet=1.0, dry=0.0)] * 2
    y, st = run(m, B, L, irs, x, pr, tiers="auto", flags=m.FLAG_GRAPH)
    assert list(st.tier_block[:st.n_tiers]) == [64, 512, 4096, 16384]
    for o in range(2):
        want = np.zeros(n)
        want[5:5 + L] += irs[0][o][:n - 5]
        want[70:70 + L] += -0.5 * irs[1][o][:n - 70]
        assert O.rel_l2(y[o], want) < 2e-6, (o, O.rel_l2(y[o], want))
    xn = np.stack([O.synth_audio(n, 2000 + i) for i in range(2)])
    yn, _ = run(m, B, L, irs, xn, pr, tiers="auto")
    truth = O.engine_truth(xn, irs, pr)
    for o in range(2):
        assert O.rel_l2(yn[o], truth[o]) < 5e-6
    idx = np.random.default_rng(3).integers(L // 2, n, 128)
    d = O.direct_conv_at(xn[0], irs[0][0], idx) + O.direct_conv_at(xn[1], irs[1][0], idx)
    assert O.rel_l2(yn[0][idx], d) < 1e-4


def test_tiers_graph_batch_and_crossfade():
    m = ca()
    B, L, K = 64, 12000, 3
    tiers = [(64, 8), (512, 7), (4096, 0)]
    irs = [irs2x2(L, 1000 + 8 * s) for s in range(K)]
    n = B * 900
    x = np.stack([np.stack([O.synth_audio(n, 3000 + 2 * s + i) for i in range(2)]) for s in range(K)])

    def go(flags, tiers):
        with m.Engine(period=B, max_ir_frames=L, n_instances=K, n_ir_slots=2 * K, tiers=tiers, flags=flags, max_voices=3) as e:
            for s in range(K):
                for i in range(2):
                    e.load_ir(2 * s + i, irs[s][i][0], irs[s][i][1])
                    e.set_params(s, i, select=2 * s + i, wet=1.0, dry=0.0)
                    e.set_glide(s, i, 1.0)
            out = np.empty((K, 2, n), np.float32)
            for t in range(n // B):
                if t == 300:   # instance 1, input 0 switches to the other IR of its pair with a 30-period glide
                    e.set_params(1, 0, select=3, wet=1.0, dry=0.0, vsteps=30)
                out[:, :, t * B:(t + 1) * B] = e.process(x[:, :, t * B:(t + 1) * B])
            return out

    yt = go(0, tiers)
    yg = go(m.FLAG_GRAPH, tiers)
    yu = go(0, None)
    assert np.array_equal(yt, yg)                      # graph replay (per fire-mask graphs) == plain launches
    assert O.rel_l2(yt, yu) < 2e-6                     # tiers == uniform, including the cross-fade of instance 1
    truth = O.engine_truth(x[0], irs[0], [dict(wet=1.0)] * 2)
    assert O.rel_l2(yt[0, 0], truth[0]) < 5e-6
    # instance 1 really cross-faded: before the switch it equals IR pair (2,3), long after it input 0 uses IR 3
    t1 = O.engine_truth(x[1], irs[1], [dict(wet=1.0)] * 2)
    assert O.rel_l2(yt[1, 0][:300 * B], t1[0][:300 * B]) < 5e-6
    assert O.rel_l2(yt[1, 0][350 * B:], t1[0][350 * B:]) > 1e-2


def test_invalid_tier_configs():
    m = ca()
    with pytest.raises(m.CaError):
        m.Engine(period=64, max_ir_frames=5000, tiers=[(64, 4), (512, 0)])      # offset 256 < block 512
    with pytest.raises(m.CaError):
        m.Engine(period=64, max_ir_frames=5000, tiers=[(128, 8), (1024, 0)])    # tier 0 != period
    with pytest.raises(m.CaError):
        m.Engine(period=64, max_ir_frames=5000, tiers=[(64, 8), (512, 0)], part_begin=0, part_count=4)


def test_staggered_batch_more_instances_than_tier_period():
    """Long tiers are phase-staggered over instances (instance s closes its tier-j block when
    (t_end + s mod m) % m == 0).  19 instances, tier periods 8 and 32: every residue class holds
    0..3 instances; each instance must still equal its own convolution."""
    m = ca()
    B, L, K = 64, 64 * 8 + 512 * 3 + 2048 * 2 - 100, 19
    tiers = [(64, 8), (512, 3), (2048, 0)]
    hs = [[O.synth_ir(L, FS, 7000 + 2 * s + o) for o in range(2)] for s in range(K)]
    n = B * 260
    x = np.stack([np.stack([O.synth_audio(n, 8000 + 2 * s + i) for i in range(2)]) for s in range(K)])
    for flags in (0, m.FLAG_GRAPH):
        with m.Engine(period=B, max_ir_frames=L, n_instances=K, n_ir_slots=K, tiers=tiers, flags=flags) as e:
            for s in range(K):
                e.load_ir(s, hs[s][0], hs[s][1])
                for i in range(2):
                    e.set_params(s, i, select=s, wet=1.0, dry=0.0)
                    e.set_glide(s, i, 1.0)
            y = e.render(x)
            e.set_active(11)                      # shrink the batch mid-run: phases are per instance, not per launch
            y2 = e.render(x[:11])
        for s in range(K):
            truth = O.engine_truth(x[s], [hs[s], hs[s]], [dict(wet=1.0)] * 2)
            for o in range(2):
                assert O.rel_l2(y[s, o], truth[o]) < 5e-6, (flags, s, o, O.rel_l2(y[s, o], truth[o]))
        xx = np.concatenate([x[:11], x[:11]], axis=-1)
        for s in (0, 5, 10):
            truth = O.engine_truth(xx[s], [hs[s], hs[s]], [dict(wet=1.0)] * 2)
            assert O.rel_l2(y2[s, 0], truth[0][n:]) < 5e-6, (flags, s)


def test_fused_tier0_kernel_matches_three_kernel_path(monkeypatch):
    """k_fused0 (forward || TMA streaming -> MAC -> inverse in one CTA per instance, default for small
    batches) against the k_forward / k_mac / k_inverse chain, including a cross-fade and predelay."""
    m = ca()
    B, L, K = 128, 128 * 8 + 1024 * 5, 3
    tiers = [(128, 8), (1024, 0)]
    irs = [irs2x2(L, 4000 + 8 * s) for s in range(K)]
    n = B * 300
    x = np.stack([np.stack([O.synth_audio(n, 5000 + 2 * s + i) for i in range(2)]) for s in range(K)])

    def go(fuse):
        monkeypatch.setenv("CA_FUSE", fuse)
        with m.Engine(period=B, max_ir_frames=L, n_instances=K, n_ir_slots=2 * K, tiers=tiers, max_voices=2) as e:
            assert e.stats().tier0_fused == int(fuse)
            for s in range(K):
                for i in range(2):
                    e.load_ir(2 * s + i, irs[s][i][0], irs[s][i][1])
                    e.set_params(s, i, select=2 * s + i, wet=0.9, dry=0.2, predelay=33 * s, panWet=0.1 * s)
                    e.set_glide(s, i, 0.9)
            out = np.empty((K, 2, n), np.float32)
            for t in range(n // B):
                if t == 120:
                    e.set_params(2, 1, select=4, wet=0.9, dry=0.2, predelay=66, panWet=0.2, vsteps=20)
                out[:, :, t * B:(t + 1) * B] = e.process(x[:, :, t * B:(t + 1) * B])
            return out

    yf, yu = go("1"), go("0")
    assert O.rel_l2(yf, yu) < 1e-6
    truth = O.engine_truth(x[1], irs[1], [dict(wet=0.9, dry=0.2, panWet=0.1)] * 2, predelay=33)
    assert O.rel_l2(yf[1, 0], truth[0]) < 5e-6


@pytest.mark.parametrize("slots", ["1", "3"])
def test_persistent_mac_schedule_matches_one_cta_per_item(monkeypatch, slots):
    """k_mac_p (batches: a CTA walks several (instance, bin tile) work items and its TMA ring stays
    full across them) against k_mac (one CTA per item): same partial sums in the same order, so the
    output must agree to the last bit -- through a cross-fade (row lists of different lengths per
    instance, a second voice), staggered long tiers, and instances that have nothing to read yet."""
    m = ca()
    B, K = 64, 11
    tiers = [(64, 8), (512, 3), (2048, 0)]
    L = 64 * 8 + 512 * 3 + 2048 * 2 - 100
    irs = [irs2x2(L, 9000 + 8 * s) for s in range(K)]
    n = B * 200
    x = np.stack([np.stack([O.synth_audio(n, 9500 + 2 * s + i) for i in range(2)]) for s in range(K)])

    def go(persist, tr):
        monkeypatch.setenv("CA_MAC_PERSIST", persist)
        monkeypatch.setenv("CA_MAC_SLOTS", slots)
        monkeypatch.setenv("CA_FUSE", "0")
        with m.Engine(period=B, max_ir_frames=L, n_instances=K, n_ir_slots=2 * K, tiers=tr, max_voices=2, mac_split=1) as e:
            for s in range(K):
                for i in range(2):
                    e.load_ir(2 * s + i, irs[s][i][0], irs[s][i][1])
                    e.set_params(s, i, select=2 * s + i, wet=0.8, dry=0.1, predelay=7 * s, panWet=0.05 * s - 0.2)
                    e.set_glide(s, i, 0.8)
            out = np.empty((K, 2, n), np.float32)
            for t in range(n // B):
                if t == 90:
                    e.set_params(4, 0, select=3, wet=0.8, dry=0.1, predelay=28, panWet=0.0, vsteps=30)
                out[:, :, t * B:(t + 1) * B] = e.process(x[:, :, t * B:(t + 1) * B])
            return out

    for tr in (tiers, None):
        yp, y1 = go("1", tr), go("0", tr)
        assert np.array_equal(yp, y1), (tr, O.rel_l2(yp, y1))
        truth = O.engine_truth(x[2], irs[2], [dict(wet=0.8, dry=0.1, panWet=-0.1)] * 2, predelay=14)
        assert O.rel_l2(yp[2, 1], truth[1]) < 5e-6


def test_pipelined_batch_schedule_matches_sequential(monkeypatch):
    """run_pipelined (batches: FFT kernels and MAC kernels on two overlapping lanes, the long tiers'
    inverse transforms one period late, partial sums double-buffered) against the sequential schedule:
    same kernels on the same data, so bit-identical -- host path, device path, a cross-fade, a change
    of the active count mid-run and an IR load while the pipeline is in flight."""
    import torch
    m = ca()
    B, K = 64, 11
    tiers = [(64, 8), (512, 3), (2048, 0)]
    L = 64 * 8 + 512 * 3 + 2048 * 2 - 100
    irs = [irs2x2(L, 9100 + 8 * s) for s in range(K)]
    n = B * 150
    x = np.stack([np.stack([O.synth_audio(n, 9700 + 2 * s + i) for i in range(2)]) for s in range(K)])

    def go(pipe, device_path):
        monkeypatch.setenv("CA_PIPELINE", pipe)
        monkeypatch.setenv("CA_FUSE", "0")
        with m.Engine(period=B, max_ir_frames=L, n_instances=K, n_ir_slots=2 * K + 1, tiers=tiers, max_voices=2) as e:
            for s in range(K):
                for i in range(2):
                    e.load_ir(2 * s + i, irs[s][i][0], irs[s][i][1])
                    e.set_params(s, i, select=2 * s + i, wet=0.8, dry=0.1, predelay=5 * s, panWet=0.05 * s - 0.2)
                    e.set_glide(s, i, 0.8)
            out = np.zeros((K, 2, n), np.float32)
            xd = torch.from_numpy(x).cuda()
            yd = torch.empty(K, 2, B, device="cuda")
            for t in range(n // B):
                if t == 60:
                    e.set_params(4, 0, select=3, wet=0.8, dry=0.1, predelay=20, panWet=0.0, vsteps=30)
                if t == 70:
                    e.load_ir(2 * K, irs[0][0][0], irs[0][0][1])
                if t == 100:
                    e.set_active(7)
                k = e.n_active
                if device_path:
                    xb = xd[:k, :, t * B:(t + 1) * B].contiguous()
                    torch.cuda.synchronize()
                    e.process_device(xb.data_ptr(), yd.data_ptr())
                    e.sync()
                    out[:k, :, t * B:(t + 1) * B] = yd[:k].cpu().numpy()
                else:
                    out[:k, :, t * B:(t + 1) * B] = e.process(x[:k, :, t * B:(t + 1) * B])
            return out

    ref = go("0", False)
    for device_path in (False, True):
        y = go("1", device_path)
        assert np.array_equal(y, ref), (device_path, O.rel_l2(y, ref))
    truth = O.engine_truth(x[2], irs[2], [dict(wet=0.8, dry=0.1, panWet=-0.1)] * 2, predelay=10)
    assert O.rel_l2(ref[2, 1], truth[1]) < 5e-6


def test_pipelined_chunked_host_path_large_batch(monkeypatch):
    """600 instances (two host-pipeline chunks, every tier-residue class populated): pipelined against
    sequential through ca_process, and instance 0 / 599 against the fp64 oracle."""
    m = ca()
    B, K = 32, 600
    tiers = [(32, 8), (256, 3), (1024, 0)]
    L = 32 * 8 + 256 * 3 + 1024 * 2 - 30
    bank = [[O.synth_ir(L, FS, 9900 + 2 * b + o) for o in range(2)] for b in range(4)]
    n = B * 100
    rng = np.random.default_rng(5)
    x = (rng.standard_normal((K, 2, n)) * 0.1).astype(np.float32)

    def go(pipe):
        monkeypatch.setenv("CA_PIPELINE", pipe)
        with m.Engine(period=B, max_ir_frames=L, n_instances=K, n_ir_slots=4, tiers=tiers, flags=m.FLAG_STREAMING) as e:
            for b in range(4):
                e.load_ir(b, bank[b][0], bank[b][1])
            for s in range(K):
                for i in range(2):
                    e.set_params(s, i, select=(s + i) % 4, wet=1.0, dry=0.0)
                    e.set_glide(s, i, 1.0)
            return e.render(x)

    y1, y0 = go("1"), go("0")
    assert np.array_equal(y1, y0), O.rel_l2(y1, y0)
    for s in (0, 599):
        truth = O.engine_truth(x[s], [bank[s % 4], bank[(s + 1) % 4]], [dict(wet=1.0)] * 2)
        assert O.rel_l2(y1[s, 0], truth[0]) < 5e-6


@pytest.mark.parametrize("graph", [False, True])
def test_async_tiers_match_fp64_and_never_race(graph):
    """CA_FLAG_ASYNC_TIERS: every long tier starts one period later in the IR, its work runs on a
    low-priority stream beside the next period (result due two periods after its block closes).
    Checked against the fp64 oracle (the partition plan differs from the synchronous one, so not
    bitwise against it) for one and several instances, with and without a CUDA graph, through a
    cross-fade, a predelay, an IR load and a change of the active count while tiers are in flight;
    two identical runs must agree to the last bit (no race between the tier stream and the periods)."""
    m = ca()
    B, K = 64, 5
    L = 64 * 9 + 512 * 4 + 2048 * 3 - 77
    irs = [irs2x2(L, 9300 + 8 * s) for s in range(K)]
    n = B * 220
    x = np.stack([np.stack([O.synth_audio(n, 9800 + 2 * s + i) for i in range(2)]) for s in range(K)])
    flags = m.FLAG_ASYNC_TIERS | (m.FLAG_GRAPH if graph else 0)

    def go():
        with m.Engine(period=B, max_ir_frames=L, n_instances=K, n_ir_slots=2 * K + 1, tiers="auto", tier_growth=8, tier_max_block=2048,
                      max_voices=2, flags=flags) as e:
            st = e.stats()
            blocks, offs = list(st.tier_block[:st.n_tiers]), list(st.tier_offset[:st.n_tiers])
            assert len(blocks) == 3 and all(offs[j] >= blocks[j] + B for j in range(1, 3)), (blocks, offs)
            for s in range(K):
                for i in range(2):
                    e.load_ir(2 * s + i, irs[s][i][0], irs[s][i][1])
                    e.set_params(s, i, select=2 * s + i, wet=0.8, dry=0.1, predelay=9 * s, panWet=0.1 * s - 0.2)
                    e.set_glide(s, i, 0.8)
            out = np.zeros((K, 2, n), np.float32)
            for t in range(n // B):
                if t == 70:
                    e.set_params(3, 1, select=2, wet=0.8, dry=0.1, predelay=27, panWet=0.1, vsteps=25)
                if t == 90:
                    e.load_ir(2 * K, irs[0][0][0], irs[0][0][1])
                if t == 150:
                    e.set_active(3)
                k = e.n_active
                out[:k, :, t * B:(t + 1) * B] = e.process(x[:k, :, t * B:(t + 1) * B])
            return out

    y = go()
    assert np.array_equal(y, go())
    for s in (0, 2, 4):
        truth = O.engine_truth(x[s], irs[s], [dict(wet=0.8, dry=0.1, panWet=0.1 * s - 0.2)] * 2, predelay=9 * s)
        stop = n if s < 3 else B * 150
        for o in range(2):
            assert O.rel_l2(y[s, o, :stop], truth[o][:stop]) < 5e-6, (graph, s, o, O.rel_l2(y[s, o, :stop], truth[o][:stop]))


def test_async_tiers_reject_plans_without_slack():
    m = ca()
    with pytest.raises(m.CaError) as ei:
        m.Engine(period=64, max_ir_frames=64 * 8 + 512 * 4, tiers=[(64, 8), (512, 0)], flags=m.FLAG_ASYNC_TIERS)
    assert ei.value.code == -1
    with m.Engine(period=64, max_ir_frames=64 * 9 + 512 * 4, tiers=[(64, 9), (512, 0)], flags=m.FLAG_ASYNC_TIERS) as e:
        assert e.stats().n_tiers == 2


def test_async_tiers_block_4096_max_predelay_ring_has_slack():
    """Largest tier block 4096 with the maximal predelay: the time ring must be long enough that the
    forward of period t_end (clear-ahead at +8192, scatter up to +8191 + B) never aliases the window
    [t_end*B - 2*4096, t_end*B) a tier forward still reads on the asynchronous stream."""
    m = ca()
    B = 64
    L = 64 * 9 + 512 * 8 + 4096 * 3 - 5
    irs = irs2x2(L, 9700)
    n = B * 600
    x = np.stack([O.synth_audio(n, 9900 + i) for i in range(2)])
    pr = [dict(wet=0.9, dry=0.1), dict(wet=0.7, dry=0.2)]

    def go():
        with m.Engine(period=B, max_ir_frames=L, tiers="auto", tier_growth=8, tier_max_block=4096, flags=m.FLAG_ASYNC_TIERS) as e:
            st = e.stats()
            assert list(st.tier_block[:st.n_tiers]) == [64, 512, 4096]
            for i in range(2):
                e.load_ir(i, irs[i][0], irs[i][1])
                e.set_params(0, i, select=i, predelay=8191, **pr[i])
                e.set_glide(0, i, pr[i]["wet"])
            return e.render(x[None])[0]

    y = go()
    for _ in range(3):
        assert np.array_equal(y, go())
    truth = O.engine_truth(x, irs, pr, predelay=8191)
    for o in range(2):
        assert O.rel_l2(y[o], truth[o]) < 5e-6, (o, O.rel_l2(y[o], truth[o]))


def test_reactivated_instances_start_clean():
    """ca_set_active(n) followed by a later increase: the instances that were parked restart like new
    ones (no replay of pre-deactivation audio as a tail)."""
    m = ca()
    B, L, K = 64, 64 * 8 + 512 * 5 - 3, 3
    irs = irs2x2(L, 9750)
    n1, n2, n3 = 40, 30, 120
    x = np.stack([np.stack([O.synth_audio(B * (n1 + n2 + n3), 9950 + 2 * s + i) for i in range(2)]) for s in range(K)])
    pr = [dict(wet=1.0, dry=0.0)] * 2
    with m.Engine(period=B, max_ir_frames=L, n_instances=K, tiers=[(64, 8), (512, 0)]) as e:
        for i in range(2):
            e.load_ir(i, irs[i][0], irs[i][1])
        for s in range(K):
            for i in range(2):
                e.set_params(s, i, select=i, **pr[i])
                e.set_glide(s, i, 1.0)
        out = np.zeros((K, 2, B * (n1 + n2 + n3)), np.float32)
        for t in range(n1 + n2 + n3):
            if t == n1:
                e.set_active(1)
            if t == n1 + n2:
                e.set_active(K)
                for s in range(1, K):
                    for i in range(2):
                        e.set_glide(s, i, 1.0)      # skip the fade-in of the restarted voices
            k = e.n_active
            out[:k, :, t * B:(t + 1) * B] = e.process(x[:k, :, t * B:(t + 1) * B])
    t0 = B * (n1 + n2)
    truth0 = O.engine_truth(x[0], irs, pr)
    assert O.rel_l2(out[0, 0], truth0[0]) < 5e-6               # instance 0 never stopped
    for s in range(1, K):
        truth = O.engine_truth(x[s][:, t0:], irs, pr)          # restarted: history before t0 is gone
        for o in range(2):
            assert O.rel_l2(out[s, o, t0:], truth[o]) < 5e-6, (s, o, O.rel_l2(out[s, o, t0:], truth[o]))


@pytest.mark.parametrize("B,tiers", [
    (256, [(256, 4), (1024, 4), (4096, 4), (16384, 0)]),    # row-FFT family: M1 = 4, 16 (one CTA), 64 (columns + rows)
    (256, [(256, 2), (512, 3), (2048, 3), (8192, 0)]),      # M1 = 2, 8, 32
    (32, [(32, 8), (256, 8), (2048, 0)]),                   # M1 = 1, 8 under a warp-FFT tier 0
])
def test_row_fft_family_matches_legacy_and_fp64(B, tiers):
    """Every transform size of the row-FFT kernels (kernels_rows.cuh) against the round-1 FFT kernels on the
    same input (CA_FLAG_LEGACY_FFT: same layouts, so only fp32 rounding differs) and against fp64, three
    phase-staggered instances, predelay, pans, a mid-run IR switch (two voices active)."""
    m = ca()
    K = 3
    L = sum(b * p for b, p in tiers[:-1]) + 3 * tiers[-1][0] - 11
    n = ((L + 5 * tiers[-1][0]) // B) * B
    irs = [irs2x2(L, 7100 + 8 * s) for s in range(K)]
    x = np.stack([np.stack([O.synth_audio(n, 7200 + 2 * s + i) for i in range(2)]) for s in range(K)])
    pr = [dict(wet=0.9, dry=0.2, level=0.8, panWet=0.2, panDry=-0.3), dict(wet=0.7, dry=0.1, level=1.0, panWet=-0.4, panDry=0.1)]
    t_switch = (n // B) // 2

    def go(flags, schedule=0):
        with m.Engine(period=B, max_ir_frames=L, n_instances=K, n_ir_slots=2 * K, tiers=tiers, flags=flags, schedule=schedule) as e:
            for s in range(K):
                for i in range(2):
                    e.load_ir(2 * s + i, irs[s][i][0], irs[s][i][1])
                    e.set_params(s, i, select=2 * s + i, predelay=5 + 40 * s, **pr[i])
                    e.set_glide(s, i, pr[i]["wet"])
            out = np.zeros((K, 2, n), np.float32)
            for t in range(n // B):
                if t == t_switch:
                    e.set_params(2, 0, select=1, predelay=85, vsteps=20, **pr[0])   # instance 2 cross-fades input 0 to another IR
                out[:, :, t * B:(t + 1) * B] = e.process(x[:, :, t * B:(t + 1) * B])
            return out

    y_new, y_old, y_r8 = go(0), go(m.FLAG_LEGACY_FFT), go(0, m.SCHED_ROWS8)   # 16 x 16 rows (default), round-1 kernels, 8 x 8 x 4 rows
    for s in range(K):
        for o in range(2):
            assert O.rel_l2(y_new[s, o], y_old[s, o]) < 2e-6, (s, o, O.rel_l2(y_new[s, o], y_old[s, o]))
            assert O.rel_l2(y_r8[s, o], y_old[s, o]) < 2e-6, (s, o, O.rel_l2(y_r8[s, o], y_old[s, o]))
    for s in range(2):
        truth = O.engine_truth(x[s], irs[s], pr, predelay=5 + 40 * s)
        for o in range(2):
            assert O.rel_l2(y_new[s, o], truth[o]) < 5e-6, (s, o, O.rel_l2(y_new[s, o], truth[o]))


def test_row_fft_tier0_uniform_matches_legacy():
    """Tier 0 at B = 256 on k_fwd0_rows / k_inv0_rows (uniform partitioning, batch of 5, predelay in and
    out of the current block, clamp active) against the warp-shuffle kernels."""
    m = ca()
    B, L, K = 256, 256 * 9 - 3, 5
    irs = irs2x2(L, 7300)
    n = B * 60
    x = np.stack([np.stack([O.synth_audio(n, 7400 + 2 * s + i, rms=0.5) for i in range(2)]) for s in range(K)])

    def go(flags, schedule=0):
        with m.Engine(period=B, max_ir_frames=L, n_instances=K, flags=flags, schedule=schedule) as e:
            for i in range(2):
                e.load_ir(i, 2.0 * irs[i][0], 2.0 * irs[i][1])
            for s in range(K):
                for i in range(2):
                    e.set_params(s, i, select=i, predelay=(0, 7, 255, 256, 8191)[s], wet=0.9, dry=0.25, panWet=0.1 * s, panDry=-0.1 * s)
                    e.set_glide(s, i, 0.9)
            return e.render(x)

    y_new, y_old, y_r8 = go(0), go(m.FLAG_LEGACY_FFT), go(0, m.SCHED_ROWS8)
    assert (np.abs(y_new) > 0.999).sum() > 50          # clamp fires
    for s in range(K):
        for o in range(2):
            assert O.rel_l2(y_new[s, o], y_old[s, o]) < 2e-6, (s, o, O.rel_l2(y_new[s, o], y_old[s, o]))
            assert O.rel_l2(y_r8[s, o], y_old[s, o]) < 2e-6, (s, o, O.rel_l2(y_r8[s, o], y_old[s, o]))


def test_shared_voice_pool_crossfades_and_exhaustion():
    """Cross-fade voices come from a pool shared by all inputs of the engine (ca_config.voice_pool).  40 instances
    switch IR at once with only 6 shared voices: six inputs glide, the others fall back to a hard switch; the
    instances that do not switch are untouched, everybody ends up on the new IR exactly, the pool is handed
    back (a second wave of switches finds voices again), and with a big enough pool the batch equals
    per-instance engines that have every voice resident."""
    m = ca()
    B, K = 64, 40
    L = 64 * 8 + 512 * 3 - 7
    tiers = [(64, 8), (512, 0)]
    irs = irs2x2(L, 7700) + irs2x2(L, 7750)          # bank slots 0..3
    nper = 260
    x = np.stack([np.stack([O.synth_audio(B * nper, 7800 + 2 * s + i) for i in range(2)]) for s in range(K)])
    switchers = [s for s in range(K) if s % 4 != 3]   # 30 instances switch input 0 from IR 0 to IR 2

    def go(pool, k_inst, xs, which):
        with m.Engine(period=B, max_ir_frames=L, n_instances=k_inst, n_ir_slots=4, tiers=tiers, voice_pool=pool) as e:
            for j in range(4):
                e.load_ir(j, irs[j][0], irs[j][1])
            for s in range(k_inst):
                for i in range(2):
                    e.set_params(s, i, select=i, wet=1.0, dry=0.0)
                    e.set_glide(s, i, 1.0)
            out = np.zeros((k_inst, 2, B * nper), np.float32)
            for t in range(nper):
                if t == 60:
                    for s in which:
                        e.set_params(s, 0, select=2, wet=1.0, dry=0.0, vsteps=10)
                if t == 160:                           # second wave: back to IR 0 (needs pool entries again)
                    for s in which:
                        e.set_params(s, 0, select=0, wet=1.0, dry=0.0, vsteps=10)
                out[:, :, t * B:(t + 1) * B] = e.process(xs[:, :, t * B:(t + 1) * B])
            return out

    y_small = go(6, K, x, switchers)
    y_big = go(2 * K, K, x, switchers)
    # (1) instances that never switch: identical whatever the others do, and equal to the truth
    for s in (3, 39):
        assert np.array_equal(y_small[s], y_big[s])
        truth = O.engine_truth(x[s], [irs[0], irs[1]], [dict(wet=1.0)] * 2)
        assert O.rel_l2(y_small[s, 0], truth[0]) < 5e-6
    # (2) with every voice available the batch equals single-instance engines (all voices resident there)
    for s in (0, 17):
        one = go(0, 1, x[s:s + 1], [0])
        assert O.rel_l2(y_big[s, 0], one[0, 0]) < 1e-6, (s, O.rel_l2(y_big[s, 0], one[0, 0]))
    # (3) exhausted pool: once the old IR has rung out (IR length + glide) everybody is exactly on the new IR
    settle = 60 + 10 + 60 + (L + B - 1) // B + 4
    for s in switchers:
        seg = slice(settle * B, 160 * B)   # only input fed after the switch is still audible here
        assert O.rel_l2(y_small[s, 0, seg], y_big[s, 0, seg]) < 1e-5, s
    # (4) the second wave found pool entries again for some inputs: at least six of them glide exactly like the big pool
    same = sum(np.array_equal(y_small[s, 0, 160 * B:200 * B], y_big[s, 0, 160 * B:200 * B]) for s in switchers)
    assert same >= 6, same


def test_schedule_bits_in_the_config_select_the_same_kernels_as_the_env_knobs():
    """ca_config.schedule (CA_SCHED_*) instead of environment variables: persistent vs per-item MAC, fused vs
    three-kernel tier 0, pipelined lanes vs sequential -- all bit-identical on the same input."""
    m = ca()
    B, K = 64, 6
    tiers = [(64, 8), (512, 3), (2048, 0)]
    L = 64 * 8 + 512 * 3 + 2048 * 2 - 33
    irs = [irs2x2(L, 9100 + 8 * s) for s in range(K)]
    n = B * 120
    x = np.stack([np.stack([O.synth_audio(n, 9600 + 2 * s + i) for i in range(2)]) for s in range(K)])

    def go(schedule):
        with m.Engine(period=B, max_ir_frames=L, n_instances=K, n_ir_slots=2 * K, tiers=tiers, mac_split=1, schedule=schedule) as e:
            for s in range(K):
                for i in range(2):
                    e.load_ir(2 * s + i, irs[s][i][0], irs[s][i][1])
                    e.set_params(s, i, select=2 * s + i, wet=0.8, dry=0.1, predelay=3 * s)
                    e.set_glide(s, i, 0.8)
            st = e.stats()
            return e.render(x), st.tier0_fused

    y_item, f0 = go(m.SCHED_MAC_PER_ITEM | m.SCHED_NO_FUSED_TIER0)
    y_pers, f1 = go(m.SCHED_MAC_PERSISTENT | m.SCHED_NO_FUSED_TIER0)
    y_pipe, _ = go(m.SCHED_MAC_PERSISTENT | m.SCHED_NO_FUSED_TIER0 | m.SCHED_PIPELINED)
    y_fused, f2 = go(m.SCHED_FUSED_TIER0)
    assert (f0, f1, f2) == (0, 0, 1)
    assert np.array_equal(y_item, y_pers) and np.array_equal(y_pers, y_pipe)
    assert O.rel_l2(y_fused, y_item) < 2e-6


def test_sm_split_green_contexts_bit_identical():
    """ca_config.sm_split: MAC lane and FFT lanes of the pipelined batch schedule on disjoint SM sets (CUDA green
    contexts).  Only the placement changes: output identical to the sequential schedule bit for bit."""
    m = ca()
    B, K = 64, 24
    tiers = [(64, 8), (512, 3), (2048, 0)]
    L = 64 * 8 + 512 * 3 + 2048 * 2 - 9
    irs = [irs2x2(L, 9200 + 8 * s) for s in range(4)]
    n = B * 150
    x = np.stack([np.stack([O.synth_audio(n, 9700 + 2 * s + i) for i in range(2)]) for s in range(K)])

    def go(**kw):
        with m.Engine(period=B, max_ir_frames=L, n_instances=K, n_ir_slots=8, tiers=tiers, mac_split=1,
                      schedule=m.SCHED_MAC_PERSISTENT | m.SCHED_NO_FUSED_TIER0, **kw) as e:
            for s in range(4):
                for i in range(2):
                    e.load_ir(2 * s + i, irs[s][i][0], irs[s][i][1])
            for s in range(K):
                for i in range(2):
                    e.set_params(s, i, select=2 * (s % 4) + i, wet=0.8, dry=0.1, predelay=2 * s)
                    e.set_glide(s, i, 0.8)
            return e.render(x)

    try:
        y_split = go(sm_split=96)
    except m.CaError as ex:
        if ex.code == -5:
            pytest.skip("green contexts not available: " + str(ex))
        raise
    assert np.array_equal(y_split, go())


def test_rows16_mono_odd_batch_idle_half_warps():
    """kernels_rows16.cuh maps one HALF-warp to an (instance, input | output) item: a mono batch of 3 leaves the last
    warp with an idle half, and neighbouring halves carry different predelays (one zero, one not).  Tiered, against
    fp64 and against the 8 x 8 x 4 row kernels."""
    m = ca()
    B, K = 256, 3
    tiers = [(256, 4), (1024, 4), (4096, 0)]
    L = 256 * 4 + 1024 * 4 + 4096 * 3 - 5
    n = B * 120
    hs = [O.synth_ir(L, 48000, 7500 + s) for s in range(K)]
    x = np.stack([O.synth_audio(n, 7600 + s)[None, :] for s in range(K)])
    pds = (0, 300, 8191)

    def go(schedule):
        with m.Engine(period=B, max_ir_frames=L, n_instances=K, n_in=1, n_out=1, n_ir_slots=K, tiers=tiers, schedule=schedule) as e:
            for s in range(K):
                e.load_ir(s, hs[s])
                e.set_params(s, 0, select=s, predelay=pds[s], wet=0.8, dry=0.1)
                e.set_glide(s, 0, 0.8)
            return e.render(x)

    y, y8 = go(0), go(m.SCHED_ROWS8)
    for s in range(K):
        truth = O.engine_truth(x[s], [[hs[s]]], [dict(wet=0.8, dry=0.1)], predelay=pds[s])
        assert O.rel_l2(y[s, 0], truth[0]) < 5e-6, (s, O.rel_l2(y[s, 0], truth[0]))
        assert O.rel_l2(y[s, 0], y8[s, 0]) < 2e-6, (s, O.rel_l2(y[s, 0], y8[s, 0]))
