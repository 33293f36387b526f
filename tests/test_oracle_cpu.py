"""CPU checks of the oracles themselves: the restatement of conv.cu equals exact convolution
under the parity protocol, the fp32 CPU port matches the fp64 truth, CC mapping table."""
import numpy as np
import pytest

from oracle import oracle as O

FS = 48000


def _irs(L, seed0=1000, safe=True):
    return [[O.synth_ir(L, FS, seed0 + 2 * i + o, parity_safe=safe) for o in range(2)] for i in range(2)]


def test_synth_ir_is_dc_and_nyquist_free():
    h = O.synth_ir(4000, FS, 3).astype(np.float64)
    alt = np.where(np.arange(4000) % 2 == 0, 1.0, -1.0)
    assert abs(h.sum()) < 1e-6 and abs((h * alt).sum()) < 1e-6
    assert abs((h ** 2).sum() - 1.0) < 1e-5


def test_direct_conv_matches_numpy():
    rng = np.random.default_rng(0)
    x, h = rng.standard_normal(3000), rng.standard_normal(700)
    assert np.allclose(O.direct_conv(x, h), np.convolve(x, h)[:3000], atol=1e-12)
    idx = np.array([0, 5, 699, 700, 2999])
    assert np.allclose(O.direct_conv_at(x, h, idx), np.convolve(x, h)[idx], atol=1e-12)
    assert np.allclose(O.fft_conv(x, h), np.convolve(x, h)[:3000], atol=1e-10)


def test_restatement_equals_convolution_under_protocol():
    N, B = 4096, 64
    irs = _irs(N - B)
    x = np.stack([np.concatenate([np.zeros(100 * B, np.float32), O.synth_audio(B * 100, 2000 + i)]) for i in range(2)])
    r = O.RefConv(N)
    for i in range(2):
        r.prepare(i, irs[i][0], irs[i][1], B)
        r.set_cc(i, select=i, wet=1.0, dry=0.0)
    L, R = r.render(x[0], x[1], B)
    truth = O.engine_truth(x, irs, [dict(wet=1.0)] * 2, conv=O.direct_conv)
    assert O.rel_l2(L, truth[0]) < 1e-6 and O.rel_l2(R, truth[1]) < 1e-6


def test_restatement_quirks_without_protocol():
    """DC mis-unpack + missing Nyquist bin (conv.cu:61, 47-73): white-noise IRs are NOT convolved
    exactly (~1e-2 at N = 4096)."""
    N, B = 4096, 64
    irs = _irs(N - B, safe=False)
    x = np.stack([np.concatenate([np.zeros(100 * B, np.float32), O.synth_audio(B * 100, 2000 + i)]) for i in range(2)])
    r = O.RefConv(N)
    for i in range(2):
        r.prepare(i, irs[i][0], irs[i][1], B)
        r.set_cc(i, select=i, wet=1.0, dry=0.0)
    L, _ = r.render(x[0], x[1], B)
    truth = O.engine_truth(x, irs, [dict(wet=1.0)] * 2)
    err = O.rel_l2(L[100 * B:], truth[0][100 * B:])
    assert 1e-3 < err < 1e-1, err


def test_restatement_first_block_carries_one_fifth_of_wet():
    """A freshly started reference fades the wet path in: after the first glide step the live
    IR spectrum is wet/5 of the target (conv.cu:27 with vsteps = 0), and the WHOLE response of
    block 0 carries that gain."""
    N, B = 1024, 64
    h = np.array([1, -1, -1, 1], np.float32) * 0.5  # sum 0 and alternating sum 0
    r = O.RefConv(N)
    r.prepare(0, h, h, B)
    r.set_cc(0, wet=1.0, dry=0.0)
    r.set_cc(1, wet=0.0, dry=0.0)
    x = np.zeros(B * 4, np.float32)
    x[:B] = O.synth_audio(B, 1)
    L, _ = r.render(x, np.zeros_like(x), B)
    y = O.direct_conv(x, h)
    assert np.allclose(L, 0.2 * y, atol=1e-7)


def test_cpu_port_matches_truth_with_parameters():
    B, L = 128, 128 * 7 + 3
    irs = _irs(L)
    x = np.stack([O.synth_audio(B * 50, 2000 + i) for i in range(2)])
    pr = [dict(wet=0.7, dry=0.4, level=0.8, panWet=0.3, panDry=-0.2), dict(wet=0.5, dry=0.5, level=1.0, panWet=-0.5, panDry=0.4)]
    u = O.Upols(B, L, 2, 2, 1, 2)
    for i in range(2):
        u.load_ir(i, irs[i][0], irs[i][1])
        u.set_param(0, i, select=i, predelay=45, glide=pr[i]["wet"], **pr[i])
    y = u.render(x[None])[0]
    truth = O.engine_truth(x, irs, pr, predelay=45)
    assert O.rel_l2(y[0], truth[0]) < 5e-6 and O.rel_l2(y[1], truth[1]) < 5e-6


def test_cpu_port_mono_and_batch():
    B, L = 64, 500
    hs = [O.synth_ir(L, FS, 10 + s) for s in range(3)]
    x = np.stack([O.synth_audio(B * 30, 20 + s) for s in range(3)])[:, None, :]
    u = O.Upols(B, L, 1, 1, 3, 3)
    for s in range(3):
        u.load_ir(s, hs[s])
        u.set_param(s, 0, wet=1.0, dry=0.0, select=s, glide=1.0)
    y = u.render(x)
    for s in range(3):
        assert O.rel_l2(y[s, 0], O.fft_conv(x[s, 0], hs[s])) < 5e-6


def test_handle_cc_mapping_table():
    """conv.cu:255-276"""
    cc = O.CC(select=0, predelay=0, speed=100, vsteps=0, dry=0.5, wet=0.5, panDry=0, panWet=0, level=1)
    O.handle_cc(cc, 0, 64, 10)
    assert cc.select == 5 and cc.vsteps == 100          # select = v * nIR / 128, vsteps = speed
    O.handle_cc(cc, 1, 127, 10)
    assert cc.predelay == 127 * 8192 // 128
    O.handle_cc(cc, 2, 32, 10)
    assert cc.dry == 0.25
    O.handle_cc(cc, 3, 96, 10)
    assert cc.wet == 0.75
    O.handle_cc(cc, 4, 0, 10)
    assert cc.panDry == -1.0
    O.handle_cc(cc, 5, 96, 10)
    assert cc.panWet == 0.5
    O.handle_cc(cc, 6, 127, 10)
    assert abs(cc.level - 127 / 128) < 1e-7
    O.handle_cc(cc, 7, 4, 10)
    assert cc.speed == 32 and cc.vsteps == 32            # vsteps clipped to the new speed
    assert O.pan_gains(0.25) == (0.75, 1.0) and O.pan_gains(-0.25) == (1.0, 0.75)


def test_ref_quirk_model_rank1_terms_explain_the_reference():
    """The model behind CA_FLAG_REF_QUIRKS (cuda-audio_b200/csrc/kernels_quirks.cuh): the pinned restatement of
    conv.cu on UNCONSTRAINED IRs equals the exact convolution plus, per input block t, a constant D and an
    alternating E (-1)^(s - pd) over accumulator samples [tB + pd, tB + N)."""
    fs, N, B, pd = 48000, 4096, 64, 301
    L = N - B - 700
    irs = [[O.synth_ir(L, fs, 50 + 2 * i + o, parity_safe=False) for o in range(2)] for i in range(2)]
    warm = 100
    x = np.stack([np.concatenate([np.zeros(warm * B, np.float32), O.synth_audio(B * 200, 2000 + i)]) for i in range(2)])
    pr = [dict(wet=0.9, level=0.8, panWet=0.3), dict(wet=0.7, panWet=-0.5)]
    cpu = O.RefConv(N)
    for i in range(2):
        cpu.prepare(i, irs[i][0], irs[i][1], B)
        cpu.set_cc(i, select=i, dry=0.0, predelay=pd if i == 0 else 0, **pr[i])
    cl, cr = cpu.render(x[0], x[1], B)
    truth = O.engine_truth(x, irs, pr, predelay=pd)
    n = x.shape[1]
    sg = (-1.0) ** np.arange(N)
    hs = [[(irs[i][o].astype(np.float64).sum(), (irs[i][o].astype(np.float64) * sg[:L]).sum()) for o in range(2)] for i in range(2)]
    s = [[O.pan_gains(pr[i].get("panWet", 0.0))[o] * pr[i].get("level", 1.0) / N for o in range(2)] for i in range(2)]
    A = [[pr[i]["wet"] * hs[i][o][0] for o in range(2)] for i in range(2)]     # converged glide: c = wet
    Ap = [[pr[i]["wet"] * hs[i][o][1] for o in range(2)] for i in range(2)]
    corr = np.zeros((2, n + N + B))
    for t in range(n // B):
        blk = x[:, t * B:(t + 1) * B].astype(np.float64)
        a, ap = blk.sum(axis=1), (blk * sg[:B]).sum(axis=1)
        D = [-(s[0][0] * a[1] * A[0][1] + s[1][0] * a[1] * A[1][0]), -(s[0][1] * a[0] * A[0][1] + s[1][1] * a[1] * A[1][1])]
        E = [-(s[0][o] * ap[0] * Ap[0][o] + s[1][o] * ap[1] * Ap[1][o]) for o in range(2)]
        lo, hi = t * B + pd, t * B + N
        for o in range(2):
            corr[o, lo:hi] += D[o] + E[o] * sg[:hi - lo]
    sl = slice(warm * B, n)
    for o, r in enumerate((cl, cr)):
        assert O.rel_l2(r[sl], truth[o][sl]) > 1e-3                       # the quirks matter on such IRs
        assert O.rel_l2(r[sl], truth[o][sl] + corr[o][sl]) < 1e-6         # and the rank-1 terms are all of it
