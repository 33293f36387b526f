"""Host-side sharding logic (cuda-audio_b200/python/cuda_audio_b200/shard.py) on CPU:
partition / instance plans, and the IR-split group (reduce over `gloo`, world_size 2) with a
numpy test double standing in for the per-rank CUDA engine."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cuda_audio_b200 import shard
from oracle import oracle as O

FS = 48000


def test_plans_cover_everything_exactly_once():
    for n, w in [(750, 1), (750, 2), (750, 8), (11250, 8), (5, 8), (1024, 3)]:
        plan = shard.plan_instances(n, w)
        assert plan[0][0] == 0 and sum(c for _, c in plan) == n
        for (b0, c0), (b1, _) in zip(plan, plan[1:]):
            assert b1 == b0 + c0
        assert max(c for _, c in plan) - min(c for _, c in plan) <= 1
    assert shard.plan_partitions(2880000, 256, 8)[0] == (0, 1407)           # configs[4]: 60 s IR, P = 11250
    assert sum(c for _, c in shard.plan_partitions(2880000, 256, 8)) == 11250
    with pytest.raises(ValueError):
        shard.split_range(10, 2, 2)


class NumpyShard:
    """Test double for one rank's engine: raw wet block of the IR partition range, by direct evaluation."""

    def __init__(self, cfg, part_begin, part_count):
        self.cfg, self.lo, self.hi = cfg, part_begin * cfg.period, (part_begin + part_count) * cfg.period
        self.irs, self.par, self.hist = {}, {}, [np.zeros(0) for _ in range(cfg.n_in)]

    def load_ir(self, slot, left, right=None):
        def seg(h):
            m = np.zeros(len(h))
            m[self.lo:self.hi] = np.asarray(h, np.float64)[self.lo:self.hi]   # other ranks' partitions are not ours
            return m
        self.irs[slot] = (seg(left), seg(right if right is not None else left))

    def set_params(self, instance, inp, **kw):
        self.par[inp] = dict(self.par.get(inp, {}), **kw)

    def set_glide(self, instance, inp, g):
        pass

    def process_tensor(self, x):
        B = self.cfg.period
        xn = x.numpy()[0].astype(np.float64)
        out = np.zeros((1, self.cfg.n_out, B), np.float32)
        for i in range(self.cfg.n_in):
            self.hist[i] = np.concatenate([self.hist[i], xn[i]])
        n = len(self.hist[0])
        for o in range(self.cfg.n_out):
            acc = np.zeros(B)
            for i in range(self.cfg.n_in):
                p = self.par[i]
                pan = O.pan_gains(p.get("panWet", 0.0))[o]
                y = O.fft_conv(self.hist[i], self.irs[p["select"]][o], n)[n - B:]
                acc += pan * p.get("wet", 0.5) * p.get("level", 1.0) * y
            out[0, o] = acc
        return torch.from_numpy(out)


def _worker(rank, world, port, L, B, nper, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = shard.IrSplitConfig(period=B, ir_frames=L)
    grp = shard.IrSplitGroup(cfg, lambda pb, pc: NumpyShard(cfg, pb, pc), dist=dist, rank=rank, world=world)
    irs = [[4.0 * O.synth_ir(L, FS, 60 + 2 * i + o) for o in range(2)] for i in range(2)]     # loud: the SUM clips
    pr = [dict(select=0, wet=1.0, dry=0.3, level=0.9, panWet=0.2, panDry=-0.4), dict(select=1, wet=0.8, dry=0.2, level=1.0, panWet=-0.3, panDry=0.5)]
    for i in range(2):
        grp.load_ir(i, irs[i][0], irs[i][1])
        grp.set_params(i, **pr[i])
    x = np.stack([O.synth_audio(B * nper, 70 + i, rms=0.4) for i in range(2)])
    ys = []
    for t in range(nper):
        y = grp.process(torch.from_numpy(x[None, :, t * B:(t + 1) * B].copy()), zeros_like=lambda a: torch.zeros(1, 2, B))
        ys.append(y.numpy()[0].copy())
    if rank == 0:
        got = np.concatenate(ys, axis=-1)
        truth = O.engine_truth(x, irs, pr)
        wet = O.fft_conv(x[0], irs[0][0]) * 0.8 * 0.9 + O.fft_conv(x[1], irs[1][0]) * 0.8 * 1.3
        ret["err"] = max(O.rel_l2(got[o], truth[o]) for o in range(2))
        ret["clipped"] = int((np.abs(wet) > 1.0).sum())
        ret["plan"] = grp.plan
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 3])
def test_ir_split_group_reduces_to_the_whole_convolution(world):
    L, B, nper = 64 * 9 + 5, 64, 40
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), L, B, nper, ret), nprocs=world, join=True)
    assert ret["err"] < 1e-6, ret["err"]          # clamp applied to the SUM, dry mixed once (rank 0)
    assert ret["clipped"] > 50                     # ... and the clamp really fired
    assert sum(c for _, c in ret["plan"]) == 10
