"""Multi-GPU sharding of the convolution path, one process per GPU.

Two modes, both plain `torch.distributed` plumbing around the C ABI:

* **instances** -- independent instances (streams / channels) are dealt to the ranks in contiguous
  blocks; no data-path collective at all (BASELINE configs[3]).
* **ir_split**  -- ONE very long IR is cut by partition range; every rank convolves the same input
  with its range (`ca_config.part_begin / part_count`: the FDL of rank r is read with a delay of
  part_begin partitions) and the `n_out x B` partial output blocks are summed onto rank 0 with a
  reduce (NCCL over NVLink on GPUs, gloo in the CPU tests) once per period (BASELINE configs[4]).
  Every shard outputs its raw (unclamped, wet-only) block (CA_FLAG_RAW_WET); rank 0 applies the
  reference's clamp(+-1) to the SUM and adds the dry mix once (conv.cu:98, 418-427).

The compute backend is injected (`make_engine`) so the host logic -- partition planning, parameter
fan-out, reduce, dry-mix ownership -- is testable on CPU with world_size 2.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Tuple


def split_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [begin, begin + count) of n items for `rank`; sizes differ by at most 1."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, base + (1 if rank < rem else 0)


def plan_instances(n_instances: int, world: int) -> List[Tuple[int, int]]:
    """instance sharding: [(first_instance, count)] per rank."""
    return [split_range(n_instances, world, r) for r in range(world)]


def plan_partitions(ir_frames: int, period: int, world: int) -> List[Tuple[int, int]]:
    """IR partition-range sharding: [(part_begin, part_count)] per rank over P = ceil(L / B) partitions.
    Ranks that would get nothing (world > P) get count 0 and must sit the reduce out with zeros."""
    P = (ir_frames + period - 1) // period
    return [split_range(P, world, r) for r in range(world)]


@dataclass
class IrSplitConfig:
    period: int
    ir_frames: int
    n_in: int = 2
    n_out: int = 2
    n_ir_slots: int = 2
    sample_rate: float = 48000.0
    flags: int = 0


class IrSplitGroup:
    """One rank's share of an IR-split convolution.  `make_engine(part_begin, part_count)` must return
    an object with load_ir(slot, left, right), set_params(instance, input, **kw), set_glide(instance,
    input, g) and process_tensor(x) -> tensor [1][n_out][B] on this rank's device."""

    def __init__(self, cfg: IrSplitConfig, make_engine: Callable[[int, int], object], dist=None, rank: int = 0, world: int = 1):
        self.cfg, self.dist, self.rank, self.world = cfg, dist, rank, world
        self.plan = plan_partitions(cfg.ir_frames, cfg.period, world)
        self.part_begin, self.part_count = self.plan[rank]
        self.engine = make_engine(self.part_begin, self.part_count) if self.part_count > 0 else None
        self.params = {}

    def load_ir(self, slot, left, right=None):
        if self.engine is not None:
            self.engine.load_ir(slot, left, right)

    def set_params(self, inp, **kw):
        """Same parameters on every rank (the wet path is linear in the IR, so it shards exactly)."""
        self.params[inp] = dict(self.params.get(inp, {}), **kw)
        if self.engine is not None:
            self.engine.set_params(0, inp, **kw)

    def set_glide(self, inp, g):
        if self.engine is not None:
            self.engine.set_glide(0, inp, g)

    def dry_gains(self):
        """[n_out][n_in] gain of the dry mix: dry * panDry * level (conv.cu:418-427)."""
        g = []
        for o in range(self.cfg.n_out):
            row = []
            for i in range(self.cfg.n_in):
                p = self.params.get(i, {})
                pan = p.get("panDry", 0.0)
                pg = 1.0 if self.cfg.n_out == 1 else ((1 - pan if pan >= 0 else 1.0) if o == 0 else (1 + pan if pan <= 0 else 1.0))
                row.append(p.get("dry", 0.5) * pg * p.get("level", 1.0))
            g.append(row)
        return g

    def process(self, x, zeros_like: Optional[Callable] = None):
        """x: tensor [1][n_in][B] (the same block on every rank).  Rank 0 returns the finished output
        block clamp(sum of partial wet blocks) + dry mix; other ranks return their raw partial."""
        if self.engine is not None:
            y = self.engine.process_tensor(x)
        else:
            y = zeros_like(x)
        if self.world > 1:
            self.dist.reduce(y, dst=0, op=self.dist.ReduceOp.SUM)
        if self.rank == 0:
            y = y.clamp(-1.0, 1.0)
            g = self.dry_gains()
            for o in range(self.cfg.n_out):
                for i in range(self.cfg.n_in):
                    if g[o][i] != 0.0:
                        y[0, o] += g[o][i] * x[0, i]
        return y


class TorchEngine:
    """GPU backend for IrSplitGroup: one raw-wet engine over torch device tensors."""

    def __init__(self, cfg: IrSplitConfig, device_index: int, part_begin: int, part_count: int, flags: int = 0):
        import torch

        from . import FLAG_RAW_WET, Engine
        self.torch = torch
        self.e = Engine(period=cfg.period, max_ir_frames=cfg.ir_frames, n_in=cfg.n_in, n_out=cfg.n_out, n_ir_slots=cfg.n_ir_slots,
                        device=device_index, flags=flags | cfg.flags | FLAG_RAW_WET, part_begin=part_begin, part_count=part_count,
                        sample_rate=cfg.sample_rate)
        self.dev = torch.device("cuda", device_index)
        self.y = torch.empty(1, cfg.n_out, cfg.period, device=self.dev)

    def load_ir(self, slot, left, right=None):
        self.e.load_ir(slot, left, right)

    def set_params(self, instance, inp, **kw):
        self.e.set_params(instance, inp, **kw)

    def set_glide(self, instance, inp, g):
        self.e.set_glide(instance, inp, g)

    def process_tensor(self, x):
        # x was produced, and the previous block's y is still being consumed (reduce, clamp, dry mix),
        # on torch's streams; the engine runs on its own non-blocking stream: order the two explicitly.
        self.torch.cuda.current_stream(self.dev).synchronize()
        self.e.process_device(x.data_ptr(), self.y.data_ptr())
        self.e.sync()
        return self.y

    def close(self):
        self.e.close()
