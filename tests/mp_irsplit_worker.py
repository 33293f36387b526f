"""Worker for the multi-GPU IR-split test / demo: launched with
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/mp_irsplit_worker.py
Every rank convolves the same input with its partition range of one long IR; the partial output
blocks are summed with an NCCL reduce; rank 0 checks the result against the fp64 oracle and against
the single-GPU engine."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cuda-audio_b200", "python"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import cuda_audio_b200 as ca  # noqa: E402
from cuda_audio_b200 import shard  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    fs, B, L, nper = 48000, 256, 256 * 101 + 77, 260
    cfg = shard.IrSplitConfig(period=B, ir_frames=L)
    grp = shard.IrSplitGroup(cfg, lambda pb, pc: shard.TorchEngine(cfg, local, pb, pc), dist=dist, rank=rank, world=world)
    irs = [[3.0 * O.synth_ir(L, fs, 90 + 2 * i + o) for o in range(2)] for i in range(2)]
    pr = [dict(select=0, wet=1.0, dry=0.3, level=0.9, panWet=0.2, panDry=-0.4), dict(select=1, wet=0.8, dry=0.2, level=1.0, panWet=-0.3, panDry=0.5)]
    for i in range(2):
        grp.load_ir(i, irs[i][0], irs[i][1])
        grp.set_params(i, **pr[i])
        grp.set_glide(i, pr[i]["wet"])
    x = np.stack([O.synth_audio(B * nper, 95 + i, rms=0.3) for i in range(2)])
    xd = torch.from_numpy(x).to(dev)
    ys = []
    for t in range(nper):
        y = grp.process(xd[None, :, t * B:(t + 1) * B].contiguous(), zeros_like=lambda a: torch.zeros(1, 2, B, device=dev))
        if rank == 0:
            ys.append(y[0].cpu().numpy().copy())
    res = {}
    if rank == 0:
        got = np.concatenate(ys, axis=-1)
        truth = O.engine_truth(x, irs, pr)
        res["err_fp64"] = max(O.rel_l2(got[o], truth[o]) for o in range(2))
        with ca.Engine(period=B, max_ir_frames=L, device=local) as e:
            for i in range(2):
                e.load_ir(i, irs[i][0], irs[i][1])
                e.set_params(0, i, **pr[i])
                e.set_glide(0, i, pr[i]["wet"])
            one = e.render(x[None])[0]
        res["err_single_gpu"] = max(O.rel_l2(got[o], one[o]) for o in range(2))
        res["clipped"] = int((np.abs(truth - 0) > 1.0).sum())
        res["world"] = world
        res["plan"] = grp.plan
        print("IRSPLIT_RESULT " + json.dumps(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
