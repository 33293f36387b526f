#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_engine_gpu.py -m gpu -x -q -k "persistent_kernel" > gpurun_out/r2p_tests.log 2>&1; echo "persist tests rc=$?"; tail -30 gpurun_out/r2p_tests.log | cut -c1-400
for ST in 0 1; do
if [ $ST = 1 ]; then export CA_PERSIST_STAMPS=1; fi
timeout 300 python - <<'PY'
import sys, os, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "cuda-audio_b200", "python"))
import numpy as np, cuda_audio_b200 as ca
from oracle import oracle as O
B, L = 256, 192000
for name, fl in (("graph", ca.FLAG_GRAPH), ("persistent", ca.FLAG_PERSISTENT)):
    with ca.Engine(period=B, max_ir_frames=L, flags=fl) as e:
        for i in range(2):
            h = [O.synth_ir(L, 48000, 10 + 2 * i + o) for o in range(2)]
            e.load_ir(i, h[0], h[1]); e.set_params(0, i, select=i); e.set_glide(0, i, 0.5)
        a, b = ca.PinnedArray((1, 2, B)), ca.PinnedArray((1, 2, B)); a.array[...] = 0.05
        for _ in range(1000): e.process_raw(a.ptr, b.ptr)
        e.reset_stats()
        for _ in range(3000): e.process_raw(a.ptr, b.ptr)
        s = e.stats()
        print(name, "p50 %.1f p99 %.1f max %.1f us" % (s.p50_us, s.p99_us, s.max_us), "launches", s.gpu_launches, flush=True)
        if fl == ca.FLAG_PERSISTENT:
            acc = np.zeros(7)
            for _ in range(200):
                e.process_raw(a.ptr, b.ptr)
                st = e.persist_stamps()
                acc += np.diff(np.array(st, dtype=np.int64))
            print("phases ns (read input, forward, mac cta0, wait all, sum, inverse, fence+publish):", (acc / 200).round(0).tolist(), "device total", round(acc.sum() / 200))
PY
done
