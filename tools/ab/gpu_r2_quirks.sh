#!/bin/bash
timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -x -q -k "quirks" 2>&1 | tail -15
python - <<'PY'
# informational: golden case C (unconstrained IRs, clamp ACTIVE in the reference's accumulator) through the drop-in class with engine.ref_quirks
import os, sys, numpy as np
sys.path.insert(0, '.')
os.environ['CA_ENGINE_REF_QUIRKS'] = '1'
import importlib.util
spec = importlib.util.spec_from_file_location("t", "tests/test_dropin_gpu.py"); t = importlib.util.module_from_spec(spec); spec.loader.exec_module(t)
from oracle import oracle as O
for name in ("C", "A", "B", "D"):
    ml, mr, _, _ = t.drive(t.DROPIN, name)
    z = np.load(f"tests/golden/ref_{name}.npz")
    print(name, 'dropin+quirks vs golden', O.rel_l2(ml, z["L"]), O.rel_l2(mr, z["R"]), 'peak', float(np.abs(z["L"]).max()))
PY
