#!/usr/bin/env python
"""Debug aid for the GPU box: run the CUDA path and the oracle on the same columns, print per-output max errors."""
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from ecrad_b200 import inputs as I  # noqa: E402
from ecrad_b200.config import RadiationConfig  # noqa: E402
from ecrad_b200.radiation_interface import setup_radiation  # noqa: E402
from oracle_lib import Oracle  # noqa: E402

ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 32
solver = sys.argv[2] if len(sys.argv) > 2 else "McICA"
raw = {k: np.array(v, dtype=np.float64) for k, v in np.load(os.path.join(ROOT, "tests/golden/ecrad_meridian_inputs.npz")).items()}
raw = I.synthetic_columns(raw, ncol)
cfg = RadiationConfig(sw_solver_name=solver, lw_solver_name=solver, use_aerosols=len(sys.argv) > 3).consolidate()
h = setup_radiation(cfg)
for rep in range(2):
    t = time.time()
    out = h.radiation(I.to_radiation_inputs(raw), ncol, 137)
    dt = time.time() - t
print(f"gpu: {ncol} columns in {dt*1e3:.2f} ms  ({ncol/dt:.0f} col/s end-to-end, pageable host memory)")
print("stages ms:", h.last_stage_ms())
t = time.time()
ref = Oracle(cfg).radiation(I.to_radiation_inputs(raw), ncol, 137)
print(f"oracle: {time.time()-t:.2f} s")
bad = 0
for k in sorted(ref):
    if k not in out or ref[k] is None or out[k] is None:
        continue
    a, b = np.asarray(out[k]), np.asarray(ref[k])
    m = np.isfinite(b)
    err = np.abs(a[m] - b[m]).max() if m.any() else 0.0
    nan = int(np.isnan(a[m]).sum())
    flag = "" if (err <= 1e-6 and nan == 0) else "   <<<<<<"
    bad += bool(flag)
    print(f"{k:32s} max|gpu-oracle| = {err:.3e}  nan={nan}  max|ref|={np.abs(b[m]).max() if m.any() else 0:.4g}{flag}")
    if flag and a.ndim == 2 and a.shape[0] == ncol:
        e = np.abs(np.where(m, a - b, 0)).max(axis=1)
        print("      worst columns:", np.argsort(e)[-5:][::-1], e[np.argsort(e)[-5:][::-1]])
print("BAD" if bad else "ALL OK")
