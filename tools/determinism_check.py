#!/usr/bin/env python
"""Run the same 10 000 columns repeatedly through the host entry and compare every output bit for bit (race detector)."""
import sys

import numpy as np

sys.path.insert(0, ".")
from ecrad_b200 import inputs as I  # noqa: E402
from ecrad_b200.config import RadiationConfig  # noqa: E402
from ecrad_b200.radiation_interface import setup_radiation  # noqa: E402

raw = {k: np.array(v, dtype=np.float64) for k, v in np.load("tests/golden/ecrad_meridian_inputs.npz").items()}
n, reps = 10000, int(sys.argv[1]) if len(sys.argv) > 1 else 12
kw = dict(a.split("=") for a in sys.argv[2:])
cfg = RadiationConfig(**kw).consolidate()
r = I.synthetic_columns(raw, n)
h = setup_radiation(cfg)
ref = h.radiation(I.to_radiation_inputs(r, cfg), n, 137)
bad = {}
for i in range(reps):
    out = h.radiation(I.to_radiation_inputs(r, cfg), n, 137)
    for nm, a in out.items():
        if isinstance(a, np.ndarray) and not np.array_equal(a, ref[nm], equal_nan=True):
            cols = np.unique(np.argwhere(~((a == ref[nm]) | (np.isnan(a) & np.isnan(ref[nm]))))[:, 0 if a.shape[0] == n else -1])
            idx = np.argwhere(~((a == ref[nm]) | (np.isnan(a) & np.isnan(ref[nm]))))
            bad.setdefault(nm, []).append((i, len(cols), cols[:5].tolist(), idx[:12].tolist(), [float(a[tuple(k)] - ref[nm][tuple(k)]) for k in idx[:4]]))
print("mismatches:", bad if bad else "none")
h.finalize()
