#!/usr/bin/env python
"""Copy the reference's own regression fixtures into small committed files under tests/golden/.

Run in the authoring container (needs /root/reference); the GPU box only sees the .npz files.
  inputs : test/ifs/ecrad_meridian.nc                          -> tests/golden/ecrad_meridian_inputs.npz
  golden : test/ifs/ecrad_meridian_noaer_out_REFERENCE.nc      -> tests/golden/ecrad_meridian_noaer_ref.npz
           test/ifs/ecrad_meridian_cloudless_out_REFERENCE.nc  -> tests/golden/ecrad_meridian_cloudless_ref.npz
           ... and the default, expexp, tripleclouds, ecckd_mcica, ecckd_tc reference outputs likewise
The golden outputs are float32 as written by the reference driver (do_write_double_precision=false); the
per-band profiles of the cloudless file are kept at 8 half-levels only to keep the fixture small.
"""
import os
import sys

import numpy as np
from scipy.io import netcdf_file

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from ecrad_b200.inputs import NC_VARS  # noqa: E402

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
BAND_LEVELS = [0, 20, 40, 60, 80, 100, 120, 137]

os.makedirs(OUT, exist_ok=True)
with netcdf_file(f"{REF}/test/ifs/ecrad_meridian.nc", mmap=False) as f:
    np.savez_compressed(f"{OUT}/ecrad_meridian_inputs.npz", **{k: np.array(f.variables[k][...]) for k in NC_VARS})
for name in ("noaer", "cloudless", "default", "expexp", "tripleclouds", "ecckd_mcica", "ecckd_tc"):
    with netcdf_file(f"{REF}/test/ifs/ecrad_meridian_{name}_out_REFERENCE.nc", mmap=False) as f:
        d = {}
        for k, v in f.variables.items():
            a = np.array(v[...])
            if a.ndim == 3:  # (column, half_level, band)
                a = a[:, BAND_LEVELS, :]
            d[k] = a
        d["band_levels"] = np.array(BAND_LEVELS)
        d["history"] = np.array(getattr(f, "history", b"").decode(errors="replace"))
        np.savez_compressed(f"{OUT}/ecrad_meridian_{name}_ref.npz", **d)
print(os.listdir(OUT))
