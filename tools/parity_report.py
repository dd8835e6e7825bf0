#!/usr/bin/env python
"""Worst |flux - oracle| (W m-2) of the CUDA path per configuration, on synthetic IFS columns (run on the GPU box).

    python tools/parity_report.py [ncol]     -> one line per configuration: worst error over the ten flux profiles, and the rest
"""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from ecrad_b200 import inputs as I  # noqa: E402
from ecrad_b200.config import RadiationConfig  # noqa: E402
from ecrad_b200.radiation_interface import setup_radiation  # noqa: E402
from oracle_lib import Oracle  # noqa: E402

FLUX = ("lw_up", "lw_dn", "lw_up_clear", "lw_dn_clear", "sw_up", "sw_dn", "sw_dn_direct", "sw_up_clear", "sw_dn_clear", "sw_dn_direct_clear")
CASES = [("McICA RRTMG", dict()), ("McICA RRTMG aerosols", dict(use_aerosols=True)),
         ("McICA RRTMG delta-scaling+aer", dict(use_aerosols=True, do_sw_delta_scaling_with_gases=True)),
         ("McICA Exp-Exp no LW scat", dict(overlap_scheme_name="Exp-Exp", do_lw_cloud_scattering=False)),
         ("Cloudless RRTMG", dict(sw_solver_name="Cloudless", lw_solver_name="Cloudless")),
         ("Tripleclouds RRTMG", dict(sw_solver_name="Tripleclouds", lw_solver_name="Tripleclouds")),
         ("McICA ecCKD-32 aerosols", dict(gas_model_name="ECCKD", do_nearest_spectral_lw_emiss=False, use_aerosols=True)),
         ("Tripleclouds ecCKD-32", dict(gas_model_name="ECCKD", do_nearest_spectral_lw_emiss=False, sw_solver_name="Tripleclouds", lw_solver_name="Tripleclouds")),
         ("SPARTACUS RRTMG 3D", dict(sw_solver_name="SPARTACUS", lw_solver_name="SPARTACUS", do_3d_effects=True))]


def main():
    ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    only = sys.argv[2] if len(sys.argv) > 2 else ""
    raw0 = {k: np.array(v, dtype=np.float64) for k, v in np.load(os.path.join(ROOT, "tests/golden/ecrad_meridian_inputs.npz")).items()}
    raw = I.synthetic_columns(raw0, ncol)
    print(f"library: {os.environ.get('ECRAD_B200_LIB', 'ecrad_b200/libecrad_b200.so')}  scan={os.environ.get('ECRAD_B200_SCAN', 'default')}  columns: {ncol}")
    for name, kw in CASES:
        if only and only not in name:
            continue
        n = ncol if "SPARTACUS" not in name else min(ncol, 64)
        r = raw if n == ncol else I.synthetic_columns(raw0, n)
        cfg = RadiationConfig(**kw).consolidate()
        h = setup_radiation(cfg)
        out = h.radiation(I.to_radiation_inputs(r, cfg), n, 137)
        ref = Oracle(cfg).radiation(I.to_radiation_inputs(r, cfg), n, 137)
        h.finalize()
        worst, where, other = 0.0, "", 0.0
        for k in sorted(ref):
            if k not in out or ref[k] is None or out[k] is None:
                continue
            a, b = np.asarray(out[k], dtype=np.float64), np.asarray(ref[k], dtype=np.float64)
            m = np.isfinite(b)
            if not m.any():
                continue
            e = float(np.abs(a[m] - b[m]).max()) if not np.isnan(a[m]).any() else float("nan")
            if k in FLUX:
                if not e <= worst:
                    worst, where = e, k
            elif k != "cloud_fraction":
                other = max(other, e) if e == e else e
        print(f"{name:34s} worst flux profile error {worst:.3e} W m-2 ({where}), other outputs {other:.3e}   {'ok' if worst <= 5e-7 else 'ABOVE 5e-7'}")


if __name__ == "__main__":
    main()
