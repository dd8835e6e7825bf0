"""Debug helper: one SPARTACUS call on the 32-column slice (run under compute-sanitizer on the GPU box)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from ecrad_b200 import inputs as I
from ecrad_b200.config import RadiationConfig
from ecrad_b200.radiation_interface import setup_radiation
raw = {k: np.array(v, dtype=np.float64) for k, v in np.load(os.path.join(ROOT, "tests/golden/ecrad_meridian_inputs.npz")).items()}
kw = dict(sw_solver_name="SPARTACUS", lw_solver_name="SPARTACUS", do_3d_effects=True)
if len(sys.argv) > 1 and sys.argv[1] == "lw": kw["do_sw"] = False
if len(sys.argv) > 1 and sys.argv[1] == "sw": kw["do_lw"] = False
cfg = RadiationConfig(**kw).consolidate()
h = setup_radiation(cfg)
h.set_option("serial", 1)
out = h.radiation(I.to_radiation_inputs(raw, cfg), 32, 137)
print("ok", np.nanmax(out["lw_up"]) if cfg.do_lw else None, np.nanmax(out["sw_up"]) if cfg.do_sw else None)
