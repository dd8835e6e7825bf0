#!/bin/bash
# compute-sanitizer memcheck over the paths added late in round 2: mixed gas models (both directions, aerosols, more than one tile of
# night and day columns), use_general_cloud_optics on RRTMG-IFS, save_radiative_properties, the spectral solar cycle.
#   tools/sanitize_round2_paths.sh [out-dir]
OUT=${1:-gpurun_out}
mkdir -p $OUT
cat > /tmp/san_case2.py <<'P'
import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from ecrad_b200 import inputs as I
from ecrad_b200.config import RadiationConfig
from ecrad_b200.radiation_interface import setup_radiation
raw0 = {k: np.array(v, dtype=np.float64) for k, v in np.load('tests/golden/ecrad_meridian_inputs.npz').items()}
n = 96
raw = I.synthetic_columns(raw0, n)
E = dict(do_nearest_spectral_lw_emiss=False)
TC = dict(sw_solver_name='Tripleclouds', lw_solver_name='Tripleclouds')
SP = dict(sw_solver_name='SPARTACUS', lw_solver_name='SPARTACUS', do_3d_effects=True)
for kw in (dict(sw_gas_model_name='ECCKD', use_aerosols=True, **E), dict(lw_gas_model_name='ECCKD', use_aerosols=True, **E, **TC),
           dict(sw_gas_model_name='ECCKD', **E, **SP), dict(use_general_cloud_optics=True, do_lw_cloud_scattering=False, use_aerosols=True),
           dict(gas_model_name='ECCKD', use_aerosols=True, **E, **TC)):
    cfg = RadiationConfig(**kw).consolidate()
    h = setup_radiation(cfg)
    h.set_option('tile_cols', 40); h.set_option('edge_cols', 16)
    out = h.radiation(I.to_radiation_inputs(raw, cfg), n, 137)
    p = h.radiative_properties(I.to_radiation_inputs(raw, cfg), n, 137, istartcol=2, iendcol=n - 1)
    if cfg.is_ecckd_sw:
        h.set_solar_cycle_multiplier(1.0)
        out = h.radiation(I.to_radiation_inputs(raw, cfg), n, 137)
    print(sorted(kw.items())[:3], float(np.nansum(out['sw_up'])), float(np.nansum(out['lw_up'])), float(np.nansum(p['od_sw'])))
    h.finalize()
P
timeout 2400 compute-sanitizer --tool memcheck --print-limit 20 python /tmp/san_case2.py > $OUT/sanitizer_round2_paths.log 2>&1
echo "== memcheck: exit $?"; grep -E "ERROR SUMMARY|Invalid|gas_model|general" $OUT/sanitizer_round2_paths.log | head -20
