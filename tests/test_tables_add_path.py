"""The table-directory route of the Fortran shim (fortran/radiation_b200.F90): every array handed over with
ecrad_b200_tables_add, no ETB1 blob -- must give bit-identical results to the blob route; and the shim itself must be complete
(generated from the header and the blob's name list, no elisions)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from ecrad_b200 import inputs as I
from ecrad_b200 import tables
from ecrad_b200.config import RadiationConfig

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "fortran", "radiation_b200.F90")


def test_shim_is_complete_and_current():
    src = open(SHIM).read()
    assert "..." not in src, "the shim must not elide anything"
    # every array of the RRTMG blob is either registered from ifsrrtm module storage or from config_type
    blob = tables.read_blob(os.path.join(ROOT, "ecrad_b200", "data", "rrtmg_tables.bin"))
    registered = set(re.findall(r"call add_[ri]\d\(t, '([A-Za-z0-9_]+)'", src))
    # ("<name>.<model>" entries are the stand-alone blob's copies of the other cloud-optics files; a host registers the configured pair)
    host_side = {nm.split(".")[0] for nm in blob}
    # the general cloud optics tables go through register_gco(t, '<prefix>', config%cloud_optics_xx(jtype)), for RRTMG-IFS too
    for pre in set(re.findall(r"call register_gco\(t, '(gco_[ls]w_[01]_)'", src)):
        registered |= {pre + f for f in ("meta", "mass_ext", "ssa", "asymmetry")}
    assert host_side <= registered, sorted(host_side - registered)
    # all 41 outputs and every input pointer of the header are assigned
    hdr = open(os.path.join(ROOT, "include", "ecrad_b200.h")).read()
    outs = re.search(r"typedef struct ecrad_b200_outputs \{(.*?)\} ecrad_b200_outputs;", hdr, re.S).group(1)
    outs = re.sub(r"/\*.*?\*/", "", outs, flags=re.S)
    names = re.findall(r"\*(\w+)", outs)
    assert len(names) == 41
    for k, nm in enumerate(names, 1):
        assert f"o%p({k}) = c_loc(flux%{nm})" in src, nm
    if os.path.isdir("/root/reference"):   # regenerate and compare (the generator needs the reference's module declarations)
        before = src
        subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "gen_fortran_shim.py")], stdout=subprocess.DEVNULL)
        assert open(SHIM).read() == before, "fortran/radiation_b200.F90 is stale: run tools/gen_fortran_shim.py"


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(use_aerosols=True), dict(gas_model_name="ECCKD", do_nearest_spectral_lw_emiss=False, use_aerosols=True,
                                                            sw_solver_name="Tripleclouds", lw_solver_name="Tripleclouds")])
def test_tables_add_route_is_bit_identical_to_the_blob_route(meridian_raw, kw):
    from ecrad_b200.radiation_interface import setup_radiation

    cfg = RadiationConfig(**kw).consolidate()
    n = 64
    raw = I.synthetic_columns(meridian_raw, n)
    h1 = setup_radiation(cfg)
    a = h1.radiation(I.to_radiation_inputs(raw, cfg), n, 137)
    h1.finalize()
    arrays = tables.read_blob(cfg.tables_path())   # {name: ndarray in Fortran logical shape}: what the shim passes, array by array
    h2 = setup_radiation(cfg, tables_arrays=arrays)
    b = h2.radiation(I.to_radiation_inputs(raw, cfg), n, 137)
    h2.finalize()
    for nm in a:
        if a[nm] is not None:
            assert np.array_equal(a[nm], b[nm], equal_nan=True), nm
