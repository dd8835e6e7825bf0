"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol include/ecrad_b200.h declares,
ctypes structs match the header's layout, and setup fails loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from ecrad_b200 import abi
from ecrad_b200.config import RadiationConfig

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge

    ge.build()
    from ecrad_b200.radiation_interface import load_library

    return load_library()


def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "ecrad_b200.h")).read()
    declared = set(re.findall(r"\b(ecrad_b200_[a-z_0-9]+)\s*\(", hdr))
    assert len(declared) >= 13
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ecrad_b200.h but not exported"


def test_struct_layout_matches_header():
    """Compile a tiny C program against the header and compare sizeof/offsetof with the ctypes mirror."""
    import subprocess
    import tempfile

    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "ecrad_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu\n", sizeof(ecrad_b200_config), sizeof(ecrad_b200_inputs), sizeof(ecrad_b200_outputs),
         offsetof(ecrad_b200_config, cloud_fraction_threshold), offsetof(ecrad_b200_inputs, cloud_fraction),
         offsetof(ecrad_b200_outputs, sw_dn_direct_band));
  return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")])
        vals = [int(x) for x in subprocess.check_output([os.path.join(d, "t")]).split()]
    assert vals == [C.sizeof(abi.Config), C.sizeof(abi.Inputs), C.sizeof(abi.Outputs), abi.Config.cloud_fraction_threshold.offset,
                    abi.Inputs.cloud_fraction.offset, abi.Outputs.sw_dn_direct_band.offset]


def test_setup_fails_loudly_without_gpu(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from ecrad_b200.radiation_interface import RadiationError, setup_radiation

    with pytest.raises(RadiationError, match="no CUDA device"):
        setup_radiation(RadiationConfig().consolidate())


def test_config_derived_tables():
    """consolidate(): sw_albedo_weights columns sum to 1; emissivity intervals follow the namelist bounds."""
    cfg = RadiationConfig().consolidate()
    w = cfg.derived["sw_albedo_weights"]
    assert w.shape == (6, 14) and np.allclose(w.sum(axis=0), 1.0)
    assert list(cfg.derived["i_emiss_from_band_lw"]) == [1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1]


def test_radiative_properties_struct_matches_the_header():
    """abi.RadiativeProperties lists the 18 pointers of struct ecrad_b200_radiative_properties in the header's order
    (= the argument list of save_radiative_properties, radiation_save.F90:716-726)."""
    import re

    from ecrad_b200 import abi

    hdr = open(os.path.join(ROOT, "include", "ecrad_b200.h")).read()
    body = re.search(r"typedef struct ecrad_b200_radiative_properties \{(.*?)\} ecrad_b200_radiative_properties;", hdr, re.S).group(1)
    names = re.findall(r"double\*\s*(\w+);", re.sub(r"/\*.*?\*/", "", body, flags=re.S))
    assert names == [nm for nm, _ in abi.RADPROP_ARRAYS] and len(names) == 18
    assert C.sizeof(abi.RadiativeProperties) == 18 * C.sizeof(C.c_void_p)
