"""ctypes loader for the CPU oracle (oracle/liboracle.so).  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from ecrad_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
TABLES = os.path.join(ROOT, "ecrad_b200", "data", "rrtmg_tables.bin")


def build():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])
    return os.path.join(ORACLE_DIR, "liboracle.so")


def build_native():
    """The same C sources built the way the reference's own build compiles its Fortran (Makefile_include.gfortran:25 -O3, CMake adds
    -march=native; gfortran contracts multiply-adds by default): -O3 -march=native -fopenmp, FMA contraction on.  Only for TIMING the
    CPU arm of bench.py on the machine it runs on (results differ from the parity oracle in the last bits); built into oracle/_bench/."""
    out_dir = os.path.join(ORACLE_DIR, "_bench")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, "liboracle_native.so")
    srcs = [os.path.join(ORACLE_DIR, f) for f in sorted(os.listdir(ORACLE_DIR)) if f.endswith(".c")]
    newest = max(os.path.getmtime(f) for f in srcs + [os.path.join(ORACLE_DIR, "oracle.h")])
    if not os.path.exists(out) or os.path.getmtime(out) < newest:
        subprocess.check_call(["gcc", "-O3", "-march=native", "-fPIC", "-shared", "-fopenmp", "-std=c11", "-ffp-contract=fast",
                               "-I", os.path.join(ROOT, "include"), "-o", out] + srcs + ["-lm"])
    return out


class Oracle:
    def __init__(self, config, lib_path=None):
        path = lib_path or os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            build()
        self.lib = C.CDLL(path)
        L = self.lib
        L.orc_tables_load.restype = C.c_void_p
        L.orc_tables_load.argtypes = [C.c_char_p]
        L.orc_tables_add.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_int64), C.c_void_p]
        L.orc_tables_resolve.argtypes = [C.c_void_p]
        L.orc_tables_free.argtypes = [C.c_void_p]
        L.orc_set_solar_cycle_multiplier.argtypes = [C.c_void_p, C.c_double]
        L.orc_radiation.argtypes = [C.c_void_p, C.POINTER(abi.Config), C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.POINTER(abi.Inputs), C.POINTER(abi.Outputs), C.c_int]
        L.orc_radiative_properties.argtypes = [C.c_void_p, C.POINTER(abi.Config), C.c_int, C.c_int, C.c_int, C.c_int,
                                               C.POINTER(abi.Inputs), C.POINTER(abi.RadiativeProperties)]
        L.orc_gas_optics_column.argtypes = [C.c_void_p, C.POINTER(abi.Config), C.c_int, C.c_int, C.c_int,
                                            C.POINTER(abi.Inputs)] + [abi.c_dp] * 6
        L.orc_expm.argtypes = [C.c_int, abi.c_dp, C.c_int]
        L.orc_fast_expm_exchange_3.argtypes = [C.c_double] * 4 + [abi.c_dp]
        self.config = config
        self.cfg = config.to_struct()
        tables = config.tables_path()
        self.t = L.orc_tables_load(tables.encode())
        if not self.t:
            raise RuntimeError("oracle: cannot load " + tables)
        self._keep = []
        for nm, arr in config.derived.items():
            a = np.asfortranarray(arr)
            code = 1 if a.dtype.kind in "iu" else 0
            a = a.astype(np.int32 if code else np.float64, order="F")
            self._keep.append(a)
            dims = (C.c_int64 * 4)(*(list(a.shape) + [1] * (4 - a.ndim)))
            L.orc_tables_add(self.t, nm.encode(), code, a.ndim, dims, a.ctypes.data_as(C.c_void_p))
        if L.orc_tables_resolve(self.t):
            raise RuntimeError("oracle: table blob incomplete")

    def set_solar_cycle_multiplier(self, multiplier):
        if self.lib.orc_set_solar_cycle_multiplier(self.t, float(multiplier)):
            raise RuntimeError("oracle: no information present on solar cycle")

    def radiative_properties(self, inputs, ncol, nlev, istartcol=1, iendcol=None):
        """What radiation() hands to save_radiative_properties (radiation_interface.F90:405-425); cloud_fraction of `inputs` is cropped."""
        iendcol = ncol if iendcol is None else iendcol
        keep, ist = abi.make_inputs(inputs, inputs["solar_irradiance"])
        arrs, pst = abi.alloc_radiative_properties(ncol, nlev, self.cfg)
        rc = self.lib.orc_radiative_properties(self.t, C.byref(self.cfg), ncol, nlev, istartcol, iendcol, C.byref(ist), C.byref(pst))
        if rc:
            raise RuntimeError(f"oracle radiative_properties failed rc={rc}")
        return arrs

    def radiation(self, inputs, ncol, nlev, istartcol=1, iendcol=None, nthreads=0, spectral_profiles=False):
        """inputs: dict from ecrad_b200.inputs.to_radiation_inputs (cloud_fraction is modified in place)."""
        iendcol = ncol if iendcol is None else iendcol
        keep, ist = abi.make_inputs(inputs, inputs["solar_irradiance"])
        outs, ost = abi.alloc_outputs(ncol, nlev, self.cfg, spectral_profiles=spectral_profiles)
        rc = self.lib.orc_radiation(self.t, C.byref(self.cfg), ncol, nlev, istartcol, iendcol, C.byref(ist), C.byref(ost), nthreads)
        if rc:
            raise RuntimeError(f"oracle radiation failed rc={rc}")
        outs["cloud_fraction"] = keep["cloud_fraction"]
        return outs

    def gas_optics_column(self, inputs, ncol, nlev, jcol):
        keep, ist = abi.make_inputs(inputs, inputs["solar_irradiance"])
        od_lw = np.zeros((nlev, 140)); planck = np.zeros((nlev + 1, 140)); emis = np.zeros(140)
        od_sw = np.zeros((nlev, 112)); ssa_sw = np.zeros((nlev, 112)); inc = np.zeros(112)
        p = lambda a: a.ctypes.data_as(abi.c_dp)  # noqa: E731
        rc = self.lib.orc_gas_optics_column(self.t, C.byref(self.cfg), ncol, nlev, jcol, C.byref(ist),
                                            p(od_lw), p(planck), p(emis), p(od_sw), p(ssa_sw), p(inc))
        if rc:
            raise RuntimeError(f"oracle gas optics failed rc={rc}")
        return dict(od_lw=od_lw, planck_hl=planck, lw_emission=emis, od_sw=od_sw, ssa_sw=ssa_sw, incoming_sw=inc)

    def __del__(self):
        try:
            self.lib.orc_tables_free(self.t)
        except Exception:
            pass
