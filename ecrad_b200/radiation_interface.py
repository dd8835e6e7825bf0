"""Host-side mirror of `module radiation_interface` (radiation/radiation_interface.F90:29) over the C-ABI.

    setup_radiation(config)                         radiation_interface.F90:37   -> ecrad_b200_setup
    set_gas_units(config, gas)                      radiation_interface.F90:164  -> host side only (ecrad_b200.inputs.set_gas_units)
    radiation(ncol, nlev, istartcol, iendcol, ...)  radiation_interface.F90:200  -> ecrad_b200_radiation

This module only marshals numpy arrays into the POD structs of include/ecrad_b200.h and calls libecrad_b200.so.
There is no Python/numpy compute path: if the CUDA library is missing or no GPU is present the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import abi
from .config import RadiationConfig

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ECRAD_B200_LIB") or os.path.join(_HERE, "libecrad_b200.so")   # (the override is for A/B builds of the same library)
DEFAULT_TABLES = os.path.join(_HERE, "data", "rrtmg_tables.bin")

EXPORTS = [
    "ecrad_b200_tables_create", "ecrad_b200_tables_add", "ecrad_b200_tables_load_file", "ecrad_b200_tables_load_memory", "ecrad_b200_tables_free",
    "ecrad_b200_setup", "ecrad_b200_set_option", "ecrad_b200_radiation", "ecrad_b200_radiation_device", "ecrad_b200_radiation_device_ld",
    "ecrad_b200_kernel_launches",
    "ecrad_b200_last_stage_ms", "ecrad_b200_stage_name", "ecrad_b200_finalize", "ecrad_b200_last_error",
    "ecrad_b200_version", "ecrad_b200_measure_fp64", "ecrad_b200_radiation_blocked", "ecrad_b200_radiation_sp",
]

_lib = None


class RadiationError(RuntimeError):
    """Raised where the reference would call radiation_abort (utilities/radiation_io.F90:45-73)."""


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RadiationError(f"{LIB_PATH} not built: run `make -C ecrad_b200/csrc` (or __graft_entry__.build()); "
                             "there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    L.ecrad_b200_tables_create.restype = C.c_void_p
    L.ecrad_b200_tables_add.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_int64), C.c_void_p]
    L.ecrad_b200_tables_load_file.argtypes = [C.c_void_p, C.c_char_p]
    L.ecrad_b200_tables_load_memory.argtypes = [C.c_void_p, C.c_char_p, C.c_int64]
    L.ecrad_b200_tables_free.argtypes = [C.c_void_p]
    L.ecrad_b200_setup.argtypes = [C.POINTER(abi.Config), C.c_void_p, C.POINTER(C.c_void_p)]
    L.ecrad_b200_radiation.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(abi.Inputs), C.POINTER(abi.Outputs)]
    L.ecrad_b200_radiation_sp.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(abi.Inputs), C.POINTER(abi.Outputs)]
    L.ecrad_b200_radiation_device.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(abi.Inputs), C.POINTER(abi.Outputs), C.c_void_p]
    L.ecrad_b200_radiation_device_ld.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(abi.Inputs), C.POINTER(abi.Outputs), C.c_void_p]
    L.ecrad_b200_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    L.ecrad_b200_set_solar_cycle_multiplier.argtypes = [C.c_void_p, C.c_double]
    L.ecrad_b200_kernel_launches.restype = C.c_int64
    L.ecrad_b200_kernel_launches.argtypes = [C.c_void_p]
    L.ecrad_b200_last_stage_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_int]
    L.ecrad_b200_stage_name.restype = C.c_char_p
    L.ecrad_b200_stage_name.argtypes = [C.c_int]
    L.ecrad_b200_finalize.argtypes = [C.c_void_p]
    L.ecrad_b200_last_error.restype = C.c_char_p
    L.ecrad_b200_last_error.argtypes = [C.c_void_p]
    L.ecrad_b200_version.restype = C.c_char_p
    L.ecrad_b200_measure_fp64.argtypes = [C.POINTER(C.c_double)]
    L.ecrad_b200_save_radiative_properties.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(abi.Inputs), C.POINTER(abi.RadiativeProperties)]
    L.ecrad_b200_radiation_blocked.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(abi.BlockLayout), abi.c_dp, abi.c_dp]
    _lib = L
    return L


class RadiationHandle:
    """What `setup_radiation` leaves behind: the config plus the device-side tables (radiation_interface.F90:37-156)."""

    def __init__(self, config: RadiationConfig, tables_path: str = None, tables_blob: bytes = None, tables_arrays: dict = None):
        L = load_library()
        self.lib = L
        self.config = config
        tables_path = tables_path or config.tables_path()
        self.cfg = config.to_struct()
        t = L.ecrad_b200_tables_create()
        try:
            if tables_arrays is not None:
                # what the Fortran shim does (fortran/radiation_b200.F90): every array through ecrad_b200_tables_add, no blob
                for nm, arr in tables_arrays.items():
                    a = np.asfortranarray(arr)
                    code = 1 if a.dtype.kind in "iu" else 0
                    a = a.astype(np.int32 if code else np.float64, order="F")
                    dims = (C.c_int64 * 4)(*(list(a.shape) + [1] * (4 - a.ndim)))
                    if L.ecrad_b200_tables_add(t, nm.encode(), code, a.ndim, dims, a.ctypes.data_as(C.c_void_p)):
                        raise RadiationError(f"cannot add table {nm}")
                rc = 0
            else:
                rc = (L.ecrad_b200_tables_load_memory(t, tables_blob, len(tables_blob)) if tables_blob is not None
                      else L.ecrad_b200_tables_load_file(t, tables_path.encode()))
            if rc:
                raise RadiationError(L.ecrad_b200_last_error(None).decode())
            # config%sw_albedo_weights, config%i_emiss_from_band_lw: the part of config_type that is a table
            for nm, arr in config.derived.items():
                a = np.asfortranarray(arr)
                code = 1 if a.dtype.kind in "iu" else 0
                a = a.astype(np.int32 if code else np.float64, order="F")
                dims = (C.c_int64 * 4)(*(list(a.shape) + [1] * (4 - a.ndim)))
                if L.ecrad_b200_tables_add(t, nm.encode(), code, a.ndim, dims, a.ctypes.data_as(C.c_void_p)):
                    raise RadiationError(f"cannot add table {nm}")
            h = C.c_void_p()
            if L.ecrad_b200_setup(C.byref(self.cfg), t, C.byref(h)):
                raise RadiationError(L.ecrad_b200_last_error(None).decode())
            self.h = h
        finally:
            L.ecrad_b200_tables_free(t)

    def _err(self):
        return self.lib.ecrad_b200_last_error(self.h).decode()

    def radiative_properties(self, inputs, ncol, nlev, istartcol=1, iendcol=None):
        """What radiation() passes to save_radiative_properties (radiation_interface.F90:405-425): the optical properties the gas,
        aerosol and cloud optics hand to the solvers, as a dict of Fortran-ordered arrays (g-point or band, level, column).  The
        caller's cloud_fraction is left uncropped."""
        iendcol = ncol if iendcol is None else iendcol
        keep, ist = abi.make_inputs(dict(inputs, cloud_fraction=np.array(inputs["cloud_fraction"], order="F")), inputs["solar_irradiance"])
        arrs, pst = abi.alloc_radiative_properties(ncol, nlev, self.cfg)
        if self.lib.ecrad_b200_save_radiative_properties(self.h, ncol, nlev, istartcol, iendcol, C.byref(ist), C.byref(pst)):
            raise RadiationError(self._err())
        return arrs

    def radiation(self, inputs, ncol, nlev, istartcol=1, iendcol=None, outputs=None, spectral_profiles=False):
        """inputs: dict keyed like abi.INPUT_ARRAYS (+ 'solar_irradiance'); cloud_fraction is cropped in place like
        cloud%crop_cloud_fraction.  Returns the dict of flux_type components (Fortran-ordered float64)."""
        iendcol = ncol if iendcol is None else iendcol
        keep, ist = abi.make_inputs(inputs, inputs["solar_irradiance"])
        if outputs is None:
            outs, ost = abi.alloc_outputs(ncol, nlev, self.cfg, spectral_profiles=spectral_profiles)
        else:
            outs, ost = outputs
        rc = self.lib.ecrad_b200_radiation(self.h, ncol, nlev, istartcol, iendcol, C.byref(ist), C.byref(ost))
        if rc:
            raise RadiationError(self._err())
        outs["cloud_fraction"] = keep.get("cloud_fraction")
        return outs

    def radiation_sp(self, inputs, ncol, nlev, istartcol=1, iendcol=None, spectral_profiles=False):
        """The call a single-precision host makes (ecrad_b200_radiation_sp): every real array is float32 on the host side.  Returns the
        flux dict as float32 arrays (plus the cropped float32 cloud_fraction)."""
        iendcol = ncol if iendcol is None else iendcol
        st = abi.Inputs()
        st.struct_bytes = C.sizeof(abi.Inputs)
        st.solar_irradiance = float(inputs["solar_irradiance"])
        keep = {}
        for nm, dt, _ in abi.INPUT_ARRAYS:
            a = inputs.get(nm)
            if a is None:
                continue
            a = np.asfortranarray(a, dtype=np.int32 if dt == "i4" else np.float32)
            keep[nm] = a
            setattr(st, nm, C.cast(a.ctypes.data, abi.c_ip if dt == "i4" else abi.c_dp))
        outs, ost = {}, abi.Outputs()
        ost.struct_bytes = C.sizeof(abi.Outputs)
        for nm, kind in abi.OUTPUT_ARRAYS:
            if kind in ("pl", "ps") and not spectral_profiles:
                continue
            a = np.full(abi.output_shape(kind, ncol, nlev, self.cfg), np.nan, dtype=np.float32, order="F")
            if nm.startswith("cloud_cover"):
                a[...] = -1.0
            outs[nm] = a
            setattr(ost, nm, C.cast(a.ctypes.data, abi.c_dp))
        if self.lib.ecrad_b200_radiation_sp(self.h, ncol, nlev, istartcol, iendcol, C.byref(st), C.byref(ost)):
            raise RadiationError(self._err())
        outs["cloud_fraction"] = keep.get("cloud_fraction")
        return outs

    def radiation_blocked(self, inputs, ncol, nlev, nproma, spectral_profiles=False):
        """The same call through the blocked (NPROMA) entry: inputs are packed into zrgp(nproma, nfields, nblocks) as an IFS-style driver
        holds them (driver/ifs_blocking.F90), ecrad_b200_radiation_blocked does the rest; returns the flux dict."""
        outs, _ = abi.alloc_outputs(ncol, nlev, self.cfg, spectral_profiles=spectral_profiles)
        lay, zin, zout = abi.pack_blocked(inputs, outs, ncol, nlev, nproma, self.cfg, inputs["solar_irradiance"])
        rc = self.lib.ecrad_b200_radiation_blocked(self.h, ncol, nlev, C.byref(lay), zin.ctypes.data_as(abi.c_dp), zout.ctypes.data_as(abi.c_dp))
        if rc:
            raise RadiationError(self._err())
        abi.unpack_blocked(lay, zout, outs, ncol, nlev)
        return outs

    def radiation_device(self, ncol, nlev, ist: abi.Inputs, ost: abi.Outputs, stream=0):
        """Device-resident entry: every pointer in ist/ost is a device pointer with leading dimension ncol."""
        rc = self.lib.ecrad_b200_radiation_device(self.h, ncol, nlev, C.byref(ist), C.byref(ost), C.c_void_p(stream))
        if rc:
            raise RadiationError(self._err())

    def radiation_device_ld(self, ncol, nlev, ld_in, ld_out, ist: abi.Inputs, ost: abi.Outputs, stream=0):
        """As radiation_device with separate leading dimensions (outputs may be a column slice of peer-GPU arrays)."""
        rc = self.lib.ecrad_b200_radiation_device_ld(self.h, ncol, nlev, ld_in, ld_out, C.byref(ist), C.byref(ost), C.c_void_p(stream))
        if rc:
            raise RadiationError(self._err())

    def set_solar_cycle_multiplier(self, multiplier):
        """single_level%spectral_solar_cycle_multiplier for the calls that follow (-1 solar minimum .. +1 maximum; ecCKD shortwave)."""
        if self.lib.ecrad_b200_set_solar_cycle_multiplier(self.h, float(multiplier)):
            raise RadiationError(self._err())

    def set_option(self, key, value):
        if self.lib.ecrad_b200_set_option(self.h, key.encode(), int(value)):
            raise RadiationError(self._err())

    def kernel_launches(self):
        return int(self.lib.ecrad_b200_kernel_launches(self.h))

    def last_stage_ms(self):
        buf = (C.c_float * 8)()
        n = self.lib.ecrad_b200_last_stage_ms(self.h, buf, 8)
        return {self.lib.ecrad_b200_stage_name(i).decode(): float(buf[i]) for i in range(n)}

    def finalize(self):
        if getattr(self, "h", None):
            self.lib.ecrad_b200_finalize(self.h)
            self.h = None

    def __del__(self):
        try:
            self.finalize()
        except Exception:
            pass


def setup_radiation(config: RadiationConfig, tables_path: str = None, tables_blob: bytes = None, tables_arrays: dict = None) -> RadiationHandle:
    """tables_blob: the ETB1 image as bytes (e.g. received by a broadcast) instead of a file path; tables_arrays: {name: ndarray},
    each registered with ecrad_b200_tables_add like the Fortran shim does."""
    if not config.derived:
        config.consolidate()
    return RadiationHandle(config, tables_path, tables_blob, tables_arrays)


def set_gas_units(config: RadiationConfig, gas: dict) -> dict:
    """Scale the gas concentrations to the units the configured gas-optics model wants (radiation_interface.F90:164-193); stays on
    the host, like the reference's.  `gas`: file variables (`q`, `o3_mmr`, `*_vmr`); returns the C-ABI gas arrays."""
    from .inputs import set_gas_units as _set

    return _set(gas, config.is_ecckd)


def radiation(handle: RadiationHandle, ncol, nlev, istartcol, iendcol, inputs, **kw):
    return handle.radiation(inputs, ncol, nlev, istartcol, iendcol, **kw)
