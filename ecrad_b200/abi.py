"""ctypes mirror of include/ecrad_b200.h (the C-ABI structs).  Field order/types must match the header."""
from __future__ import annotations

import ctypes as C

import numpy as np

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)

# enum values = the reference's `enum, bind(c)` blocks (radiation_config.F90:51-111, radiation_cloud_cover.F90:30-33)
SOLVER = {"cloudless": 0, "homogeneous": 1, "mcica": 2, "spartacus": 3, "tripleclouds": 4}
GAS_MODEL = {"monochromatic": 0, "rrtmg-ifs": 1, "ecckd": 2}
OVERLAP = {"max-ran": 0, "exp-ran": 1, "exp-exp": 2}
# liquid_model_name / ice_model_name (radiation_config.F90:108-133); Jahangir and Nielsen are in the reference's enumeration but
# radiation_cloud_optics.F90:345-372 aborts on them, so they are not offered here either
LIQ_MODEL = {"socrates": 1, "slingo": 2}
ICE_MODEL = {"fu-ifs": 1, "baran-experimental": 2, "baran2016": 3, "baran2017-experimental": 4, "yi": 5}
# sw_entrapment_name (radiation_config.F90:69-84)
PDF_SHAPE = {"lognormal": 0, "gamma": 1}   # cloud_pdf_shape_name (radiation_config.F90:134-142)
ENTRAPMENT = {"zero": 0, "edge-only": 1, "explicit": 2, "non-fractal": 3, "maximum": 4}


class Config(C.Structure):
    _fields_ = [
        ("struct_bytes", C.c_int32),
        ("i_solver_sw", C.c_int32), ("i_solver_lw", C.c_int32),
        ("i_gas_model_sw", C.c_int32), ("i_gas_model_lw", C.c_int32),
        ("i_overlap_scheme", C.c_int32),
        ("i_liq_model", C.c_int32), ("i_ice_model", C.c_int32),
        ("do_sw", C.c_int32), ("do_lw", C.c_int32), ("do_sw_direct", C.c_int32), ("do_clear", C.c_int32),
        ("do_clouds", C.c_int32), ("use_aerosols", C.c_int32),
        ("do_lw_cloud_scattering", C.c_int32), ("do_lw_aerosol_scattering", C.c_int32), ("do_lw_derivatives", C.c_int32),
        ("do_sw_delta_scaling_with_gases", C.c_int32), ("do_fu_lw_ice_optics_bug", C.c_int32),
        ("use_beta_overlap", C.c_int32), ("use_vectorizable_generator", C.c_int32),
        ("do_surface_sw_spectral_flux", C.c_int32), ("do_canopy_fluxes_sw", C.c_int32),
        ("do_canopy_fluxes_lw", C.c_int32), ("do_save_spectral_flux", C.c_int32),
        ("do_nearest_spectral_sw_albedo", C.c_int32), ("do_nearest_spectral_lw_emiss", C.c_int32),
        ("n_g_sw", C.c_int32), ("n_g_lw", C.c_int32), ("n_bands_sw", C.c_int32), ("n_bands_lw", C.c_int32),
        ("n_albedo_sw", C.c_int32), ("n_emiss_lw", C.c_int32),
        ("n_canopy_bands_sw", C.c_int32), ("n_canopy_bands_lw", C.c_int32),
        ("n_aerosol_types", C.c_int32),
        ("do_3d_effects", C.c_int32), ("i_3d_sw_entrapment", C.c_int32), ("do_3d_lw_multilayer_effects", C.c_int32),
        ("cloud_fraction_threshold", C.c_double), ("cloud_mixing_ratio_threshold", C.c_double),
        ("min_gas_od_lw", C.c_double), ("min_gas_od_sw", C.c_double),
        ("cloud_inhom_decorr_scaling", C.c_double),
        ("max_gas_od_3d", C.c_double), ("max_cloud_od", C.c_double), ("max_3d_transfer_rate", C.c_double),
        ("min_cloud_effective_size", C.c_double),
        ("overhead_sun_factor", C.c_double), ("overhang_factor", C.c_double), ("clear_to_thick_fraction", C.c_double),
        ("do_lw_side_emissivity", C.c_int32), ("use_expm_everywhere", C.c_int32),
        ("do_toa_spectral_flux", C.c_int32), ("i_cloud_pdf_shape", C.c_int32),
        ("n_regions", C.c_int32), ("use_general_cloud_optics", C.c_int32),
    ]


INPUT_ARRAYS = [
    # name, dtype, shape-kind
    ("cos_sza", "f8", "c"), ("skin_temperature", "f8", "c"),
    ("sw_albedo", "f8", "ca"), ("sw_albedo_direct", "f8", "ca"), ("lw_emissivity", "f8", "ce"),
    ("iseed", "i4", "c"),
    ("pressure_hl", "f8", "ch"), ("temperature_hl", "f8", "ch"),
    ("h2o_mmr", "f8", "cl"), ("co2_mmr", "f8", "cl"), ("o3_mmr", "f8", "cl"), ("n2o_mmr", "f8", "cl"),
    ("ch4_mmr", "f8", "cl"), ("cfc11_mmr", "f8", "cl"), ("cfc12_mmr", "f8", "cl"), ("hcfc22_mmr", "f8", "cl"),
    ("ccl4_mmr", "f8", "cl"),
    ("cloud_fraction", "f8", "cl"), ("q_liq", "f8", "cl"), ("q_ice", "f8", "cl"), ("re_liq", "f8", "cl"),
    ("re_ice", "f8", "cl"), ("overlap_param", "f8", "ci"), ("fractional_std", "f8", "cl"),
    ("aerosol_mmr", "f8", "clt"), ("h2o_sat_liq", "f8", "cl"),
    ("inv_cloud_effective_size", "f8", "cl"), ("inv_inhom_effective_size", "f8", "cl"),
]


class Inputs(C.Structure):
    _fields_ = [("struct_bytes", C.c_int32), ("reserved", C.c_int32), ("solar_irradiance", C.c_double)] + [
        (nm, c_ip if dt == "i4" else c_dp) for nm, dt, _ in INPUT_ARRAYS
    ]


# name, shape-kind: h = (ncol, nlev+1); c = (ncol); gl/gs = (ng, ncol); bs = (nbands_sw, ncol); as/al = canopy;
# pl/ps = per-band profiles (nband, ncol, nlev+1)
OUTPUT_ARRAYS = [
    ("lw_up", "h"), ("lw_dn", "h"), ("lw_up_clear", "h"), ("lw_dn_clear", "h"),
    ("sw_up", "h"), ("sw_dn", "h"), ("sw_dn_direct", "h"),
    ("sw_up_clear", "h"), ("sw_dn_clear", "h"), ("sw_dn_direct_clear", "h"),
    ("lw_derivatives", "h"),
    ("cloud_cover_lw", "c"), ("cloud_cover_sw", "c"),
    ("lw_dn_surf_g", "gl"), ("lw_dn_surf_clear_g", "gl"), ("lw_up_toa_g", "gl"), ("lw_up_toa_clear_g", "gl"),
    ("sw_dn_diffuse_surf_g", "gs"), ("sw_dn_direct_surf_g", "gs"),
    ("sw_dn_diffuse_surf_clear_g", "gs"), ("sw_dn_direct_surf_clear_g", "gs"),
    ("sw_up_toa_g", "gs"), ("sw_up_toa_clear_g", "gs"),
    ("sw_dn_surf_band", "bs"), ("sw_dn_direct_surf_band", "bs"),
    ("sw_dn_surf_clear_band", "bs"), ("sw_dn_direct_surf_clear_band", "bs"),
    ("sw_dn_diffuse_surf_canopy", "as"), ("sw_dn_direct_surf_canopy", "as"),
    ("lw_dn_surf_canopy", "al"),
    ("lw_up_band", "pl"), ("lw_dn_band", "pl"), ("sw_up_band", "ps"), ("sw_dn_band", "ps"), ("sw_dn_direct_band", "ps"),
    ("sw_dn_toa_g", "gs"), ("sw_dn_toa_band", "bs"), ("sw_up_toa_band", "bs"), ("sw_up_toa_clear_band", "bs"),
    ("lw_up_toa_band", "bl"), ("lw_up_toa_clear_band", "bl"),
]


class Outputs(C.Structure):
    _fields_ = [("struct_bytes", C.c_int32), ("reserved", C.c_int32)] + [(nm, c_dp) for nm, _ in OUTPUT_ARRAYS]


# ecrad_b200_radiative_properties (include/ecrad_b200.h): name, shape kind -- the argument list of save_radiative_properties
# (radiation_save.F90:716-726)
RADPROP_ARRAYS = [
    ("planck_hl", "glh"), ("lw_emission", "gl"), ("lw_albedo", "gl"), ("sw_albedo_direct", "gs"), ("sw_albedo_diffuse", "gs"),
    ("incoming_sw", "gs"), ("od_lw", "glf"), ("ssa_lw", "glf"), ("g_lw", "glf"), ("od_sw", "gsf"), ("ssa_sw", "gsf"), ("g_sw", "gsf"),
    ("od_lw_cloud", "blf"), ("ssa_lw_cloud", "blf"), ("g_lw_cloud", "blf"), ("od_sw_cloud", "bsf"), ("ssa_sw_cloud", "bsf"), ("g_sw_cloud", "bsf"),
]


class RadiativeProperties(C.Structure):
    _fields_ = [(nm, c_dp) for nm, _ in RADPROP_ARRAYS]


def alloc_radiative_properties(ncol, nlev, cfg):
    """NaN-filled Fortran-ordered arrays (spectral index fastest, column slowest) + the struct that points at them."""
    shapes = {"glh": (cfg.n_g_lw, nlev + 1, ncol), "gl": (cfg.n_g_lw, ncol), "gs": (cfg.n_g_sw, ncol), "glf": (cfg.n_g_lw, nlev, ncol),
              "gsf": (cfg.n_g_sw, nlev, ncol), "blf": (cfg.n_bands_lw, nlev, ncol), "bsf": (cfg.n_bands_sw, nlev, ncol)}
    arrs, st = {}, RadiativeProperties()
    for nm, kind in RADPROP_ARRAYS:
        a = np.full(shapes[kind], np.nan, dtype=np.float64, order="F")
        arrs[nm] = a
        setattr(st, nm, a.ctypes.data_as(c_dp))
    return arrs, st


def output_shape(kind, ncol, nlev, cfg):
    """Fortran shape of a flux_type component (radiation_flux.F90:147-300)."""
    return {
        "h": (ncol, nlev + 1), "c": (ncol,), "gl": (cfg.n_g_lw, ncol), "gs": (cfg.n_g_sw, ncol),
        "bs": (cfg.n_bands_sw, ncol), "bl": (cfg.n_bands_lw, ncol), "as": (cfg.n_canopy_bands_sw, ncol), "al": (cfg.n_canopy_bands_lw, ncol),
        "pl": (cfg.n_bands_lw, ncol, nlev + 1), "ps": (cfg.n_bands_sw, ncol, nlev + 1),
    }[kind]


def alloc_outputs(ncol, nlev, cfg, spectral_profiles=False, fill=np.nan):
    """Allocate every flux component as a Fortran-ordered float64 array; returns (dict, Outputs struct)."""
    arrs = {}
    st = Outputs()
    st.struct_bytes = C.sizeof(Outputs)
    for nm, kind in OUTPUT_ARRAYS:
        if kind in ("pl", "ps") and not spectral_profiles:
            continue
        a = np.full(output_shape(kind, ncol, nlev, cfg), fill, dtype=np.float64, order="F")
        if nm.startswith("cloud_cover"):
            a[...] = -1.0  # flux%allocate initial value seen in the golden files for night columns
        arrs[nm] = a
        setattr(st, nm, a.ctypes.data_as(c_dp))
    return arrs, st


def make_inputs(arrays, solar_irradiance):
    """arrays: dict of Fortran-ordered numpy arrays keyed like INPUT_ARRAYS; returns (keepalive dict, Inputs)."""
    st = Inputs()
    st.struct_bytes = C.sizeof(Inputs)
    st.solar_irradiance = float(solar_irradiance)
    keep = {}
    for nm, dt, _ in INPUT_ARRAYS:
        a = arrays.get(nm)
        if a is None:
            continue
        want = np.int32 if dt == "i4" else np.float64
        a = np.asfortranarray(a, dtype=want)
        keep[nm] = a
        setattr(st, nm, a.ctypes.data_as(c_ip if dt == "i4" else c_dp))
    return keep, st


class BlockLayout(C.Structure):
    """struct ecrad_b200_block_layout: where each input / output lives in zrgp(nproma, nfields, nblocks)."""
    _fields_ = [("struct_bytes", C.c_int32), ("nproma", C.c_int32), ("nblocks", C.c_int32), ("nfields_in", C.c_int32), ("nfields_out", C.c_int32),
                ("in_field", C.c_int32 * 28), ("out_field", C.c_int32 * 41), ("solar_irradiance", C.c_double)]


def pack_blocked(arrays, outputs, ncol, nlev, nproma, cfg, solar_irradiance):
    """Column-layout inputs (dict keyed like INPUT_ARRAYS) and outputs (dict from alloc_outputs) -> (layout, zrgp_in, zrgp_out) in the
    blocked layout of driver/ifs_blocking.F90: Fortran (nproma, nfields, nblocks), i.e. C order (nblocks, nfields, nproma)."""
    nblocks = (ncol + nproma - 1) // nproma
    lay = BlockLayout()
    lay.struct_bytes = C.sizeof(BlockLayout)
    lay.nproma, lay.nblocks, lay.solar_irradiance = nproma, nblocks, float(solar_irradiance)
    rows_in, f = [], 0
    for k, (nm, dt, _) in enumerate(INPUT_ARRAYS):
        a = arrays.get(nm)
        if a is None:
            lay.in_field[k] = -1
            rows_in.append(None)
            continue
        a2 = np.asarray(a, dtype=np.float64).reshape(ncol, -1, order="F")   # (ncol, rows): aerosol (ncol, nlev, ntype) -> type slowest
        lay.in_field[k] = f
        rows_in.append(a2)
        f += a2.shape[1]
    lay.nfields_in = f
    zin = np.zeros((nblocks, f, nproma))
    pad = nblocks * nproma - ncol
    for k, a2 in enumerate(rows_in):
        if a2 is not None:
            full = np.concatenate([a2, np.zeros((pad, a2.shape[1]))]) if pad else a2
            zin[:, lay.in_field[k]:lay.in_field[k] + a2.shape[1], :] = full.reshape(nblocks, nproma, -1).transpose(0, 2, 1)
    f, views = 0, []
    for k, (nm, kind) in enumerate(OUTPUT_ARRAYS):
        a = outputs.get(nm)
        if a is None:
            lay.out_field[k] = -1
            views.append(None)
            continue
        if kind in ("h", "c"):
            a2 = np.asarray(a).reshape(ncol, -1, order="F")                       # (ncol, rows)
        elif kind in ("pl", "ps"):
            a2 = np.transpose(np.asarray(a), (1, 2, 0)).reshape(ncol, -1)         # (nb, ncol, nlev+1) -> (ncol, [level][band])
        else:
            a2 = np.asarray(a).T                                                  # (n, ncol) -> (ncol, n)
        lay.out_field[k] = f
        views.append(a2)
        f += a2.shape[1]
    lay.nfields_out = f
    zout = np.zeros((nblocks, f, nproma))
    for k, a2 in enumerate(views):
        if a2 is not None:
            full = np.concatenate([a2, np.zeros((pad, a2.shape[1]))]) if pad else a2
            zout[:, lay.out_field[k]:lay.out_field[k] + a2.shape[1], :] = full.reshape(nblocks, nproma, -1).transpose(0, 2, 1)
    return lay, zin, zout


def unpack_blocked(lay, zout, outputs, ncol, nlev):
    """zrgp_out -> the column-layout flux arrays of `outputs` (in place)."""
    nproma, nblocks = lay.nproma, lay.nblocks
    for k, (nm, kind) in enumerate(OUTPUT_ARRAYS):
        a = outputs.get(nm)
        if a is None or lay.out_field[k] < 0:
            continue
        if kind in ("h", "c"):
            n = int(np.prod(a.shape[1:])) if a.ndim > 1 else 1
        elif kind in ("pl", "ps"):
            n = a.shape[0] * a.shape[2]
        else:
            n = a.shape[0]
        blk = zout[:, lay.out_field[k]:lay.out_field[k] + n, :].transpose(0, 2, 1).reshape(nblocks * nproma, n)[:ncol]
        if kind in ("h", "c"):
            a[...] = blk.reshape(a.shape, order="F")
        elif kind in ("pl", "ps"):
            a[...] = np.transpose(blk.reshape(ncol, a.shape[2], a.shape[0]), (2, 0, 1))
        else:
            a[...] = blk.T
