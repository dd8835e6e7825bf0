#!/usr/bin/env python
"""Generates fortran/radiation_b200.F90: the complete ISO_C_BINDING shim between the reference's `radiation_interface` and
libecrad_b200.so (include/ecrad_b200.h).  Everything that is a list is generated from the single source of truth it mirrors:

  * type(cfg_t)/in_t/out_t and the field assignments    <- the structs of include/ecrad_b200.h, field by field
  * the RRTMG table registrations                        <- the names and shapes of the ETB1 blob (tools/extract_rrtmg_tables.py),
                                                            matched against the declarations in ifsrrtm/yoerrta*.F90, yoesrta*.F90
                                                            (an array declared larger than what the C side wants is passed as a section)

    python tools/gen_fortran_shim.py [--ref /root/reference]

The generator refuses to write the file if a header field has no mapping or a blob array has no Fortran declaration.
"""
import argparse
import os
import re
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from ecrad_b200 import tables  # noqa: E402


# ----------------------------------------------------------------------------------------------------------------------
def parse_struct(hdr, name):
    """[(ctype, field)] of `typedef struct <name> {...}` in declaration order (comments stripped)."""
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), hdr, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    out = []
    for stmt in body.split(";"):
        stmt = " ".join(stmt.split())
        if not stmt:
            continue
        m = re.match(r"(const )?(int32_t|double)\s*(.*)", stmt)
        ctype, rest = m.group(2), m.group(3)
        for f in rest.split(","):
            f = f.strip()
            ptr = f.startswith("*") or ctype + "*" in stmt.replace(" ", "")[: len(ctype) + 7]
            out.append((ctype + ("*" if "*" in f or ptr else ""), f.replace("*", "").strip()))
    return out


# how each field of ecrad_b200_config is obtained from config_type (radiation/radiation_config.F90); logicals -> merge(1, 0, x)
CFG_SPECIAL = {
    "struct_bytes": "int(c_sizeof(c), c_int32_t)",
    "n_albedo_sw": "size(config%sw_albedo_weights, 1)",
    "n_emiss_lw": "merge(maxval(config%i_emiss_from_band_lw), size(config%lw_emiss_weights, 1), config%do_nearest_spectral_lw_emiss)",
    "n_aerosol_types": "merge(config%n_aerosol_types, 0, config%use_aerosols)",
    "n_regions": "config%nregions",
}
LOGICAL_PREFIXES = ("do_", "use_")

# ecrad_b200_inputs field -> Fortran expression (target arrays of the derived types radiation() receives)
IN_MAP = {
    "cos_sza": "c_loc(single_level%cos_sza)", "skin_temperature": "c_loc(single_level%skin_temperature)",
    "sw_albedo": "c_loc(single_level%sw_albedo)", "lw_emissivity": "c_loc(single_level%lw_emissivity)",
    "iseed": "c_loc(single_level%iseed)",
    "pressure_hl": "c_loc(thermodynamics%pressure_hl)", "temperature_hl": "c_loc(thermodynamics%temperature_hl)",
    "h2o_mmr": "c_loc(gas%mixing_ratio(1,1,IH2O))", "co2_mmr": "c_loc(gas%mixing_ratio(1,1,ICO2))", "o3_mmr": "c_loc(gas%mixing_ratio(1,1,IO3))",
    "n2o_mmr": "c_loc(gas%mixing_ratio(1,1,IN2O))", "ch4_mmr": "c_loc(gas%mixing_ratio(1,1,ICH4))",
    "cfc11_mmr": "c_loc(gas%mixing_ratio(1,1,ICFC11))", "cfc12_mmr": "c_loc(gas%mixing_ratio(1,1,ICFC12))",
    "hcfc22_mmr": "c_loc(gas%mixing_ratio(1,1,IHCFC22))", "ccl4_mmr": "c_loc(gas%mixing_ratio(1,1,ICCl4))",
    "cloud_fraction": "c_loc(cloud%fraction)", "q_liq": "c_loc(cloud%mixing_ratio(1,1,1))", "q_ice": "c_loc(cloud%mixing_ratio(1,1,2))",
    "re_liq": "c_loc(cloud%effective_radius(1,1,1))", "re_ice": "c_loc(cloud%effective_radius(1,1,2))",
    "overlap_param": "c_loc(cloud%overlap_param)", "fractional_std": "c_loc(cloud%fractional_std)",
}
IN_OPTIONAL = {   # c_null_ptr unless the condition holds
    "sw_albedo_direct": ("allocated(single_level%sw_albedo_direct)", "c_loc(single_level%sw_albedo_direct)"),
    "aerosol_mmr": ("config%use_aerosols .and. present(aerosol)", "c_loc(aerosol%mixing_ratio)"),
    "h2o_sat_liq": ("config%use_aerosols .and. allocated(thermodynamics%h2o_sat_liq)", "c_loc(thermodynamics%h2o_sat_liq)"),
    "inv_cloud_effective_size": ("allocated(cloud%inv_cloud_effective_size)", "c_loc(cloud%inv_cloud_effective_size)"),
    "inv_inhom_effective_size": ("allocated(cloud%inv_inhom_effective_size)", "c_loc(cloud%inv_inhom_effective_size)"),
}


# ----------------------------------------------------------------------------------------------------------------------
def fortran_decls(path):
    """{NAME: [extent, ...]} of the REAL/INTEGER array declarations of a reference module (PARAMETERs of the file evaluated)."""
    src = open(path, errors="replace").read()
    src = re.sub(r"&\s*\n\s*&?", "", src)
    params = {}
    # dimension parameters the modules import (ifsrrtm/parrrtm.F90, parsrtm.F90)
    for par in ("parrrtm.F90", "parsrtm.F90"):
        psrc = re.sub(r"&\s*\n\s*&?", "", open(os.path.join(os.path.dirname(path), par), errors="replace").read())
        for m in re.finditer(r"PARAMETER\s*::\s*(.*)", psrc, re.I):
            for a in m.group(1).split("!")[0].split(","):
                if "=" in a:
                    k, v = a.split("=")
                    try:
                        params[k.strip().upper()] = int(eval(v.strip().upper(), {}, params))
                    except Exception:
                        pass
    for m in re.finditer(r"PARAMETER\s*::\s*(.*)", src, re.I):
        for a in m.group(1).split(","):
            if "=" in a:
                k, v = a.split("=")
                try:
                    params[k.strip().upper()] = int(eval(v.strip(), {}, params))
                except Exception:
                    pass
    out = {}

    def extents(dims):
        ext = []
        for e in dims.split(","):
            e = e.strip()
            if ":" in e:
                lo, hi = e.split(":")
                ext.append(int(eval(hi.upper(), {}, params)) - int(eval(lo.upper(), {}, params)) + 1)
            else:
                ext.append(int(eval(e.upper(), {}, params)))
        return ext

    for m in re.finditer(r"^\s*(REAL|INTEGER)\s*\(KIND=(\w+)\)\s*((?:,\s*\w+\s*(?:\([^()]*\))?\s*)*)::\s*(.*)$", src, re.I | re.M):
        kind, attrs, rest = m.group(2).upper(), m.group(3), m.group(4)
        if re.search(r"PARAMETER", attrs, re.I):
            continue
        dattr = re.search(r"DIMENSION\s*\(([^()]*)\)", attrs, re.I)
        rest = rest.split("!")[0]
        for d in re.finditer(r"(\w+)\s*(\(([^()]*)\))?", rest):
            nm, dims = d.group(1).upper(), d.group(3)
            if nm in params:
                continue
            if dims:
                ext = extents(dims)
            elif dattr:
                ext = extents(dattr.group(1))
            else:
                ext = []
            out[nm] = (ext, kind)
    return out


def rrtmg_registrations(ref):
    """[(blob name, module, variable, section, rank, is_int)] for every RRTMG array of the blob that lives in ifsrrtm module storage."""
    blob = tables.read_blob(os.path.join(ROOT, "ecrad_b200", "data", "rrtmg_tables.bin"))
    shared = {"lw_TOTPLNK": ("yoerrtwn", "TOTPLNK"), "lw_DELWAVE": ("yoerrtwn", "DELWAVE"), "lw_NSPA": ("yoerrtwn", "NSPA"), "lw_NSPB": ("yoerrtwn", "NSPB"),
              "lw_PREFLOG": ("yoerrtrf", "PREFLOG"), "lw_TREF": ("yoerrtrf", "TREF"), "lw_CHI_MLS": ("yoerrtrf", "CHI_MLS"),
              "lw_NGB": ("yoerrtftr", "NGB"), "lw_NGC": ("yoerrtftr", "NGC"),
              "sw_PREFLOG": ("yoesrtwn", "PREFLOG"), "sw_TREF": ("yoesrtwn", "TREF"), "sw_NSPA": ("yoesrtwn", "NSPA"), "sw_NSPB": ("yoesrtwn", "NSPB"),
              "sw_NGC": ("yoesrtwn", "NGC"), "sw_NGBSW": ("yoesrtm", "NGBSW")}
    sw_rename = {"ABSA": "ABSA", "ABSB": "ABSB"}   # (the blob's sw ABSA/ABSB are KAC/KBC, EQUIVALENCEd with ABSA/ABSB in yoesrta*)
    regs, decl_cache = [], {}
    for name in sorted(blob):
        a = blob[name]
        m = re.match(r"(lw|sw)(\d+)_(\w+)$", name)
        if m:
            mod = ("yoerrta" if m.group(1) == "lw" else "yoesrta") + m.group(2)
            var = sw_rename.get(m.group(3), m.group(3))
        elif name in shared:
            mod, var = shared[name]
        else:
            continue   # cloud / aerosol / pdf / config-derived tables: registered from config_type, not from ifsrrtm modules
        if mod not in decl_cache:
            decl_cache[mod] = fortran_decls(os.path.join(ref, "ifsrrtm", mod + ".F90"))
        decls = decl_cache[mod]
        if var.upper() not in decls:
            raise SystemExit(f"{name}: no declaration of {var} in {mod}.F90")
        ext, kind = decls[var.upper()]
        want = list(a.shape)
        if not ext:                     # scalar in the module, 1-element array in the blob
            regs.append((name, mod, var, None, 0, a.dtype.kind in "iu"))
            continue
        if len(ext) != len(want):
            raise SystemExit(f"{name}: rank {len(want)} in the blob, {len(ext)} in {mod}.F90")
        sec = []
        for e, w in zip(ext, want):
            if w > e:
                raise SystemExit(f"{name}: blob extent {w} exceeds the declared {e}")
            sec.append(":" if e == w else f"1:{w}")
        regs.append((name, mod, var, "(" + ",".join(sec) + ")" if any(s != ":" for s in sec) else "", len(want), a.dtype.kind in "iu"))
    return regs


# ----------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    args = ap.parse_args()
    hdr = open(os.path.join(ROOT, "include", "ecrad_b200.h")).read()
    cfg = parse_struct(hdr, "ecrad_b200_config")
    inp = parse_struct(hdr, "ecrad_b200_inputs")
    outp = parse_struct(hdr, "ecrad_b200_outputs")
    cfg_src = open(os.path.join(args.ref, "radiation", "radiation_config.F90"), errors="replace").read().lower()

    L = []
    w = L.append
    w("! radiation_b200.F90 -- ISO_C_BINDING shim between radiation_interface and libecrad_b200.so (include/ecrad_b200.h).")
    w("! GENERATED by tools/gen_fortran_shim.py from the header's structs and the table blob's name list: do not edit by hand.")
    w("!")
    w("! Call sites (two lines in radiation/radiation_interface.F90, see INTEGRATION.md):")
    w("!   end of setup_radiation (:153)            call b200_setup(config)")
    w("!   start of the else branch of radiation (:318)   if (b200_active) then; call b200_radiation(<the arguments of radiation>); else <existing body> end if")
    w("! The caller's arrays are handed over by pointer, nothing is copied on the host: double-precision builds (JPRB = JPRD) call")
    w("! ecrad_b200_radiation, single-precision builds (-DPARKIND1_SINGLE, ifsaux/parkind1.F90:46-50) call ecrad_b200_radiation_sp, which takes")
    w("! the same structs with float arrays behind the pointers (widened on the device; the kernels compute in double precision).")
    w("module radiation_b200")
    w("  use, intrinsic :: iso_c_binding")
    w("  use parkind1, only : jprb, jpim")
    w("  implicit none")
    w("  private")
    w("  public :: b200_setup, b200_radiation, b200_save_radiative_properties, b200_finalize, b200_set_option, b200_active")
    w("")
    w("  logical     :: b200_active = .false.")
    w("  type(c_ptr) :: handle = c_null_ptr")
    w("")
    # ---- structs
    w("  type, bind(c) :: cfg_t                       ! == struct ecrad_b200_config, field by field")
    for t, f in cfg:
        w(f"    {'integer(c_int32_t)' if t == 'int32_t' else 'real(c_double)    '} :: {f}")
    w("  end type")
    w("  type, bind(c) :: in_t                        ! == struct ecrad_b200_inputs")
    for t, f in inp:
        ft = {"int32_t": "integer(c_int32_t)", "double": "real(c_double)    "}.get(t, "type(c_ptr)       ")
        w(f"    {ft} :: {f}")
    w("  end type")
    w("  type, bind(c) :: out_t                       ! == struct ecrad_b200_outputs: %d pointers in the header's order, c_null_ptr = not allocated" % (len(outp) - 2))
    w("    integer(c_int32_t) :: struct_bytes, reserved")
    w(f"    type(c_ptr)        :: p({len(outp) - 2})")
    w("  end type")
    w("  type, bind(c) :: props_t                     ! == struct ecrad_b200_radiative_properties: 18 pointers in the header's order")
    w("    type(c_ptr) :: p(18)")
    w("  end type")
    w("")
    w("  interface")
    w("    function ecrad_b200_tables_create() bind(c) result(t)")
    w("      import; type(c_ptr) :: t")
    w("    end function")
    w("    function ecrad_b200_tables_add(t, name, dtype, ndim, dims, data) bind(c) result(rc)")
    w("      import; type(c_ptr), value :: t, data; character(kind=c_char) :: name(*)")
    w("      integer(c_int), value :: dtype, ndim; integer(c_int64_t) :: dims(*); integer(c_int) :: rc")
    w("    end function")
    w("    subroutine ecrad_b200_tables_free(t) bind(c)")
    w("      import; type(c_ptr), value :: t")
    w("    end subroutine")
    w("    function ecrad_b200_setup(cfg, tab, h) bind(c) result(rc)")
    w("      import; type(cfg_t) :: cfg; type(c_ptr), value :: tab; type(c_ptr) :: h; integer(c_int) :: rc")
    w("    end function")
    w("    function ecrad_b200_radiation(h, ncol, nlev, istartcol, iendcol, inp, outp) bind(c) result(rc)")
    w("      import; type(c_ptr), value :: h; integer(c_int), value :: ncol, nlev, istartcol, iendcol")
    w("      type(in_t) :: inp; type(out_t) :: outp; integer(c_int) :: rc")
    w("    end function")
    w("    function ecrad_b200_radiation_sp(h, ncol, nlev, istartcol, iendcol, inp, outp) bind(c) result(rc)")
    w("      import; type(c_ptr), value :: h; integer(c_int), value :: ncol, nlev, istartcol, iendcol")
    w("      type(in_t) :: inp; type(out_t) :: outp; integer(c_int) :: rc")
    w("    end function")
    w("    function ecrad_b200_save_radiative_properties(h, ncol, nlev, istartcol, iendcol, inp, props) bind(c) result(rc)")
    w("      import; type(c_ptr), value :: h; integer(c_int), value :: ncol, nlev, istartcol, iendcol")
    w("      type(in_t) :: inp; type(props_t) :: props; integer(c_int) :: rc")
    w("    end function")
    w("    function ecrad_b200_set_solar_cycle_multiplier(h, multiplier) bind(c) result(rc)")
    w("      import; type(c_ptr), value :: h; real(c_double), value :: multiplier; integer(c_int) :: rc")
    w("    end function")
    w("    function ecrad_b200_set_option(h, key, val) bind(c) result(rc)")
    w("      import; type(c_ptr), value :: h; character(kind=c_char) :: key(*); integer(c_int), value :: val; integer(c_int) :: rc")
    w("    end function")
    w("    subroutine ecrad_b200_finalize(h) bind(c)")
    w("      import; type(c_ptr), value :: h")
    w("    end subroutine")
    w("    function ecrad_b200_last_error(h) bind(c) result(msg)")
    w("      import; type(c_ptr), value :: h; type(c_ptr) :: msg")
    w("    end function")
    w("  end interface")
    w("")
    w("contains")
    w("")
    w("  ! ---- table registration helpers: the directory copies the data, so sections and temporaries are fine; everything is")
    w("  !      handed over as real(c_double) / integer(c_int32_t) in Fortran element order with its Fortran dimensions")
    for rank in (1, 2, 3, 4):
        dims = ",".join(":" * 1 for _ in range(rank))
        w(f"  subroutine add_r{rank}(t, name, a)")
        w(f"    type(c_ptr), intent(in) :: t; character(*), intent(in) :: name; real(jprb), intent(in) :: a({dims})")
        w(f"    real(c_double), allocatable, target :: tmp({dims}); integer(c_int) :: rc")
        w("    allocate(tmp, source=real(a, c_double))")
        w(f"    rc = ecrad_b200_tables_add(t, name//c_null_char, 0_c_int, {rank}_c_int, int(shape(tmp), c_int64_t), c_loc(tmp))")
        w("    if (rc /= 0) call abort_with(ecrad_b200_last_error(c_null_ptr))")
        w("  end subroutine")
    w("  subroutine add_i1(t, name, a)")
    w("    type(c_ptr), intent(in) :: t; character(*), intent(in) :: name; integer(jpim), intent(in) :: a(:)")
    w("    integer(c_int32_t), allocatable, target :: tmp(:); integer(c_int) :: rc")
    w("    allocate(tmp, source=int(a, c_int32_t))")
    w("    rc = ecrad_b200_tables_add(t, name//c_null_char, 1_c_int, 1_c_int, int(shape(tmp), c_int64_t), c_loc(tmp))")
    w("    if (rc /= 0) call abort_with(ecrad_b200_last_error(c_null_ptr))")
    w("  end subroutine")
    w("")
    # ---- RRTMG tables
    regs = rrtmg_registrations(args.ref)
    w("  ! ---- RRTMG-IFS: what RRTM_INIT_140GP / SRTM_INIT left in module storage (%d arrays; names = those of the ETB1 blob)" % len(regs))
    w("  subroutine register_rrtmg(t)")
    mods = {}
    for name, mod, var, sec, rank, is_int in regs:
        mods.setdefault(mod, []).append((name, var))
    for mod in sorted(mods):
        ren = ", ".join(f"{mod}_{var} => {var}" for var in sorted({v for _, v in mods[mod]}))
        w(f"    use {mod}, only : {ren}")
    w("    type(c_ptr), intent(in) :: t")
    for name, mod, var, sec, rank, is_int in regs:
        v = f"{mod}_{var}"
        if rank == 0:
            w(f"    call add_{'i' if is_int else 'r'}1(t, '{name}', [{v}])")
        elif is_int:
            w(f"    call add_i1(t, '{name}', {v}{sec})")
        else:
            w(f"    call add_r{rank}(t, '{name}', {v}{sec})")
    w("  end subroutine")
    w("")
    w("  ! ---- tables that live in config_type: cloud optics coefficients, McICA PDF look-up table, spectral mappings, aerosol optics")
    w("  subroutine register_config_tables(t, config)")
    w("    use radiation_config, only : config_type")
    w("    type(c_ptr), intent(in) :: t; type(config_type), intent(in) :: config")
    w("    integer :: j")
    w("    if (.not. config%use_general_cloud_optics) then   ! setup_cloud_optics ran (radiation_interface.F90:116-120)")
    w("      call add_r2(t, 'liq_coeff_lw', config%cloud_optics%liq_coeff_lw); call add_r2(t, 'liq_coeff_sw', config%cloud_optics%liq_coeff_sw)")
    w("      call add_r2(t, 'ice_coeff_lw', config%cloud_optics%ice_coeff_lw); call add_r2(t, 'ice_coeff_sw', config%cloud_optics%ice_coeff_sw)")
    w("      ! (the arrays of the configured liquid_model_name / ice_model_name; the library checks their coefficient counts against i_liq_model / i_ice_model)")
    w("      if (allocated(config%cloud_optics%ice_coeff_gen)) call add_r1(t, 'ice_coeff_gen', config%cloud_optics%ice_coeff_gen)   ! Baran-2017")
    w("    end if")
    w("    call add_r2(t, 'pdf_val', config%pdf_sampler%val)                 ! val(ncdf, nfsd), radiation_pdf_sampler.F90:83-93")
    w("    call add_r1(t, 'pdf_fsd', [(config%pdf_sampler%fsd1 + real(j-1,jprb) / config%pdf_sampler%inv_fsd_interval, j = 1, config%pdf_sampler%nfsd)])")
    w("    call add_r2(t, 'sw_albedo_weights', config%sw_albedo_weights)     ! (n_albedo_sw, n_bands_sw)")
    w("    if (allocated(config%i_albedo_from_band_sw)) call add_i1(t, 'i_albedo_from_band_sw', config%i_albedo_from_band_sw)")
    w("    if (allocated(config%i_emiss_from_band_lw)) call add_i1(t, 'i_emiss_from_band_lw', config%i_emiss_from_band_lw)")
    w("    ! g-point order of the solvers (radiation_ifs_rrtm.F90:122-130, :167-174): a permutation for SPARTACUS, the identity otherwise")
    w("    if (allocated(config%i_g_from_reordered_g_sw)) call add_i1(t, 'i_g_from_reordered_g_sw', config%i_g_from_reordered_g_sw)")
    w("    if (allocated(config%i_g_from_reordered_g_lw)) call add_i1(t, 'i_g_from_reordered_g_lw', config%i_g_from_reordered_g_lw)")
    w("    if (allocated(config%lw_emiss_weights)) call add_r2(t, 'lw_emiss_weights', config%lw_emiss_weights)")
    w("    if (config%use_aerosols) then                                      ! config%aerosol_optics (radiation_aerosol_optics_data.F90:35-120)")
    w("      call add_i1(t, 'aerosol_iclass', config%aerosol_optics%iclass); call add_i1(t, 'aerosol_itype', config%aerosol_optics%itype)")
    w("      call add_r1(t, 'aer_rh_lower', config%aerosol_optics%rh_lower)")
    for sp in ("sw", "lw"):
        for q, comp in (("mass_ext", "mass_ext"), ("ssa", "ssa"), ("g", "g")):
            w(f"      call add_r2(t, 'aer_{q}_{sp}_phobic', config%aerosol_optics%{comp}_{sp}_phobic)")
            w(f"      call add_r3(t, 'aer_{q}_{sp}_philic', config%aerosol_optics%{comp}_{sp}_philic)")
    w("    end if")
    w("  end subroutine")
    w("")
    w("  ! ---- ecCKD: config%gas_optics_lw/sw (ckd_model_type, radiation_ecckd.F90:60-126), registered per spectrum in b200_setup, and the")
    w("  !      generalised cloud optics config%cloud_optics_lw/sw(1:2) (general_cloud_optics_type, radiation_general_cloud_optics_data.F90:30-68)")
    w("  subroutine register_ckd_model(t, pre, go, is_sw)")
    w("    use radiation_ecckd, only : ckd_model_type")
    w("    use radiation_ecckd_gas, only : IConcDependenceLUT")
    w("    type(c_ptr), intent(in) :: t; character(*), intent(in) :: pre; type(ckd_model_type), intent(in) :: go; logical, intent(in) :: is_sw")
    w("    real(jprb) :: gas_meta(6, go%ngas); integer :: j; character(len=8) :: cj")
    w("    call add_r1(t, pre//'meta', [real(go%ng, jprb), real(go%npress, jprb), real(go%ntemp, jprb), real(go%nplanck, jprb), real(go%ngas, jprb), &")
    w("         &  go%log_pressure1, go%d_log_pressure, go%d_temperature, go%temperature1_planck, go%d_temperature_planck, merge(1.0_jprb, 0.0_jprb, is_sw)])")
    w("    call add_r1(t, pre//'temperature1', go%temperature1)")
    w("    if (is_sw) then")
    w("      call add_r1(t, pre//'norm_solar_irradiance', go%norm_solar_irradiance); call add_r1(t, pre//'rayleigh_molar_scat', go%rayleigh_molar_scat)")
    w("      ! read_spectral_solar_cycle ran (config%use_spectral_solar_cycle, radiation_ecckd_interface.F90:79-82)")
    w("      if (allocated(go%norm_amplitude_solar_irradiance)) call add_r1(t, pre//'norm_amplitude_solar_irradiance', go%norm_amplitude_solar_irradiance)")
    w("    else")
    w("      call add_r2(t, pre//'planck_function', go%planck_function)       ! (ng, nplanck)")
    w("    end if")
    w("    do j = 1, go%ngas")
    w("      write(cj, '(i0)') j - 1")
    w("      associate (sg => go%single_gas(j))")
    w("        gas_meta(:, j) = [real(sg%i_gas_code, jprb), real(sg%i_conc_dependence, jprb), sg%reference_mole_frac, real(sg%n_mole_frac, jprb), &")
    w("             &           sg%log_mole_frac1, sg%d_log_mole_frac]")
    w("        if (sg%i_conc_dependence == IConcDependenceLUT) then")
    w("          call add_r4(t, pre//'gas'//trim(cj)//'_molar_abs', sg%molar_abs_conc)   ! (ng, npress, ntemp, nconc)")
    w("        else")
    w("          call add_r3(t, pre//'gas'//trim(cj)//'_molar_abs', sg%molar_abs)        ! (ng, npress, ntemp)")
    w("        end if")
    w("      end associate")
    w("    end do")
    w("    call add_r2(t, pre//'gas_meta', gas_meta)")
    w("  end subroutine")
    w("  subroutine register_gco(t, pre, co)")
    w("    use radiation_general_cloud_optics_data, only : general_cloud_optics_type")
    w("    type(c_ptr), intent(in) :: t; character(*), intent(in) :: pre; type(general_cloud_optics_type), intent(in) :: co")
    w("    call add_r1(t, pre//'meta', [real(co%n_effective_radius, jprb), co%effective_radius_0, co%d_effective_radius])")
    w("    call add_r2(t, pre//'mass_ext', co%mass_ext); call add_r2(t, pre//'ssa', co%ssa); call add_r2(t, pre//'asymmetry', co%asymmetry)   ! (ng, nre)")
    w("  end subroutine")
    w("")
    # ---- setup
    w("  subroutine b200_setup(config)")
    w("    use radiation_config, only : config_type, IGasModelECCKD")
    w("    type(config_type), intent(in) :: config")
    w("    type(cfg_t) :: c; type(c_ptr) :: t")
    w("    t = ecrad_b200_tables_create()")
    w("    ! one gas model per spectrum (radiation_interface.F90:333-355): the ifsrrtm module storage as soon as one spectrum runs RRTMG-IFS,")
    w("    ! the ckd_model_type of each ecCKD spectrum; the library tells the spectra apart by which tables it finds")
    w("    if (config%i_gas_model_lw /= IGasModelECCKD .or. config%i_gas_model_sw /= IGasModelECCKD) call register_rrtmg(t)")
    w("    if (config%i_gas_model_lw == IGasModelECCKD) call register_ckd_model(t, 'ckd_lw_', config%gas_optics_lw, .false.)")
    w("    if (config%i_gas_model_sw == IGasModelECCKD) call register_ckd_model(t, 'ckd_sw_', config%gas_optics_sw, .true.)")
    w("    if (config%use_general_cloud_optics) then   ! look-up tables per g-point (ecCKD) or per RRTMG band (radiation_config.F90:1078-1090)")
    w("      call register_gco(t, 'gco_lw_0_', config%cloud_optics_lw(1)); call register_gco(t, 'gco_lw_1_', config%cloud_optics_lw(2))")
    w("      call register_gco(t, 'gco_sw_0_', config%cloud_optics_sw(1)); call register_gco(t, 'gco_sw_1_', config%cloud_optics_sw(2))")
    w("    end if")
    w("    call register_config_tables(t, config)")
    missing = []
    for tname, f in cfg:
        if f in CFG_SPECIAL:
            w(f"    c%{f} = {CFG_SPECIAL[f]}")
            continue
        if not re.search(r"\b%s\b" % re.escape(f), cfg_src):
            missing.append(f)
        if tname == "int32_t" and f.startswith(LOGICAL_PREFIXES):
            w(f"    c%{f} = merge(1_c_int32_t, 0_c_int32_t, config%{f})")
        elif tname == "int32_t":
            w(f"    c%{f} = int(config%{f}, c_int32_t)")
        else:
            w(f"    c%{f} = real(config%{f}, c_double)")
    if missing:
        raise SystemExit("ecrad_b200_config fields without a component of config_type: " + ", ".join(missing))
    w("    if (ecrad_b200_setup(c, t, handle) /= 0) call abort_with(ecrad_b200_last_error(c_null_ptr))")
    w("    call ecrad_b200_tables_free(t)")
    w("    b200_active = .true.")
    w("  end subroutine")
    w("")
    # ---- radiation
    w("  subroutine b200_radiation(ncol, nlev, istartcol, iendcol, config, single_level, thermodynamics, gas, cloud, flux, aerosol)")
    w("    use radiation_config, only : config_type")
    w("    use radiation_single_level, only : single_level_type")
    w("    use radiation_thermodynamics, only : thermodynamics_type")
    w("    use radiation_gas, only : gas_type, IH2O, ICO2, IO3, IN2O, ICH4, ICFC11, ICFC12, IHCFC22, ICCl4")
    w("    use radiation_cloud, only : cloud_type")
    w("    use radiation_aerosol, only : aerosol_type")
    w("    use radiation_flux, only : flux_type")
    w("    integer, intent(in) :: ncol, nlev, istartcol, iendcol")
    w("    type(config_type), intent(in) :: config")
    w("    type(single_level_type), intent(in), target :: single_level")
    w("    type(thermodynamics_type), intent(in), target :: thermodynamics")
    w("    type(gas_type), intent(in), target :: gas")
    w("    type(cloud_type), intent(inout), target :: cloud")
    w("    type(flux_type), intent(inout), target :: flux")
    w("    type(aerosol_type), intent(in), target, optional :: aerosol")
    w("    type(in_t) :: i; type(out_t) :: o")
    w("    i%struct_bytes = int(c_sizeof(i), c_int32_t); i%reserved = 0")
    w("    o%struct_bytes = int(c_sizeof(o), c_int32_t); o%reserved = 0; o%p = c_null_ptr")
    in_lines = ["    i%struct_bytes = int(c_sizeof(i), c_int32_t); i%reserved = 0",
                "    i%solar_irradiance = real(single_level%solar_irradiance, c_double)"]
    unmapped = []
    for tname, f in inp:
        if f in ("struct_bytes", "reserved", "solar_irradiance"):
            continue
        if f in IN_MAP:
            in_lines.append(f"    i%{f} = {IN_MAP[f]}")
        elif f in IN_OPTIONAL:
            cond, expr = IN_OPTIONAL[f]
            in_lines.append(f"    i%{f} = c_null_ptr")
            if "present(aerosol)" in cond:
                in_lines.append("    if (present(aerosol)) then")
                in_lines.append(f"      if (config%use_aerosols) i%{f} = {expr}")
                in_lines.append("    end if")
            else:
                in_lines.append(f"    if ({cond}) i%{f} = {expr}")
        else:
            unmapped.append(f)
    if unmapped:
        raise SystemExit("ecrad_b200_inputs fields without a mapping: " + ", ".join(unmapped))
    for ln in in_lines[1:]:
        w(ln)
    k = 0
    for tname, f in outp:
        if f in ("struct_bytes", "reserved"):
            continue
        k += 1
        w(f"    if (allocated(flux%{f})) o%p({k}) = c_loc(flux%{f})")
    w("    if (config%use_spectral_solar_cycle) then   ! calc_incoming_sw with the multiplier, radiation_ecckd_interface.F90:284-290")
    w("      if (ecrad_b200_set_solar_cycle_multiplier(handle, real(single_level%spectral_solar_cycle_multiplier, c_double)) /= 0) &")
    w("           &  call abort_with(ecrad_b200_last_error(handle))")
    w("    end if")
    w("#ifdef PARKIND1_SINGLE")
    w("    if (ecrad_b200_radiation_sp(handle, int(ncol, c_int), int(nlev, c_int), int(istartcol, c_int), int(iendcol, c_int), i, o) /= 0) &")
    w("         &  call abort_with(ecrad_b200_last_error(handle))")
    w("#else")
    w("    if (ecrad_b200_radiation(handle, int(ncol, c_int), int(nlev, c_int), int(istartcol, c_int), int(iendcol, c_int), i, o) /= 0) &")
    w("         &  call abort_with(ecrad_b200_last_error(handle))")
    w("#endif")
    w("  end subroutine")
    w("")
    props = ["planck_hl", "lw_emission", "lw_albedo", "sw_albedo_direct", "sw_albedo_diffuse", "incoming_sw", "od_lw", "ssa_lw", "g_lw",
             "od_sw", "ssa_sw", "g_sw", "od_lw_cloud", "ssa_lw_cloud", "g_lw_cloud", "od_sw_cloud", "ssa_sw_cloud", "g_sw_cloud"]
    hdr_props = re.search(r"typedef struct ecrad_b200_radiative_properties \{(.*?)\} ecrad_b200_radiative_properties;", hdr, re.S).group(1)
    hdr_props = re.findall(r"double\*\s*(\w+);", re.sub(r"/\*.*?\*/", "", hdr_props, flags=re.S))
    if hdr_props != props:
        raise SystemExit("ecrad_b200_radiative_properties changed: " + ", ".join(hdr_props))
    w("  ! The arrays radiation() passes to save_radiative_properties (radiation_interface.F90:405-425), filled by the library from the same")
    w("  ! inputs; call it where config%do_save_radiative_properties is tested and hand the arrays on to radiation_save.  Double precision only.")
    w("  subroutine b200_save_radiative_properties(ncol, nlev, istartcol, iendcol, config, single_level, thermodynamics, gas, cloud, &")
    w("       &  planck_hl, lw_emission, lw_albedo, sw_albedo_direct, sw_albedo_diffuse, incoming_sw, od_lw, ssa_lw, g_lw, od_sw, ssa_sw, g_sw, &")
    w("       &  od_lw_cloud, ssa_lw_cloud, g_lw_cloud, od_sw_cloud, ssa_sw_cloud, g_sw_cloud, aerosol)")
    w("    use radiation_config, only : config_type")
    w("    use radiation_single_level, only : single_level_type")
    w("    use radiation_thermodynamics, only : thermodynamics_type")
    w("    use radiation_gas, only : gas_type, IH2O, ICO2, IO3, IN2O, ICH4, ICFC11, ICFC12, IHCFC22, ICCl4")
    w("    use radiation_cloud, only : cloud_type")
    w("    use radiation_aerosol, only : aerosol_type")
    w("    integer, intent(in) :: ncol, nlev, istartcol, iendcol")
    w("    type(config_type), intent(in) :: config")
    w("    type(single_level_type), intent(in), target :: single_level")
    w("    type(thermodynamics_type), intent(in), target :: thermodynamics")
    w("    type(gas_type), intent(in), target :: gas")
    w("    type(cloud_type), intent(in), target :: cloud")
    w("    ! (spectral index, level, column) with ALL ncol columns, like the library's other arrays; columns istartcol:iendcol are written")
    w("    real(jprb), intent(inout), target, contiguous, dimension(:,:,:) :: planck_hl, od_lw, ssa_lw, g_lw, od_sw, ssa_sw, g_sw, &")
    w("         &  od_lw_cloud, ssa_lw_cloud, g_lw_cloud, od_sw_cloud, ssa_sw_cloud, g_sw_cloud")
    w("    real(jprb), intent(inout), target, contiguous, dimension(:,:) :: lw_emission, lw_albedo, sw_albedo_direct, sw_albedo_diffuse, incoming_sw")
    w("    type(aerosol_type), intent(in), target, optional :: aerosol")
    w("    type(in_t) :: i; type(props_t) :: p")
    for ln in in_lines:
        w(ln)
    for k, nm in enumerate(props, 1):
        w(f"    p%p({k}) = c_loc({nm})")
    w("    if (ecrad_b200_save_radiative_properties(handle, int(ncol, c_int), int(nlev, c_int), int(istartcol, c_int), int(iendcol, c_int), i, p) /= 0) &")
    w("         &  call abort_with(ecrad_b200_last_error(handle))")
    w("  end subroutine")
    w("")
    w("  subroutine b200_set_option(key, val)          ! e.g. call b200_set_option('register_host', 1): page-lock the caller's arrays once")
    w("    character(*), intent(in) :: key; integer, intent(in) :: val")
    w("    if (ecrad_b200_set_option(handle, key//c_null_char, int(val, c_int)) /= 0) call abort_with(ecrad_b200_last_error(handle))")
    w("  end subroutine")
    w("")
    w("  subroutine b200_finalize()")
    w("    call ecrad_b200_finalize(handle); handle = c_null_ptr; b200_active = .false.")
    w("  end subroutine")
    w("")
    w("  subroutine abort_with(cmsg)                   ! reference error convention: message on nulerr, then radiation_abort")
    w("    use radiation_io, only : nulerr, radiation_abort")
    w("    type(c_ptr), intent(in) :: cmsg")
    w("    character(kind=c_char), pointer :: s(:); integer :: n")
    w("    call c_f_pointer(cmsg, s, [512])")
    w("    n = 0")
    w("    do while (n < 512)")
    w("      if (s(n+1) == c_null_char) exit")
    w("      n = n + 1")
    w("    end do")
    w("    write(nulerr, '(a,512a1)') '*** Error (ecrad_b200): ', s(1:n)")
    w("    call radiation_abort()")
    w("  end subroutine")
    w("")
    w("end module radiation_b200")
    os.makedirs(os.path.join(ROOT, "fortran"), exist_ok=True)
    path = os.path.join(ROOT, "fortran", "radiation_b200.F90")
    open(path, "w").write("\n".join(L) + "\n")
    print(f"wrote {path}: {len(L)} lines, {len(regs)} RRTMG tables, {len(cfg)} config fields, {len(inp) - 3} input arrays, {len(outp) - 2} outputs")


if __name__ == "__main__":
    main()
