"""Liquid / ice optics parameterisations other than the default pair (config%i_liq_model, i_ice_model): Slingo + Lindner-Li liquid;
Baran, Baran-2016, Baran-2017 and Yi ice (radiation_cloud_optics.F90:345-447 dispatches exactly these).

The reference holds no output for any of them, so the pins are: (1) a vectorised numpy restatement of each published formula, written
here from the Fortran independently of oracle/cloud.c, against the oracle's orc_cloud_optics on the reference's own test columns;
(2) ecrad_b200/csrc/cloud_core.h (what the GPU kernel compiles) replayed on the CPU against the oracle, bit for bit;
(3) the fluxes of a full oracle run stay physical and close to the default pair's.  The GPU parity test is in test_gpu_parity.py.
"""
import ctypes as C
import os

import numpy as np
import pytest

import hostcheck_lib
from ecrad_b200 import abi, tables
from ecrad_b200 import inputs as I
from ecrad_b200.config import RadiationConfig
from oracle_lib import Oracle

NLEV = 137
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODELS = [("Slingo", "Fu-IFS"), ("SOCRATES", "Baran-EXPERIMENTAL"), ("SOCRATES", "Baran2016"), ("SOCRATES", "Baran2017-EXPERIMENTAL"),
          ("SOCRATES", "Yi"), ("Slingo", "Yi")]
LIQ_TAG = {"socrates": "", "slingo": ".slingo"}
ICE_TAG = {"fu-ifs": "", "baran-experimental": ".baran", "baran2016": ".baran2016", "baran2017-experimental": ".baran2017", "yi": ".yi"}


@pytest.fixture(scope="module")
def blob():
    return tables.read_blob(os.path.join(ROOT, "ecrad_b200", "data", "rrtmg_tables.bin"))


def delta_eddington(od, scat, g):
    f = g * g
    return od - scat * f, scat * (1.0 - f), g / (1.0 + g)


def numpy_cloud_optics(blob, liq, ice, p_hl, t_hl, frac, q_liq, q_ice, re_liq, re_ice):
    """(od, ssa, g) x (lw, sw), [nlev][nb]; do_lw_cloud_scattering on, delta-Eddington applied to the particles (defaults)."""
    out = {}
    for sp, nb in (("lw", 16), ("sw", 14)):
        cl, ci = blob[f"liq_coeff_{sp}{LIQ_TAG[liq]}"], blob[f"ice_coeff_{sp}{ICE_TAG[ice]}"]   # (nb, ncoeff)
        od_o, ssa_o, g_o = np.zeros((NLEV, nb)), np.zeros((NLEV, nb)), np.zeros((NLEV, nb))
        for l in range(NLEV):
            if not frac[l] > 0.0:
                continue
            factor = (p_hl[l + 1] - p_hl[l]) / (9.80665 * frac[l])
            lwp, iwp = factor * q_liq[l], factor * q_ice[l]
            odl = scl = gl = odi = sci = gi = np.zeros(nb)
            if lwp > 0.0:
                if liq == "slingo" and sp == "sw":
                    re_um = min(max(4.2, re_liq[l] * 1.0e6), 16.6)
                    odl = lwp * 1000.0 * (cl[:, 0] + cl[:, 1] / re_um)
                    scl = odl * (1.0 - cl[:, 2] - re_um * cl[:, 3])
                    gl = cl[:, 4] + re_um * cl[:, 5]
                elif liq == "slingo":
                    r = min(max(2.0, re_liq[l] * 1.0e6), 40.0)
                    ir = 1.0 / r
                    odl = lwp * 1000.0 * (cl[:, 0] + r * cl[:, 1] + ir * (cl[:, 2] + ir * (cl[:, 3] + ir * cl[:, 4])))
                    scl = odl * (1.0 - (cl[:, 5] + ir * cl[:, 6] + r * (cl[:, 7] + r * cl[:, 8])))
                    gl = cl[:, 9] + ir * cl[:, 10] + r * (cl[:, 11] + r * cl[:, 12])
                else:
                    r = max(float(np.float32(1.2e-6)), min(re_liq[l], float(np.float32(50.0e-6))))
                    odl = lwp * (cl[:, 0] + r * (cl[:, 1] + r * cl[:, 2])) / (1.0 + r * (cl[:, 3] + r * (cl[:, 4] + r * cl[:, 5])))
                    scl = odl * (1.0 - (cl[:, 6] + r * (cl[:, 7] + r * cl[:, 8])) / (1.0 + r * (cl[:, 9] + r * cl[:, 10])))
                    gl = (cl[:, 11] + r * (cl[:, 12] + r * cl[:, 13])) / (1.0 + r * (cl[:, 14] + r * cl[:, 15]))
                if sp == "sw":
                    odl, scl, gl = delta_eddington(odl, scl, gl)
            if iwp > 0.0:
                qi, T = q_ice[l], 0.5 * (t_hl[l] + t_hl[l + 1])
                if ice == "baran-experimental":
                    odi = iwp * (ci[:, 0] + ci[:, 1] / (1.0 + qi * ci[:, 2]))
                    sci = odi * (ci[:, 3] + ci[:, 4] / (1.0 + qi * ci[:, 5]))
                    gi = ci[:, 6] + ci[:, 7] / (1.0 + qi * ci[:, 8])
                elif ice == "baran2016":
                    qi_T = min(qi, 1.0e-3) * T
                    odi = iwp * ci[:, 0] * (1.0 / ((T * T) * (T * T)))
                    sci = odi * (ci[:, 1] + ci[:, 2] * qi_T)
                    gi = ci[:, 3] + ci[:, 4] * qi_T
                elif ice == "baran2017-experimental":
                    cg = blob["ice_coeff_gen.baran2017"]
                    qm = qi * np.exp(cg[0] * (T - cg[1]))
                    odi = iwp * (ci[:, 0] + ci[:, 1] / (1.0 + qm ** cg[2] * ci[:, 2]))
                    sci = odi * (ci[:, 3] + ci[:, 4] / (1.0 + qm ** cg[3] * ci[:, 5]))
                    gi = ci[:, 6] + ci[:, 7] / (1.0 + qm ** cg[4] * ci[:, 8])
                elif ice == "yi":
                    de = min(max(re_ice[l] * 2.0e6, 10.0), 119.99)
                    x = de * 0.2 - 1.0
                    i0 = int(np.floor(x))          # 1-based index of the reference = 0-based column i0 - 1
                    w2 = x - i0
                    w1 = 1.0 - w2
                    odi = 0.001 * (iwp * 1000.0) * (w1 * ci[:, i0 - 1] + w2 * ci[:, i0])
                    sci = odi * (w1 * ci[:, i0 - 1 + 23] + w2 * ci[:, i0 + 23])
                    gi = w1 * ci[:, i0 - 1 + 46] + w2 * ci[:, i0 + 46]
                else:   # Fu (Fu 1996 shortwave, Fu et al. 1998 longwave); do_fu_lw_ice_optics_bug off
                    de = min(re_ice[l], 100.0e-6) * (1.0e6 / 0.64952)
                    ide, w = 1.0 / de, iwp * 1000.0
                    maxg = 1.0 - 10.0 * np.finfo(np.float64).eps
                    if sp == "sw":
                        odi = w * (ci[:, 0] + ci[:, 1] * ide)
                        sci = odi * (1.0 - (ci[:, 2] + de * (ci[:, 3] + de * (ci[:, 4] + de * ci[:, 5]))))
                        gi = np.minimum(ci[:, 6] + de * (ci[:, 7] + de * (ci[:, 8] + de * ci[:, 9])), maxg)
                    else:
                        odi = w * (ci[:, 0] + ide * (ci[:, 1] + ide * ci[:, 2]))
                        sci = odi - w * ide * (ci[:, 3] + de * (ci[:, 4] + de * (ci[:, 5] + de * ci[:, 6])))
                        gi = np.minimum(ci[:, 7] + de * (ci[:, 8] + de * (ci[:, 9] + de * ci[:, 10])), maxg)
                odi, sci, gi = delta_eddington(odi, sci, gi)
            od_o[l] = odl + odi
            with np.errstate(invalid="ignore", divide="ignore"):
                g_o[l] = np.where(scl + sci > 0.0, (gl * scl + gi * sci) / (scl + sci), 0.0) if sp == "lw" else (gl * scl + gi * sci) / (scl + sci)
                ssa_o[l] = (scl + sci) / (odl + odi)
        out[sp] = (od_o, ssa_o, g_o)
    return out


def column_inputs(raw, cfg, c):
    inp = I.to_radiation_inputs(raw, cfg)
    f = lambda nm: np.ascontiguousarray(inp[nm][c], dtype=np.float64)   # noqa: E731
    frac = f("cloud_fraction").copy()
    # crop_cloud_fraction (radiation_cloud.F90:700-740), which the full run applies before the cloud optics
    frac[(frac < cfg.cloud_fraction_threshold) | (f("q_liq") + f("q_ice") < cfg.cloud_mixing_ratio_threshold)] = 0.0
    return f("pressure_hl"), f("temperature_hl"), frac, f("q_liq"), f("q_ice"), f("re_liq"), f("re_ice")


def oracle_cloud_optics(orc, cols):
    L = orc.lib
    dp = C.POINTER(C.c_double)
    L.orc_cloud_optics.argtypes = [C.c_void_p, C.POINTER(abi.Config), C.c_int] + [dp] * 13
    L.orc_cloud_optics.restype = C.c_int
    arrs = [np.zeros((NLEV, 16)) for _ in range(3)] + [np.zeros((NLEV, 14)) for _ in range(3)]
    P = lambda a: a.ctypes.data_as(dp)   # noqa: E731
    rc = L.orc_cloud_optics(orc.t, C.byref(orc.cfg), NLEV, *[P(a) for a in cols], *[P(a) for a in arrs])
    assert rc == 0
    return {"lw": tuple(arrs[:3]), "sw": tuple(arrs[3:])}


@pytest.mark.parametrize("liq,ice", MODELS + [("SOCRATES", "Fu-IFS")])
def test_oracle_matches_numpy_restatement(meridian_raw, blob, liq, ice):
    cfg = RadiationConfig(liquid_model_name=liq, ice_model_name=ice).consolidate()
    orc = Oracle(cfg)
    ncloudy = 0
    for c in (3, 9, 12, 17, 24, 31):
        cols = column_inputs(meridian_raw, cfg, c)
        got = oracle_cloud_optics(orc, cols)
        ref = numpy_cloud_optics(blob, liq.lower(), ice.lower(), *cols)
        ncloudy += int((cols[2] > 0.0).sum())
        for sp in ("lw", "sw"):
            for k, nm in enumerate(("od", "ssa", "g")):
                a, b = got[sp][k], ref[sp][k]
                assert np.allclose(a, b, rtol=1e-12, atol=1e-300, equal_nan=True), (sp, nm, c, float(np.nanmax(np.abs(a - b))))
    assert ncloudy > 50


@pytest.mark.parametrize("liq,ice", MODELS)
def test_device_core_replay_is_bit_identical_to_oracle(meridian_raw, liq, ice):
    """cloud_core.h (the header the CUDA kernel compiles) on the CPU, against oracle/cloud.c: same operations, same bits."""
    cfg = RadiationConfig(liquid_model_name=liq, ice_model_name=ice).consolidate()
    orc = Oracle(cfg)
    lib, _ = hostcheck_lib.load()
    dp = C.POINTER(C.c_double)
    lib.hc_load_models.restype = C.c_void_p
    lib.hc_load_models.argtypes = [C.c_char_p, C.c_int, C.c_int]
    lib.hc_cloud_optics.argtypes = [C.c_void_p, C.c_int] + [dp] * 7 + [C.c_int] * 3 + [dp] * 6
    h = lib.hc_load_models(hostcheck_lib.TABLES.encode(), abi.LIQ_MODEL[liq.lower()], abi.ICE_MODEL[ice.lower()])
    assert h
    P = lambda a: a.ctypes.data_as(dp)   # noqa: E731
    for c in (3, 12, 17, 31):
        cols = column_inputs(meridian_raw, cfg, c)
        got = oracle_cloud_optics(orc, cols)
        arrs = [np.zeros((NLEV, 16)) for _ in range(3)] + [np.zeros((NLEV, 14)) for _ in range(3)]
        lib.hc_cloud_optics(h, NLEV, *[P(a) for a in cols], 1, 0, 0, *[P(a) for a in arrs])
        for k in range(3):
            assert np.array_equal(arrs[k], got["lw"][k], equal_nan=True), ("lw", k, c)
            assert np.array_equal(arrs[3 + k], got["sw"][k], equal_nan=True), ("sw", k, c)
    lib.hc_free(h)


def test_fluxes_stay_close_to_the_default_pair(meridian_raw):
    """Different fits to the same physics: a few W m-2 apart from SOCRATES + Fu on average (the Baran fits ignore the effective radius
    and can move a single thick-cirrus column by 100-200 W m-2 in the shortwave), clear-sky fluxes untouched."""
    n = 32
    base = Oracle(RadiationConfig().consolidate()).radiation(I.to_radiation_inputs(meridian_raw, RadiationConfig().consolidate()), n, NLEV)
    for liq, ice in MODELS:
        cfg = RadiationConfig(liquid_model_name=liq, ice_model_name=ice).consolidate()
        out = Oracle(cfg).radiation(I.to_radiation_inputs(meridian_raw, cfg), n, NLEV)
        for nm in ("lw_up_clear", "sw_dn_clear"):
            assert np.array_equal(out[nm], base[nm]), (liq, ice, nm)
        for nm in ("lw_up", "lw_dn", "sw_up", "sw_dn"):
            assert np.isfinite(out[nm]).all()
            d, dm = float(np.abs(out[nm] - base[nm]).max()), float(np.abs(out[nm] - base[nm]).mean())
            assert 0.0 < d < 300.0 and dm < 10.0, (liq, ice, nm, d, dm)
        assert (out["sw_up"] >= 0.0).all() and (out["sw_dn"] >= out["sw_dn_direct"] - 1e-9).all()


def test_missing_model_is_refused_like_the_reference():
    """radiation_config.F90:1020-1025 get_enum_code aborts on an unknown name; Jahangir / Nielsen abort in cloud_optics."""
    with pytest.raises(KeyError):
        RadiationConfig(liquid_model_name="Nielsen").consolidate().to_struct()
    with pytest.raises(KeyError):
        RadiationConfig(ice_model_name="Monochromatic").consolidate().to_struct()
