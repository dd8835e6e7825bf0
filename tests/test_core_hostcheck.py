"""CPU checks of the math that the CUDA kernels inline (ecrad_b200/csrc/*_core.h), replayed by the TEST-ONLY harness
tests/hostcheck.cpp against the oracle: the table packer + stencil builders for all 30 RRTMG bands, and the McICA
generator walk (bit-exact cloud masks)."""
import ctypes as C

import numpy as np
import pytest

from ecrad_b200 import inputs as I
from ecrad_b200.config import RadiationConfig
from hostcheck_lib import load
from oracle_lib import Oracle

NLEV = 137
dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731


@pytest.fixture(scope="module")
def env(meridian_raw):
    cfg = RadiationConfig().consolidate()
    lib, h = load()
    return lib, h, Oracle(cfg), cfg


def test_stencil_gas_optics_matches_oracle(env, meridian_raw):
    lib, h, orc, _ = env
    raw = I.synthetic_columns(meridian_raw, 48)
    inp = I.to_radiation_inputs(raw)
    gases = ["h2o_mmr", "co2_mmr", "ch4_mmr", "n2o_mmr", "cfc11_mmr", "cfc12_mmr", "hcfc22_mmr", "ccl4_mmr", "o3_mmr"]
    worst = {}
    for c in range(48):
        ref = orc.gas_optics_column(inp, 48, NLEV, c + 1)
        p = np.ascontiguousarray(inp["pressure_hl"][c]); t = np.ascontiguousarray(inp["temperature_hl"][c])
        gas = np.ascontiguousarray(np.stack([inp[k][c] for k in gases]))
        od_lw = np.zeros((NLEV, 140)); pf = np.zeros((NLEV, 140)); pl = np.zeros((NLEV + 1, 140)); ps = np.zeros(140)
        od_sw = np.zeros((NLEV, 112)); ssa = np.zeros((NLEV, 112)); inc = np.zeros(112)
        k1, k2 = C.c_int(), C.c_int()
        lib.hc_gas_column(h, NLEV, dp(p), dp(t), dp(gas), float(inp["skin_temperature"][c]), dp(od_lw), dp(pf), dp(pl), dp(ps),
                          dp(od_sw), dp(ssa), dp(inc), C.byref(k1), C.byref(k2))
        assert k1.value <= 21 and k2.value <= 13   # slot sizes used by the kernels
        rel = lambda a, b: np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))  # noqa: E731
        r = {"od_lw": rel(np.maximum(od_lw, 1e-15), ref["od_lw"]), "planck": rel(pl, ref["planck_hl"])}
        if inp["cos_sza"][c] > 0:
            r["od_sw"] = rel(od_sw, ref["od_sw"]); r["ssa_sw"] = rel(ssa, ref["ssa_sw"])
            r["incoming"] = rel(inc * inp["solar_irradiance"] / inc.sum(), ref["incoming_sw"])
        for k, v in r.items():
            worst[k] = max(worst.get(k, 0.0), v)
    for k, v in worst.items():
        assert v < 1e-12, (k, v)   # re-association only: ~1e-15 relative


@pytest.mark.parametrize("scheme,beta", [(1, 0), (0, 0), (1, 1), (2, 0), (2, 1)])
def test_generator_walk_is_bit_exact(env, meridian_raw, scheme, beta):
    lib, h, orc, _ = env
    L = orc.lib
    L.orc_cloud_generator.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int32, C.c_double, C.POINTER(C.c_double),
                                      C.POINTER(C.c_double), C.c_double, C.POINTER(C.c_double), C.c_int, C.c_int, C.POINTER(C.c_double),
                                      C.POINTER(C.c_double)]
    raw = I.synthetic_columns(meridian_raw, 96)
    inp = I.to_radiation_inputs(raw)
    for c in range(96):
        fr = np.ascontiguousarray(inp["cloud_fraction"][c]).copy()
        fr[(fr < 1e-6) | (inp["q_liq"][c] + inp["q_ice"][c] < 1e-9)] = 0
        op = np.ascontiguousarray(inp["overlap_param"][c]); fsd = np.ascontiguousarray(inp["fractional_std"][c])
        for ng, seed in ((112, int(inp["iseed"][c])), (140, int(inp["iseed"][c]) + 997)):
            a = np.zeros((NLEV, ng)); b = np.zeros((NLEV, ng)); ta, tb = C.c_double(), C.c_double()
            L.orc_cloud_generator(orc.t, ng, NLEV, scheme, seed, 1e-6, dp(fr), dp(op), 0.5, dp(fsd), beta, 0, dp(a), C.byref(ta))
            lib.hc_cloud_generator(h, ng, NLEV, scheme, seed, 1e-6, dp(fr), dp(op), 0.5, dp(fsd), beta, dp(b), C.byref(tb))
            assert ta.value == tb.value and np.array_equal(a, b), (c, ng)
