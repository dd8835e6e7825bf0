"""Builds and loads tests/_hostcheck.so: a TEST-ONLY CPU replay of ecrad_b200/csrc/*_core.h (see hostcheck.cpp)."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SO = os.path.join(HERE, "_hostcheck.so")
SRC = os.path.join(HERE, "hostcheck.cpp")
TABLES = os.path.join(ROOT, "ecrad_b200", "data", "rrtmg_tables.bin")


def load():
    deps = [SRC] + [os.path.join(ROOT, "ecrad_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "ecrad_b200", "csrc")) if f.endswith(".h")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-o", SO, SRC])
    lib = C.CDLL(SO)
    lib.hc_load.restype = C.c_void_p
    lib.hc_load.argtypes = [C.c_char_p]
    lib.hc_free.argtypes = [C.c_void_p]
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    lib.hc_gas_column.argtypes = [C.c_void_p, C.c_int, dp, dp, dp, C.c_double] + [dp] * 7 + [ip, ip]
    lib.hc_cloud_generator.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int32, C.c_double, dp, dp, C.c_double, dp, C.c_int, dp, dp]
    lib.hc_table_sizes.argtypes = [C.c_void_p, ip, ip]
    lib.hc_expm.argtypes = [C.c_int, dp, C.c_int]
    lib.hc_fast_expm_exchange_3.argtypes = [C.c_double] * 4 + [dp]
    lib.hc_m3_solve_mat.argtypes = [dp, dp, dp]
    h = lib.hc_load(TABLES.encode())
    assert h, "hostcheck: cannot pack tables"
    return lib, h
