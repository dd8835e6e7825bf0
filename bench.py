#!/usr/bin/env python
"""bench.py -- columns/s of the ecRad hot path (LW+SW, McICA + RRTMG 140+112 g-points, 137 levels) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--ncol C]          (N>1: launched by torch.distributed.run)
    python bench.py --impl reference [...]                                  (CPU arm: the oracle port on all host threads)

One step = one pass of the whole hot path (gas optics -> cloud optics/generator -> LW and SW solvers) over one batch
of synthetic IFS-shaped columns (BASELINE.json configs[1]: 10 000 columns per GPU).
  value : whole-job columns/s with the inputs resident in HBM (device-resident C-ABI entry), CUDA-event timed on the
          launching stream, max over ranks.
  e2e   : the same through the host-buffer C-ABI entry ecrad_b200_radiation (what the Fortran shim calls), pinned host
          buffers, H2D of every input and D2H of every flux_type component inside the timed region.
Weak scaling: every rank owns `ncol` columns of one global synthetic problem (columns are independent; no data-path
collective; the barrier/max-reduce of the timing go over NCCL).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NLEV = 137
B_MIN = 30924.0  # algorithmic bytes per column, fp64 (SURVEY.md section 8d / DESIGN.md "Roofline accounting")
METRIC = "columns/sec LW+SW McICA+RRTMG 137-lev"


def load_raw():
    raw = dict(np.load(os.path.join(ROOT, "tests", "golden", "ecrad_meridian_inputs.npz")))
    return {k: np.array(v, dtype=np.float64) for k, v in raw.items()}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.proc, self.lines = gpu, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


CPU_BUILD = "gcc -O3 -march=native -fopenmp, FMA contraction on (as Makefile_include.gfortran:25 / the reference's CMake build), built on this host"


def cpu_arm(cfg, raw, ncol_sample, first, nthreads, reps):
    """Times the oracle port (C, OpenMP over columns) on `ncol_sample` columns of the same synthetic workload.  The library timed is a
    separate build of the oracle's sources with the reference build's optimisation flags (oracle_lib.build_native), made here on the
    host that runs it; the parity oracle itself (-O2, no contraction) is not what is timed."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from ecrad_b200 import inputs as I
    from oracle_lib import Oracle, build_native

    orc = Oracle(cfg, lib_path=build_native())
    inp = I.to_radiation_inputs(I.synthetic_columns(raw, ncol_sample, first=first), cfg)
    orc.radiation(dict(inp), ncol_sample, NLEV, nthreads=nthreads)  # warm-up (page-in, thread pool)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        orc.radiation(dict(inp), ncol_sample, NLEV, nthreads=nthreads)
        ts.append(time.perf_counter() - t0)
    return ncol_sample / float(np.mean(ts)), float(np.mean(ts))


def bind_to_gpu_numa_node(gpu):
    """Run this rank on the cores of the NUMA node its GPU hangs off (sysfs), so that the pinned host buffers of the end-to-end path
    are allocated (first touch) in the memory next to the GPU's PCIe root.  Returns the node or None where the platform has no
    such information (single-node hosts, containers without sysfs)."""
    try:
        import torch

        bus = torch.cuda.get_device_properties(gpu).pci_bus_id if hasattr(torch.cuda.get_device_properties(gpu), "pci_bus_id") else None
        if bus is None:
            q = subprocess.run(["nvidia-smi", f"--id={gpu}", "--query-gpu=pci.bus_id", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
            busid = q.lower()[4:] if q.lower().startswith("0000") and len(q) > 12 else q.lower()
        else:
            busid = f"0000:{bus:02x}:00.0"
        node = int(open(f"/sys/bus/pci/devices/{busid}/numa_node").read())
        if node < 0:
            return None
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


# --workload: BASELINE.json configs (the default, configs[1], is the one the metric is quoted on; the others are extra lines)
WORKLOADS = {
    "mcica_rrtmg": (dict(), 10000, "McICA LW+SW, RRTMG 140+112 g-points", "configCY49R1.nam, use_aerosols=false", "BASELINE.json configs[1]"),
    "cloudless_ecckd32": (dict(gas_model_name="ECCKD", do_nearest_spectral_lw_emiss=False, sw_solver_name="Cloudless", lw_solver_name="Cloudless"),
                          10000, "Cloudless LW+SW, ecCKD 32+32 g-points", "configCY49R1_ecckd.nam, Cloudless, use_aerosols=false", "BASELINE.json configs[0] on the GPU"),
    "mcica_ecckd32": (dict(gas_model_name="ECCKD", do_nearest_spectral_lw_emiss=False), 10000, "McICA LW+SW, ecCKD 32+32 g-points",
                      "configCY49R1_ecckd.nam, McICA, use_aerosols=false", "extra"),
    "tripleclouds_ecckd64": (dict(gas_model_name="ECCKD", do_nearest_spectral_lw_emiss=False, sw_solver_name="Tripleclouds", lw_solver_name="Tripleclouds",
                                  ecckd_tables="ecckd_tables_64b.bin"), 100000, "Tripleclouds LW+SW, ecCKD 64+64 g-points",
                             "configCY49R1_ecckd.nam + 64-term models, use_aerosols=false", "BASELINE.json configs[2]"),
    "tripleclouds_rrtmg": (dict(sw_solver_name="Tripleclouds", lw_solver_name="Tripleclouds"), 10000, "Tripleclouds LW+SW, RRTMG 140+112 g-points",
                           "configCY49R1.nam, Tripleclouds, use_aerosols=false", "extra"),
    "tripleclouds_mixed": (dict(sw_gas_model_name="ECCKD", do_nearest_spectral_lw_emiss=False, sw_solver_name="Tripleclouds", lw_solver_name="Tripleclouds",
                                use_aerosols=True), 10000, "Tripleclouds, SW ecCKD 32 g-points + LW RRTMG 140 g-points (mixed gas models), aerosols",
                           "configCY49R1_mixed.nam, lw_gas_model_name=RRTMG-IFS (test/ifs `test_mixed_gas`, second run)", "extra"),
    "spartacus_rrtmg": (dict(sw_solver_name="SPARTACUS", lw_solver_name="SPARTACUS", do_3d_effects=True), 50000,
                        "SPARTACUS LW+SW 3 regions with 3D effects, RRTMG 140+112 g-points",
                        "configCY49R1.nam, SPARTACUS, do_3d_effects=true, use_aerosols=false (ctest `spartacus`)", "BASELINE.json configs[4]"),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="mcica_rrtmg", choices=sorted(WORKLOADS))
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--ncol", type=int, default=0, help="columns per GPU (default: the workload's, 10 000 for BASELINE configs[1])")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="columns in the CPU-baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-sweep", default="", help="tuning aid: 'tile:edge:ramp[:tail],...' host-entry tile schedules timed after the e2e measurement (key e2e.sweep)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    from ecrad_b200 import inputs as I
    from ecrad_b200.config import RadiationConfig

    wkw, wncol, wname, wnam, wref = WORKLOADS[args.workload]
    args.ncol = args.ncol or wncol
    cfg = RadiationConfig(**wkw).consolidate()   # default: test/ifs/configCY49R1.nam with use_aerosols=false
    raw = load_raw()
    ncores = os.cpu_count() or 1
    config = {"workload": f"{wname}, {NLEV} levels, {args.ncol} synthetic IFS columns per GPU ({wref})",
              "ncol_per_gpu": args.ncol, "nlev": NLEV, "namelist": wnam,
              "sharding": f"columns x{world}, no data-path collective",
              "numa": "each rank bound to the NUMA node of its GPU (sysfs) before the pinned host buffers are allocated",
              "l2": "inputs+scratch per step (>3 GB) exceed the 126 MB L2; no explicit flush"}

    # ------------------------------------------------------------------ CPU reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        sample = args.cpu_sample or min(args.ncol, 2000 if args.workload.startswith("spartacus") else 10000)
        warm = max(args.warmup, 1)
        for _ in range(warm - 1):
            cpu_arm(cfg, raw, sample, 0, ncores, 1)
        v, sec = cpu_arm(cfg, raw, sample, 0, ncores, max(args.steps, 1))
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "columns/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": "columns/s", "cores": ncores, "kind": "port",
                                 "sample": f"{sample} columns of the same synthetic workload per step; C/OpenMP oracle port, {CPU_BUILD} "
                                           "(the Fortran reference cannot be built: no Fortran compiler in the image or on the GPU box, profiles/r2a_probe_box.txt)"},
                "e2e": {"value": v, "unit": "columns/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ B200 arm
    import torch

    from ecrad_b200 import abi
    from ecrad_b200.radiation_interface import setup_radiation

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank)   # pinned host buffers (first touch) and the calling thread next to this rank's GPU
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ncol = args.ncol
    blob = None
    if dist is not None:   # rank 0 reads the optics tables, one NCCL broadcast hands them to every GPU
        from ecrad_b200.sharding import broadcast_table_blob
        blob = broadcast_table_blob(cfg.tables_path(), dist, device=dev)
    h = setup_radiation(cfg, tables_blob=blob)
    inp = I.to_radiation_inputs(I.synthetic_columns(raw, ncol, first=rank * ncol), cfg)

    # ---- host buffers (pinned) for the e2e path ----
    pinned, host_in = {}, {}
    in_arrays = [(nm, dt) for nm, dt, _ in abi.INPUT_ARRAYS if nm in inp]   # optional inputs (aerosols, cloud sizes) may be absent
    for nm, dt in in_arrays:
        a = np.asfortranarray(inp[nm], dtype=np.int32 if dt == "i4" else np.float64)
        t = torch.from_numpy(np.ascontiguousarray(a.T)).pin_memory()   # (rows, ncol) C order == Fortran (ncol, rows)
        pinned[nm] = t
        host_in[nm] = t.numpy().T
    host_in["solar_irradiance"] = inp["solar_irradiance"]
    out_names = [(nm, kind) for nm, kind in abi.OUTPUT_ARRAYS if kind not in ("pl", "ps")]
    host_out, ost_host = {}, abi.Outputs()
    ost_host.struct_bytes = C.sizeof(abi.Outputs)
    for nm, kind in out_names:
        shp = abi.output_shape(kind, ncol, NLEV, h.cfg)
        t = torch.empty(tuple(reversed(shp)), dtype=torch.float64).pin_memory()
        t.fill_(-1.0 if nm.startswith("cloud_cover") else float("nan"))
        host_out[nm] = t
        setattr(ost_host, nm, C.cast(t.data_ptr(), abi.c_dp))
    keep_h, ist_host = abi.make_inputs(host_in, inp["solar_irradiance"])
    for nm, dt in in_arrays:   # make_inputs must not have copied: point at the pinned memory
        setattr(ist_host, nm, C.cast(pinned[nm].data_ptr(), abi.c_ip if dt == "i4" else abi.c_dp))
    h2d_bytes = sum(t.numel() * t.element_size() for t in pinned.values()) + ncol * 8
    d2h_bytes = sum(t.numel() * 8 for t in host_out.values()) + ncol * NLEV * 8

    # ---- device-resident buffers for the kernel-only path ----
    dev_in, ist_dev = {}, abi.Inputs()
    ist_dev.struct_bytes = C.sizeof(abi.Inputs)
    ist_dev.solar_irradiance = inp["solar_irradiance"]
    for nm, dt in in_arrays:
        dev_in[nm] = pinned[nm].to(dev)
        setattr(ist_dev, nm, C.cast(dev_in[nm].data_ptr(), abi.c_ip if dt == "i4" else abi.c_dp))
    dev_out, ost_dev = {}, abi.Outputs()
    ost_dev.struct_bytes = C.sizeof(abi.Outputs)
    prof_names = [nm for nm, kind in out_names if kind == "h"]
    arena = torch.zeros((len(prof_names), NLEV + 1, ncol), dtype=torch.float64, device=dev)   # all flux profiles, one slab
    for nm, kind in out_names:
        shp = abi.output_shape(kind, ncol, NLEV, h.cfg)
        dev_out[nm] = arena[prof_names.index(nm)] if kind == "h" else torch.zeros(tuple(reversed(shp)), dtype=torch.float64, device=dev)
        setattr(ost_dev, nm, C.cast(dev_out[nm].data_ptr(), abi.c_dp))
    stream = torch.cuda.current_stream()

    def step_device():
        # (cloud%crop_cloud_fraction is in place and idempotent, so repeated steps see the same effective input)
        h.radiation_device(ncol, NLEV, ist_dev, ost_dev, stream=stream.cuda_stream)

    def step_host():
        rc = h.lib.ecrad_b200_radiation(h.h, ncol, NLEV, 1, ncol, C.byref(ist_host), C.byref(ost_host))
        if rc:
            raise RuntimeError(h._err())

    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = h.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_ms = {}
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    barrier()
    launches = h.kernel_launches() - l0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    # per-stage kernel durations for the roofline: one extra pass with the three chains serialised on the launching stream
    # (in the timed region they overlap on three streams, so per-kernel durations are not separable there)
    h.set_option("serial", 1)
    step_device(); step_device()
    torch.cuda.synchronize()
    stage_ms = h.last_stage_ms()
    h.set_option("serial", 0)
    ms_step = ms_total / args.steps
    value = world * ncol / (ms_step * 1e-3)

    # ---- end-to-end through the host-buffer entry ----
    for _ in range(max(args.warmup, 3)):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    barrier()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / args.steps)
    e2e_value = world * ncol / e2e_s
    clocks = sampler.stop()

    if args.e2e_sweep and world == 1:   # tuning aid: the host entry's tile schedule (tile_cols : edge_cols : ramp)
        sweep = {}
        for spec in args.e2e_sweep.split(","):
            tile, edge, ramp, tail = (int(x) for x in (spec + ":1").split(":")[:4])
            h.set_option("tile_cols", tile); h.set_option("edge_cols", edge); h.set_option("tile_ramp", ramp); h.set_option("tail_tiles", tail)
            for _ in range(2):
                step_host()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                step_host()
            sweep[spec] = round(ncol / ((time.perf_counter() - t0) / args.steps))
        print("e2e sweep (columns/s):", json.dumps(sweep), file=sys.stderr)
        raise SystemExit(0)

    # What an unmodified Fortran host passes: ordinary (pageable) allocatables.  Same call, same arrays, copied out of pinned memory
    # into plain numpy arrays: (a) as they are -- cudaMemcpy2DAsync from pageable memory is staged by the driver and serialises with
    # the kernels; (b) with set_option("register_host", 1): the library page-locks the caller's arrays once (cudaHostRegister).
    e2e_extra = {}
    if world == 1:
        page_in = {nm: np.array(t.numpy(), copy=True) for nm, t in pinned.items()}
        page_out = {nm: np.array(t.numpy(), copy=True) for nm, t in host_out.items()}
        ist_p, ost_p = abi.Inputs(), abi.Outputs()
        C.memmove(C.byref(ist_p), C.byref(ist_host), C.sizeof(abi.Inputs))
        C.memmove(C.byref(ost_p), C.byref(ost_host), C.sizeof(abi.Outputs))
        for nm, dt in in_arrays:
            setattr(ist_p, nm, C.cast(page_in[nm].ctypes.data, abi.c_ip if dt == "i4" else abi.c_dp))
        for nm, kind in out_names:
            setattr(ost_p, nm, C.cast(page_out[nm].ctypes.data, abi.c_dp))

        def step_page():
            if h.lib.ecrad_b200_radiation(h.h, ncol, NLEV, 1, ncol, C.byref(ist_p), C.byref(ost_p)):
                raise RuntimeError(h._err())

        for key, reg in (("pageable", 0), ("registered", 1)):
            h.set_option("register_host", reg)
            for _ in range(2):
                step_page()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                step_page()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / args.steps
            e2e_extra[key] = {"value": ncol / dt, "ms_per_step": dt * 1e3}
        h.set_option("register_host", 0)
        # the single-precision boundary (ecrad_b200_radiation_sp): float arrays on the host side, half the PCIe bytes, fp64 kernels
        f_in = {nm: (t if dt == "i4" else t.to(torch.float32).pin_memory()) for (nm, dt), t in zip(in_arrays, [pinned[nm] for nm, _ in in_arrays])}
        f_out = {nm: torch.empty_like(t, dtype=torch.float32).pin_memory() for nm, t in host_out.items()}
        ist_f, ost_f = abi.Inputs(), abi.Outputs()
        C.memmove(C.byref(ist_f), C.byref(ist_host), C.sizeof(abi.Inputs))
        C.memmove(C.byref(ost_f), C.byref(ost_host), C.sizeof(abi.Outputs))
        for nm, dt in in_arrays:
            setattr(ist_f, nm, C.cast(f_in[nm].data_ptr(), abi.c_ip if dt == "i4" else abi.c_dp))
        for nm, kind in out_names:
            setattr(ost_f, nm, C.cast(f_out[nm].data_ptr(), abi.c_dp))
        for _ in range(2):
            if h.lib.ecrad_b200_radiation_sp(h.h, ncol, NLEV, 1, ncol, C.byref(ist_f), C.byref(ost_f)):
                raise RuntimeError(h._err())
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            h.lib.ecrad_b200_radiation_sp(h.h, ncol, NLEV, 1, ncol, C.byref(ist_f), C.byref(ost_f))
        torch.cuda.synchronize()
        dt_sp = (time.perf_counter() - t0) / args.steps
        sp_err = float((f_out["lw_up"].double() - host_out["lw_up"]).abs().max())
        assert sp_err <= 1e-3, f"single-precision boundary: lw_up differs from the fp64 result by {sp_err} W m-2"
        e2e_extra["sp"] = {"value": ncol / dt_sp, "ms_per_step": dt_sp * 1e3, "max_abs_diff_lw_up_vs_fp64": sp_err,
                           "h2d_bytes_per_step": h2d_bytes // 2, "d2h_bytes_per_step": d2h_bytes // 2}
        assert np.array_equal(page_out["lw_up"], host_out["lw_up"].numpy(), equal_nan=True), "pageable-host results differ from the pinned-host results"

    # Exchange step of a host model that wants all fluxes on one GPU.  Not part of `value` (the path itself needs
    # no collective).  Two ways: (1) NCCL gather of the profiles after the step; (2) no gather at all -- every rank's flux kernels
    # store their column slice straight into rank 0's arrays over NVLink (peer-mapped memory, ecrad_b200_radiation_device_ld).
    gather = None
    if dist is not None:
        from ecrad_b200.sharding import PeerFluxArrays, gather_slab
        prof = prof_names
        nbytes = arena.numel() * 8 * world
        # (1) one NCCL gather of the whole slab into a preallocated destination, after the step
        dest = torch.empty((world,) + tuple(arena.shape), dtype=torch.float64, device=dev) if rank == 0 else None
        gather_slab(arena, dist, 0, dest)   # warm-up: NCCL sets up its channels on the first collective
        barrier()
        g0 = torch.cuda.Event(enable_timing=True); g1 = torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(args.steps):
            gather_slab(arena, dist, 0, dest)
        g1.record()
        barrier()
        nccl_ms = max_over_ranks(g0.elapsed_time(g1)) / args.steps
        if rank == 0:
            assert torch.equal(dest[0], arena), "gathered slab differs from the local fluxes"
        # step + gather back to back, as a host model that needs the fluxes on one GPU every step would run it
        for _ in range(2):
            step_device(); gather_slab(arena, dist, 0, dest)
        barrier()
        g0.record(stream)
        for _ in range(args.steps):
            step_device(); gather_slab(arena, dist, 0, dest)
        g1.record(stream)
        barrier()
        step_gather_ms = max_over_ranks(g0.elapsed_time(g1)) / args.steps
        del dest
        fused = push = None
        try:
            peer = PeerFluxArrays(prof, NLEV + 1, world * ncol, dist, dst=0)
        except RuntimeError as exc:   # raised on every rank together (sharding.PeerFluxArrays agrees over the group)
            peer, fused = None, {"unavailable": str(exc)}
        if peer is not None:
            # (2) push: after its flux kernels every rank copies its slab into its column slice of rank 0's arrays over NVLink (peer-mapped
            # memory), rows of ncol * 8 bytes per store run; two output arenas alternate so that the push of step i overlaps step i+1
            arena2 = torch.zeros_like(arena)
            ost_b = abi.Outputs()
            C.memmove(C.byref(ost_b), C.byref(ost_dev), C.sizeof(abi.Outputs))
            for k, nm in enumerate(prof):
                setattr(ost_b, nm, C.cast(arena2[k].data_ptr(), abi.c_dp))
            view = peer.slice_view(rank * ncol, ncol)
            side = torch.cuda.Stream(device=dev)
            evs = [torch.cuda.Event(), torch.cuda.Event()]
            pushed = [torch.cuda.Event(), torch.cuda.Event()]

            def step_push(i):
                a, o = (arena, ost_dev) if i % 2 == 0 else (arena2, ost_b)
                stream.wait_event(pushed[i % 2])            # this arena's previous push has left
                h.radiation_device(ncol, NLEV, ist_dev, o, stream=stream.cuda_stream)
                evs[i % 2].record(stream)
                side.wait_event(evs[i % 2])
                with torch.cuda.stream(side):
                    view.copy_(a, non_blocking=True)
                    pushed[i % 2].record(side)

            for i in range(4):
                step_push(i)
            side.synchronize(); barrier()
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p0.record(stream)
            for i in range(args.steps):
                step_push(i)
            stream.wait_stream(side)
            p1.record(stream)
            barrier()
            push_ms = max_over_ranks(p0.elapsed_time(p1)) / args.steps
            sums = torch.stack([dev_out[nm].view(torch.int64).sum() for nm in prof])
            allsums = [torch.empty_like(sums) for _ in range(world)]
            dist.all_gather(allsums, sums)
            if rank == 0:
                ok = True
                for k, nm in enumerate(prof):
                    full = peer.tensor(nm)
                    for r in range(world):
                        ok = ok and bool(full[:, r * ncol:(r + 1) * ncol].contiguous().view(torch.int64).sum() == allsums[r][k])
                assert ok, "pushed flux arrays differ from the local results"
            push = {"ms_per_step": push_ms, "value": world * ncol / (push_ms * 1e-3), "unit": "columns/s",
                    "how": "every rank copies its flux slab into its column slice of rank 0's arrays over NVLink (peer-mapped memory, one strided device copy "
                           "per step on a side stream, overlapping the next step), checked bit-exact"}
            # (3) no copy at all: the flux kernels store straight into the peer arrays (isolated 8-byte remote stores: one CTA = one column)
            ost_peer = abi.Outputs()
            C.memmove(C.byref(ost_peer), C.byref(ost_dev), C.sizeof(abi.Outputs))
            for nm in prof:
                setattr(ost_peer, nm, C.cast(peer.pointer(nm, rank * ncol), abi.c_dp))

            def step_peer():
                h.radiation_device_ld(ncol, NLEV, ncol, world * ncol, ist_dev, ost_peer, stream=stream.cuda_stream)

            for _ in range(3):
                step_peer()
            barrier()
            p0.record(stream)
            for _ in range(args.steps):
                step_peer()
            p1.record(stream)
            barrier()
            peer_ms = max_over_ranks(p0.elapsed_time(p1)) / args.steps
            peer.close()
            fused = {"ms_per_step": peer_ms, "value": world * ncol / (peer_ms * 1e-3), "unit": "columns/s",
                     "how": "flux kernels of every rank write their column slice into rank 0's arrays over NVLink (CUDA IPC peer mapping)"}
            step_device(); torch.cuda.synchronize()   # dev_out holds this rank's own results again for the parity guard below
        best = min([x for x in (step_gather_ms, push and push["ms_per_step"], fused and fused.get("ms_per_step")) if x])
        gather = {"profiles": len(prof), "bytes_per_step": nbytes, "nccl_gather_ms": nccl_ms, "nccl_gather_gbs": nbytes * (world - 1) / world / (nccl_ms * 1e-3) / 1e9,
                  "step_plus_nccl_gather_ms": step_gather_ms, "p2p_push": push, "fused_p2p_stores": fused,
                  "value_with_gather": world * ncol / (best * 1e-3),
                  "note": "value_with_gather = columns/s when all flux profiles must be on GPU 0 after every step (best of the three ways)"}

    # parity guard: the timed outputs are the real thing (first 32 columns of rank 0 = the golden test slice)
    if rank == 0:
        g = np.load(os.path.join(ROOT, "tests", "golden", "ecrad_meridian_noaer_ref.npz"))
        for nm, gn in (("sw_dn", "flux_dn_sw"), ("lw_up", "flux_up_lw")):
            a = dev_out[nm].cpu().numpy().T[:32]
            b = host_out[nm].numpy().T[:32]
            # the golden file is the RRTMG one: other gas models / solvers differ from it by physics (a few W m-2), not by bugs
            err = np.abs(a - g[gn]).max() if args.workload == "mcica_rrtmg" else 0.0
            assert err <= 1e-3 and np.array_equal(a, b) and np.isfinite(a).all(), f"bench outputs are wrong: {nm} {err}"

    # ---- roofline of the dominant kernel ----
    dom = max(stage_ms, key=stage_ms.get) if stage_ms else None
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and dom and args.workload in ("mcica_rrtmg", "spartacus_rrtmg"):   # ncu figures exist for these two
        tj = json.load(open(tpath))
        tj = tj if args.workload == "mcica_rrtmg" else tj.get(args.workload, {})
        if dom in tj:
            traffic = tj[dom]["dram_bytes_per_column"] * ncol
    # whole-step figures from the ncu captures (profiles/traffic.json: DRAM bytes and fp64 flops per column and stage) and the fp64
    # multiply-add throughput measured on this device (ecrad_b200_measure_fp64)
    alu = whole = None
    if rank == 0:
        tfl = C.c_double(0.0)
        fp64_peak = float(tfl.value) if h.lib.ecrad_b200_measure_fp64(C.byref(tfl)) == 0 else 0.0
        tj = json.load(open(tpath)) if os.path.exists(tpath) else {}
        tj = tj if args.workload == "mcica_rrtmg" else tj.get(args.workload, {})
        st = {k: v for k, v in tj.items() if isinstance(v, dict) and "dram_bytes_per_column" in v}
        if st and stage_ms:
            step_s = ms_step * 1e-3
            tot_bytes = sum(v["dram_bytes_per_column"] for v in st.values()) * ncol
            tot_flop = sum(v.get("fp64_flop_per_column", 0.0) for v in st.values()) * ncol
            whole = {"dram_bytes_per_step": tot_bytes, "dram_gbs": tot_bytes / step_s / 1e9, "dram_frac": tot_bytes / step_s / 1e9 / peak,
                     "algorithmic_bytes_per_step": B_MIN * ncol, "algorithmic_frac": B_MIN * ncol / step_s / 1e9 / peak,
                     "source": "sum over the stages of profiles/traffic.json (ncu --set full captures of this code) / measured step time"}
            if tot_flop and fp64_peak:
                alu = {"tflops_achieved": tot_flop / step_s / 1e12, "tflops_peak": fp64_peak, "frac": tot_flop / step_s / 1e12 / fp64_peak,
                       "flop_per_column": tot_flop / ncol, "peak_source": "measured: ecrad_b200_measure_fp64 (8 independent DFMA chains per thread, 8 CTAs per SM)",
                       "how": "2 x DFMA + DMUL + DADD thread instructions of every kernel (ncu) / measured step time"}
    roofline = None
    if dom:
        achieved = B_MIN * ncol / (stage_ms[dom] * 1e-3) / 1e9
        # which resource is closer to its ceiling over the whole step: DRAM traffic actually moved, or fp64 arithmetic actually issued
        bound = "hbm"
        if whole and alu:
            bound = "hbm" if whole["dram_frac"] >= alu["frac"] else "fp64-alu"
        roofline = {"bound": bound, "kernel": dom, "whole_step": whole, "alu": alu, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic,
                    # the same stage on the DRAM bytes it really moves (ncu): how close the adding-method scratch traffic runs to the HBM peak
                    "dram_achieved": (traffic / (stage_ms[dom] * 1e-3) / 1e9) if traffic else None,
                    "dram_frac": (traffic / (stage_ms[dom] * 1e-3) / 1e9 / peak) if traffic else None, "peak_source": peak_src, "kernel_ms": stage_ms[dom], "stage_ms": stage_ms, "stage_ms_mode": "serialised extra pass (CUDA events on the launching stream)",
                    "note": "`achieved`/`frac` are on ALGORITHMIC bytes (30.9 kB/column) of the dominant stage, per the contract; `whole_step` and `alu` say how "
                            "close the step runs to the DRAM and fp64 ceilings on what it really moves and computes; neither is saturated: the kernels are "
                            "latency-bound at 25-35 % occupancy (DESIGN.md section 4)"}

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            sample = args.cpu_sample or min(ncol, 2000 if args.workload.startswith("spartacus") else 10000)
            v, sec = cpu_arm(cfg, raw, sample, 0, ncores, 3)
            cpu = {"value": v, "unit": "columns/s", "cores": ncores, "kind": "port",
                   "sample": f"{sample} columns of the same workload, 3 repetitions ({sec:.2f} s each), C/OpenMP oracle port, {CPU_BUILD}"}
        line = {"metric": METRIC, "value": value, "unit": "columns/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": config, "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "columns/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                        "ms_per_step": e2e_s * 1e3, "host_memory": "pinned",
                        # the same call with the host arrays an unmodified Fortran driver has (pageable), without and with
                        # set_option("register_host", 1)
                        "pageable_host": e2e_extra.get("pageable"), "pageable_host_registered": e2e_extra.get("registered"),
                        # the same call from a single-precision host (float arrays, ecrad_b200_radiation_sp; kernels still fp64)
                        "single_precision_host": e2e_extra.get("sp")},
                "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu}
        if gather is not None:
            line["gather"] = gather
        print(json.dumps(line))
    h.finalize()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
