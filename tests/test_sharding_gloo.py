"""N>1 host logic on CPU: two gloo ranks shard one synthetic problem, compute their columns, rank 0 gathers the flux
profiles; the result must equal the single-process run bit for bit (columns are independent, shards are contiguous).
The per-rank compute here is the CPU oracle (this is a test of sharding/gather plumbing; the GPU path is covered by
tests/test_gpu_parity.py)."""
import os
import socket
import sys

import numpy as np
import pytest

from ecrad_b200.sharding import shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_covers_all_columns_once():
    for n, g in ((10000, 8), (1000000, 8), (33, 4), (3, 8), (1, 1), (137, 2)):
        seen = np.zeros(n, dtype=int)
        for r in range(g):
            s, e = shard_range(n, r, g)
            if e >= s:
                seen[s - 1:e] += 1
        assert (seen == 1).all(), (n, g)


def _worker(rank, world, port, ncol, tmpdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    from ecrad_b200 import inputs as I
    from ecrad_b200.config import RadiationConfig
    from ecrad_b200.radiation_interface import DEFAULT_TABLES
    from ecrad_b200.sharding import broadcast_table_blob, gather_profiles, gather_slab
    from oracle_lib import Oracle

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        blob = broadcast_table_blob(DEFAULT_TABLES, dist)
        assert blob[:4] == b"ETB1" and len(blob) == os.path.getsize(DEFAULT_TABLES)
        raw = {k: np.array(v, dtype=np.float64) for k, v in np.load(os.path.join(ROOT, "tests", "golden", "ecrad_meridian_inputs.npz")).items()}
        s, e = shard_range(ncol, rank, world)
        n_loc = e - s + 1
        cfg = RadiationConfig().consolidate()
        out = Oracle(cfg).radiation(I.to_radiation_inputs(I.synthetic_columns(raw, n_loc, first=s - 1)), n_loc, 137, nthreads=2)
        got = {}
        for nm in ("lw_up", "sw_dn", "lw_dn_clear"):
            local = torch.from_numpy(np.ascontiguousarray(out[nm].T))      # (nlev+1, ncol_local), column fastest
            g = gather_profiles(local, ncol, dist)
            if rank == 0:
                got[nm] = g.numpy().T
        # the single-collective route: all profiles in one slab per rank, shards padded to a common width by the caller
        per = -(-ncol // world)
        slab = torch.zeros((3, 138, per), dtype=torch.float64)
        for k, nm in enumerate(("lw_up", "sw_dn", "lw_dn_clear")):
            slab[k, :, :n_loc] = torch.from_numpy(np.ascontiguousarray(out[nm].T))
        dest = gather_slab(slab, dist, 0)
        if rank == 0:
            for k, nm in enumerate(("lw_up", "sw_dn", "lw_dn_clear")):
                cols = [dest[r, k, :, : shard_range(ncol, r, world)[1] - shard_range(ncol, r, world)[0] + 1] for r in range(world)]
                got[nm + "_slab"] = torch.cat(cols, dim=1).numpy().T
            np.savez(os.path.join(tmpdir, "gathered.npz"), **got)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_shard_and_gather_matches_single_process(tmp_path, meridian_raw):
    import torch.multiprocessing as mp

    from ecrad_b200 import inputs as I
    from ecrad_b200.config import RadiationConfig
    from oracle_lib import Oracle

    ncol = 75   # ragged: 38 + 37
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, ncol, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "gathered.npz")
    ref = Oracle(RadiationConfig().consolidate()).radiation(I.to_radiation_inputs(I.synthetic_columns(meridian_raw, ncol)), ncol, 137)
    for nm in ("lw_up", "sw_dn", "lw_dn_clear"):
        assert np.array_equal(got[nm], ref[nm]), nm
        assert np.array_equal(got[nm + "_slab"], ref[nm]), nm + " (gather_slab)"
