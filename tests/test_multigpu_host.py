"""Host-side logic of the y-slab decomposition, exercised with world_size-2 gloo on CPU:
slab geometry, splitting/joining arrays, and the all_gather of exchange handles that
fv2d_b200.multigpu.connect performs before any GPU-to-GPU traffic."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fv2d_b200 import capi, multigpu


def test_slab_rows_and_split_join_roundtrip():
    Ng, Ny, Ntx = 2, 24, 11
    rng = np.random.default_rng(0)
    Q = rng.normal(size=(4, Ny + 2 * Ng, Ntx))
    for n in (1, 2, 3, 4, 5, 7, 8):  # 24 rows over 5 or 7 slabs: uneven (the first Ny % n slabs own one row more)
        slabs = [multigpu.split_global(Q, Ng, r, n) for r in range(n)]
        owned, first = [], 0
        for r, s in enumerate(slabs):
            j0, rows = multigpu.slab_rows(Ny, Ng, r, n)
            nyl = Ny // n + (1 if r < Ny % n else 0)
            assert (j0, rows) == (first, nyl + 2 * Ng) and s.shape == (4, rows, Ntx)
            first += nyl
            owned.append(nyl)
            # a slab's ghost rows are its neighbours' edge rows
            if r > 0:
                assert np.array_equal(s[:, :Ng], slabs[r - 1][:, -2 * Ng:-Ng])
            if r < n - 1:
                assert np.array_equal(s[:, -Ng:], slabs[r + 1][:, Ng:2 * Ng])
        assert sum(owned) == Ny and max(owned) - min(owned) <= 1
        assert np.array_equal(multigpu.join_slabs(slabs, Ng), Q)
    with pytest.raises(ValueError):
        multigpu.slab_rows(3, 2, 0, 4)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = bytes([rank + 1]) * capi.FV2D_IPC_HANDLE_BYTES
        allh = multigpu.gather_handles(mine, dist)
        ok = len(allh) == world * capi.FV2D_IPC_HANDLE_BYTES and all(
            allh[r * capi.FV2D_IPC_HANDLE_BYTES:(r + 1) * capi.FV2D_IPC_HANDLE_BYTES] == bytes([r + 1]) * capi.FV2D_IPC_HANDLE_BYTES
            for r in range(world))
        # global dt = max over slabs is what the device mailboxes compute; same reduction here
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        q.put((rank, ok, float(t.item())))
    finally:
        dist.destroy_process_group()


def test_handle_allgather_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True, 2.0), (1, True, 2.0)]


def test_numa_binding_is_best_effort():
    """bench.py binds every rank to its GPU's NUMA node before it allocates host arrays; on a box without
    a GPU (or without NUMA information) the helper must report, not raise, and leave the affinity alone."""
    import os

    from fv2d_b200 import multigpu

    before = os.sched_getaffinity(0)
    info = multigpu.bind_to_gpu_numa_node(0)
    assert info["device"] == 0 and ("error" in info or info["numa_node"] is not None)
    if info.get("cpus") is None:
        assert os.sched_getaffinity(0) == before
