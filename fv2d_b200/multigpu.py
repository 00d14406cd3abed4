"""One-process-per-GPU plumbing for the y-slab decomposition.

torch.distributed is used for exactly two things: telling ranks apart and all-gathering the
512-byte exchange handles once at start-up.  After `connect`, every per-step exchange (ghost
rows, the global CFL reduction) happens inside the CUDA kernels over NVLink peer mappings;
no collective is called per step.
"""
from __future__ import annotations

import numpy as np

from . import capi


def slab_rows(Ny: int, Ng: int, rank: int, nranks: int):
    """(first local row in the global array incl. ghosts, number of local rows incl. ghosts)
    of y-slab `rank`: owned rows [rank*Ny/nranks, (rank+1)*Ny/nranks) plus Ng ghost rows a side."""
    if Ny < nranks:
        raise ValueError("more slabs than rows")
    # rows are dealt out as evenly as they go: the first Ny % nranks slabs get one row more
    # (fv2d_ctx_create_slab does the same)
    nyl = Ny // nranks + (1 if rank < Ny % nranks else 0)
    return rank * (Ny // nranks) + min(rank, Ny % nranks), nyl + 2 * Ng


def split_global(Qglobal: np.ndarray, Ng: int, rank: int, nranks: int) -> np.ndarray:
    """The local slab (with its ghost rows, taken from the neighbours' rows) of a global
    [f][Nty][Ntx] array whose ghosts are already filled."""
    Ny = Qglobal.shape[1] - 2 * Ng
    j0, n = slab_rows(Ny, Ng, rank, nranks)
    return np.ascontiguousarray(Qglobal[:, j0:j0 + n, :])


def join_slabs(slabs, Ng: int) -> np.ndarray:
    """Inverse of split_global for the domain rows (ghost rows of the result are the outer
    slabs' own ghost rows)."""
    parts = [s[:, Ng:-Ng, :] for s in slabs]
    return np.concatenate([slabs[0][:, :Ng, :]] + parts + [slabs[-1][:, -Ng:, :]], axis=1)


def gather_handles(local: bytes, dist, device=None) -> bytes:
    """all_gather of the per-rank exchange handles (works with the nccl and the gloo backend)."""
    import torch

    world = dist.get_world_size()
    mine = torch.tensor(list(local), dtype=torch.uint8, device=device)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)
    return b"".join(bytes(t.cpu().tolist()) for t in out)


def connect(ctx: "capi.Context", dist) -> None:
    """Exchange IPC handles and map the neighbours' buffers (call once, collectively)."""
    import torch

    world = dist.get_world_size()
    device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else None
    handles = gather_handles(ctx.halo_export(), dist, device)
    ctx.halo_connect(handles, world)
    dist.barrier()


def connect_local(ctxs) -> None:
    """Single-process variant: contexts on several GPUs of this process."""
    handles = b"".join(c.halo_export() for c in ctxs)
    for c in ctxs:
        c.halo_connect(handles, len(ctxs))


def bind_to_gpu_numa_node(device_index: int) -> dict:
    """Pins this process to the CPUs of the NUMA node the GPU hangs off, so that the host arrays it
    allocates from here on (first touch, pinned or not) live in the memory next to the GPU's PCIe root
    port.  One process per GPU on a multi-socket box otherwise sends half of the host<->device traffic
    across the socket interconnect.  Best effort: returns what it did, never raises."""
    import os

    info = {"device": device_index, "numa_node": None, "cpus": None}
    try:
        import torch

        prop = torch.cuda.get_device_properties(device_index)
        if hasattr(prop, "pci_bus_id"):
            bdf = f"{getattr(prop, 'pci_domain_id', 0):04x}:{prop.pci_bus_id:02x}:{getattr(prop, 'pci_device_id', 0):02x}.0"
        else:  # older torch: ask NVML (same enumeration order unless CUDA_VISIBLE_DEVICES reorders)
            import pynvml

            pynvml.nvmlInit()
            bus_id = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(device_index)).busId
            bus_id = bus_id.decode() if isinstance(bus_id, bytes) else bus_id
            bdf = bus_id.lower()[-12:]  # "00000000:04:00.0" -> "0000:04:00.0"
        path = f"/sys/bus/pci/devices/{bdf}/numa_node"
        node = int(open(path).read().strip())
        info["numa_node"] = node
        if node < 0:
            return info
        cpulist = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
        cpus = set()
        for part in cpulist.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            info["cpus"] = len(cpus)
    except Exception as e:  # noqa: BLE001 - best effort by design
        info["error"] = str(e)
    return info
