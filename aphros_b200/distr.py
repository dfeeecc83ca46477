"""z-slab decomposition over the GPUs of one node, one process per GPU.

The reference cuts the mesh into px*py*pz subdomains (src/distr/distr.ipp:74-84,
src/distr/native.ipp:90-104) and exchanges halos with MPI; inside the CUDA
module the rank-wide index space is cut along z only (the slowest index, so a
halo face is one contiguous xy-plane) and the exchange happens inside the
kernels over NVLink peer memory.  This file is the host-side plumbing:
``torch.distributed`` carries the NCCL id and the CUDA IPC handles between the
ranks; nothing here touches field data.
"""

from __future__ import annotations

from .solver import Mesh


def slab_partition(nz: int, nranks: int):
    """[(z0, nz_local)] per rank: contiguous planes, sizes differ by at most one,
    larger slabs first (like the reference's equal blocks when nz % nranks == 0)."""
    if nranks < 1 or nz < nranks:
        raise ValueError("cannot split %d planes over %d ranks" % (nz, nranks))
    base, extra = divmod(nz, nranks)
    out = []
    z0 = 0
    for r in range(nranks):
        n = base + (1 if r < extra else 0)
        out.append((z0, n))
        z0 += n
    return out


def neighbours(rank: int, nranks: int, periodic_z: bool):
    """(lo, hi) neighbour ranks in z, None at a non-periodic domain boundary."""
    lo = rank - 1 if rank > 0 else (nranks - 1 if periodic_z else None)
    hi = rank + 1 if rank < nranks - 1 else (0 if periodic_z else None)
    return lo, hi


def local_mesh(shape, periodic, rank, nranks, device=None, cell_volume=None) -> Mesh:
    z0, nzl = slab_partition(shape[0], nranks)[rank]
    return Mesh(shape=tuple(shape), periodic=tuple(periodic), cell_volume=cell_volume,
                rank=rank, nranks=nranks, z0=z0, nz_local=nzl,
                device=rank if device is None else device)


def all_gather_bytes(blob: bytes, group=None):
    """All-gather equal-length byte strings over a (gloo or nccl) process group."""
    import torch
    import torch.distributed as dist

    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if "nccl" in str(backend) else "cpu"
    mine = torch.tensor(list(blob), dtype=torch.uint8, device=dev)
    world = dist.get_world_size(group)
    outs = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(outs, mine, group=group)
    return [bytes(o.cpu().tolist()) for o in outs]


def broadcast_bytes(blob, nbytes: int, src: int = 0, group=None) -> bytes:
    import torch
    import torch.distributed as dist

    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if "nccl" in str(backend) else "cpu"
    if dist.get_rank(group) == src:
        t = torch.tensor(list(blob), dtype=torch.uint8, device=dev)
    else:
        t = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=src, group=group)
    return bytes(t.cpu().tolist())


def connect(solver, group=None):
    """Collective: gives `solver` (a SolverConjugateCuda built on a local_mesh) its
    NCCL communicator and its neighbours' ghost planes (include/aphcg.h, multi-GPU
    section).  No-op for a single rank."""
    import ctypes

    import torch.distributed as dist

    from . import capi

    m = solver.mesh
    if m.nranks == 1:
        return
    if dist.get_world_size(group) != m.nranks or dist.get_rank(group) != m.rank:
        raise ValueError("mesh rank/nranks do not match the process group")
    if m.rank == 0:
        buf = ctypes.create_string_buffer(capi.UNIQUE_ID_BYTES)
        capi.check(capi.lib().aphcg_comm_unique_id(buf))
        uid = buf.raw
    else:
        uid = None
    uid = broadcast_bytes(uid, capi.UNIQUE_ID_BYTES, 0, group)
    solver.CommInit(uid)
    blobs = all_gather_bytes(solver.IpcExport(), group)
    solver.IpcConnect(blobs)
    dist.barrier(group)


def bind_to_gpu_numa_node(device: int):
    """Pins the calling process to the CPUs of the NUMA node its GPU hangs off, so that the
    pinned staging buffers allocated afterwards (first touch) and the copy-engine traffic stay
    on that socket: with 8 ranks uploading rows at once, buffers that land on the other
    socket halve the aggregate host-to-device rate.  Best effort: returns a short description,
    or None when the topology cannot be read (no sysfs entry, one node, VM)."""
    import os
    import subprocess
    try:
        out = subprocess.run(["nvidia-smi", "-i", str(device), "--query-gpu=pci.bus_id",
                              "--format=csv,noheader"], capture_output=True, text=True, timeout=20)
        bus = out.stdout.strip().lower()
        if bus.count(":") == 2 and len(bus.split(":")[0]) == 8:
            bus = bus[4:]  # 00000000:1b:00.0 -> 0000:1b:00.0
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return "gpu %d (%s) -> numa node %d, %d cpus" % (device, bus, node, len(allowed))
    except Exception:
        return None
