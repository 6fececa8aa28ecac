"""Multi-GPU plumbing for the query path: one process per GPU (torchrun), torch.distributed
for the collectives (NCCL over NVLink on GPUs; gloo in the CPU tests).

The path shards by queries only — the trees are read-only:
  1. rank `src` builds the mesh handle on its GPU;
  2. the handle's flat arena (one contiguous device allocation, position independent) is
     replicated with ONE broadcast (ncclBroadcast under the hood);
  3. every rank answers its contiguous slice of the queries; no communication during traversal;
  4. results are gathered to rank 0 (or stay sharded).
"""
from __future__ import annotations

import ctypes

import numpy as np

from .api import Api, Mesh


def shard_range(n: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of n queries for `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _arena_tensor(api: Api, ptr: int, nbytes: int, device):
    """Zero-copy torch view of [ptr, ptr+nbytes)."""
    import torch
    if "EMULATION" in api.version():  # host-emulated kernels (tests): "device" memory is host memory
        buf = (ctypes.c_uint8 * nbytes).from_address(ptr)
        return torch.from_numpy(np.ctypeslib.as_array(buf))

    class _Raw:
        pass

    raw = _Raw()
    raw.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}
    return torch.as_tensor(raw, device=device)


def broadcast_mesh(api: Api, mesh, src: int = 0, device=None, group=None):
    """Replicate `mesh` (valid on rank `src`, None elsewhere) to every rank.
    Returns (mesh, keepalive): non-source ranks get a handle that borrows `keepalive`."""
    import torch
    import torch.distributed as dist

    rank = dist.get_rank(group)
    emul = "EMULATION" in api.version()
    dev = torch.device("cpu") if emul else (device or torch.device("cuda", torch.cuda.current_device()))
    nb = torch.zeros(1, dtype=torch.int64, device=dev)
    if rank == src:
        ptr, nbytes = mesh.arena()
        nb[0] = nbytes
    dist.broadcast(nb, src, group=group)
    nbytes = int(nb.item())
    if rank == src:
        buf = _arena_tensor(api, ptr, nbytes, dev)
    else:
        buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    dist.broadcast(buf, src, group=group)
    if rank == src:
        return mesh, None
    return Mesh.from_arena(api, buf.data_ptr(), nbytes, copy=False), buf


def gather_rows(local: np.ndarray, n_total: int, dst: int = 0, group=None):
    """Concatenate per-rank row blocks (in rank order) on rank `dst`; None elsewhere."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    parts = [None] * world if rank == dst else None
    dist.gather_object(np.ascontiguousarray(local), parts, dst=dst, group=group)
    if rank != dst:
        return None
    out = np.concatenate(parts, axis=0)
    assert out.shape[0] == n_total
    return out


class SharedArrays:
    """ONE query set and its outputs in POSIX shared memory, mapped by every process of the job (one per GPU):
    P [3, n] and C [3, n] doubles (column-major n x 3), S, W [n] doubles, I [n] int32, and the int64 chunk counter the
    cooperative C-ABI calls (iglb200_signed_distance_coop) deal work from.  The mapping can be page-locked
    (cudaHostRegister) so that every GPU DMA-writes its chunks straight into it."""

    def __init__(self, name: str, n: int, create: bool):
        from multiprocessing import shared_memory
        self.n = int(n)
        al = lambda x: (x + 4095) & ~4095
        self._off = {}
        off = 4096
        for key, nbytes in (("P", 24 * n), ("S", 8 * n), ("I", 4 * n), ("C", 24 * n), ("W", 8 * n)):
            self._off[key] = off
            off += al(nbytes)
        self.nbytes = off
        if create:
            try:
                old = shared_memory.SharedMemory(name=name)
                old.close()
                old.unlink()
            except FileNotFoundError:
                pass
            self.shm = shared_memory.SharedMemory(name=name, create=True, size=self.nbytes)
        else:
            import time
            for _ in range(600):
                try:
                    self.shm = shared_memory.SharedMemory(name=name)
                    if self.shm.size >= self.nbytes:
                        break
                    self.shm.close()
                except FileNotFoundError:
                    pass
                time.sleep(0.05)
            else:
                raise RuntimeError("shared query set never appeared")
        buf = self.shm.buf
        self.counter = np.frombuffer(buf, np.int64, 8, 0)   # [0], [1]: ping-pong chunk counters
        self.P = np.frombuffer(buf, np.float64, 3 * n, self._off["P"]).reshape(3, n)
        self.S = np.frombuffer(buf, np.float64, n, self._off["S"])
        self.I = np.frombuffer(buf, np.int32, n, self._off["I"])
        self.C = np.frombuffer(buf, np.float64, 3 * n, self._off["C"]).reshape(3, n)
        self.W = np.frombuffer(buf, np.float64, n, self._off["W"])
        self._registered = False

    def address(self):
        return self.counter.ctypes.data

    def register(self, api):
        """Page-lock this process's mapping (cudaHostRegister): every GPU then DMA-reads P and DMA-writes its chunks of
        S/I/C/W straight from/into the shared arrays."""
        import ctypes
        api.check(api.lib.iglb200_host_register(ctypes.c_void_p(self.address()), self.nbytes))
        self._registered = api

    def close(self, unlink=False):
        if self._registered:
            import ctypes
            self._registered.lib.iglb200_host_unregister(ctypes.c_void_p(self.address()))
            self._registered = False
        self.counter = self.P = self.S = self.I = self.C = self.W = None
        try:
            self.shm.close()
            if unlink:
                self.shm.unlink()
        except Exception:
            pass
