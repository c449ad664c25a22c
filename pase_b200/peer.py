"""Peer-mapped device buffers for the data-parallel update (``pase_adam_flat_dp``).

One process per GPU (torchrun), all on one node: every rank allocates its flat parameter /
gradient / flag buffers with ``pase_dp_alloc`` (a dedicated ``cudaMalloc`` + CUDA IPC handle),
the 64-byte handles travel through ``torch.distributed`` (plumbing), and every rank maps every
peer's buffers into its own address space (``pase_dp_open``: NVLink / NVSwitch peer access).
After that the update kernel reads the peers' gradients and writes the peers' parameters with
plain loads and stores; no collective library is involved in the step.
SURVEY.md 8e (pure DP, one flat gradient buffer); reference: no counterpart (single GPU).
"""
import ctypes

import torch

from . import _lib

_DT = {torch.float32: ("<f4", 4), torch.int32: ("<i4", 4), torch.int64: ("<i8", 8)}


class _Blob(object):
    """__cuda_array_interface__ holder: torch wraps the pointer without copying and keeps
    this object (hence the allocation) alive."""

    def __init__(self, ptr, numel, dtype, owner):
        self.__cuda_array_interface__ = {"shape": (numel,), "typestr": _DT[dtype][0],
                                         "data": (ptr, False), "version": 2}
        self._owner = owner


class PeerAlloc(object):
    """One zero-initialised device buffer of this rank + its IPC handle."""

    def __init__(self, numel, dtype, device):
        L = _lib.lib()
        self.numel, self.dtype, self.device = int(numel), dtype, torch.device(device)
        self.nbytes = self.numel * _DT[dtype][1]
        ptr = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        with torch.cuda.device(self.device):
            rc = L.pase_dp_alloc(ctypes.c_long(self.nbytes), ctypes.byref(ptr), handle)
        if rc != 0:
            raise RuntimeError("pase_dp_alloc failed: " + _lib.last_error())
        self.ptr, self.handle = int(ptr.value), bytes(handle.raw)
        self.tensor = torch.as_tensor(_Blob(self.ptr, self.numel, dtype, self), device=self.device)

    def free(self):
        if self.ptr:
            _lib.lib().pase_dp_free(ctypes.c_void_p(self.ptr))
            self.ptr = 0


def open_peer(handle, device):
    """-> device pointer (int) of a peer's buffer mapped on `device`."""
    L = _lib.lib()
    ptr = ctypes.c_void_p()
    buf = ctypes.create_string_buffer(handle, 64)
    with torch.cuda.device(device):
        rc = L.pase_dp_open(buf, ctypes.byref(ptr))
    if rc != 0:
        raise RuntimeError("pase_dp_open failed: " + _lib.last_error())
    return int(ptr.value)


class PeerGroup(object):
    """The peer-mapped view of `names` buffers across the ranks of a process group:
    ``table(name)`` is an int64 device tensor of `world` pointers (this rank's own pointer at
    its own index)."""

    def __init__(self, allocs, group=None):
        import torch.distributed as dist
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.allocs = allocs                                  # name -> PeerAlloc (local)
        names = sorted(allocs)
        mine = {n: allocs[n].handle for n in names}
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=group)
        dev = allocs[names[0]].device
        self._opened, self._tables = [], {}
        err = None
        try:
            for n in names:
                ptrs = []
                for q in range(self.world):
                    if q == self.rank:
                        ptrs.append(allocs[n].ptr)
                    else:
                        p = open_peer(everyone[q][n], dev)
                        self._opened.append(p)
                        ptrs.append(p)
                self._tables[n] = torch.tensor(ptrs, dtype=torch.int64, device=dev)
        except Exception as exc:                  # noqa: BLE001 -- reported on every rank below
            err = exc
        # every rank learns whether EVERY rank mapped everything (also the barrier that keeps
        # anyone from touching a buffer a peer has not mapped yet)
        ok = torch.tensor([0 if err is not None else 1], dtype=torch.int32, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            self.close()
            raise RuntimeError("peer mapping (CUDA IPC) failed on at least one rank: %r" % (err,))

    def table(self, name):
        return self._tables[name]

    def close(self):
        L = _lib.lib()
        for p in self._opened:
            L.pase_dp_close(ctypes.c_void_p(p))
        self._opened = []
