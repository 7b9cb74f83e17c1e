"""Multi-GPU replicas: one process per GPU, tree replicated, queries sharded.

The reference has no distributed path; queries are independent, so the only exchange is the
one-time broadcast of the tree image (points + nodes + indices, < 200 MB at 7.7M points) over
NVLink. ``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests) is the plumbing; nothing
here touches the query path.
"""
import ctypes as C

import numpy as np

from . import _lib


def shard_range(n, rank, world):
    """Contiguous block of ceil(n / world) items for `rank` (last ranks may get fewer or none)."""
    per = -(-n // world)
    begin = min(rank * per, n)
    return begin, min(begin + per, n)


def broadcast_bytes(payload, src, device=None):
    """Broadcast a byte buffer from `src` to every rank. `payload` is a torch.uint8 tensor on
    `src` (ignored elsewhere). Returns a tensor holding the bytes on every rank."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank()
    if device is None:
        device = payload.device if payload is not None else torch.device("cpu")
    size = torch.zeros(1, dtype=torch.int64, device=device)
    if rank == src:
        size[0] = payload.numel()
    dist.broadcast(size, src)
    buf = payload if rank == src else torch.empty(int(size.item()), dtype=torch.uint8, device=device)
    dist.broadcast(buf, src)
    return buf


def rebase_ragged(offsets_per_rank, flats_per_rank):
    """Concatenate per-rank ragged results ((offsets[n_r + 1], flat)) into one (offsets, flat)
    pair in rank order — how radius / box results of query shards are put back together."""
    total = 0
    offs = [np.zeros(1, dtype=np.uint64)]
    for o in offsets_per_rank:
        o = np.asarray(o, dtype=np.uint64)
        offs.append(o[1:] + np.uint64(total))
        total += int(o[-1])
    return np.concatenate(offs), np.concatenate(flats_per_rank)


def replicate_tree(tree, src=0, device_index=0):
    """Broadcast `tree` (a KdTree on rank `src`, None elsewhere) to every rank's GPU over NCCL.
    Returns the C handle usable with the pico_b200_* calls on every rank."""
    import torch
    import torch.distributed as dist
    L = _lib.lib()
    rank = dist.get_rank()
    dev = torch.device("cuda", device_index)
    image = None
    if rank == src:
        size = C.c_uint64()
        _lib.check(L.pico_b200_tree_serialize_size(tree._h, C.byref(size)))
        image = torch.empty(size.value, dtype=torch.uint8, device=dev)
        _lib.check(L.pico_b200_tree_serialize(tree._h, C.c_void_p(image.data_ptr()), 1))
        torch.cuda.synchronize()
    image = broadcast_bytes(image, src, dev)
    torch.cuda.synchronize()
    if rank == src:
        return tree._h
    handle = C.c_void_p()
    _lib.check(L.pico_b200_tree_deserialize(C.c_void_p(image.data_ptr()), image.numel(), 1, device_index,
                                            C.byref(handle)))
    return handle


# ---------------------------------------------------------------------------------- raw NCCL communicator
class _NcclUniqueId(C.Structure):
    _fields_ = [("internal", C.c_char * 128)]


def _nccl():
    """The NCCL library of this process (torch's bundled one when torch is loaded)."""
    import glob
    import os
    libs = []
    try:
        import torch
        libs = sorted(glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "nccl", "lib",
                                             "libnccl.so*")))
    except ImportError:
        pass
    lib = C.CDLL(libs[0] if libs else "libnccl.so.2", mode=C.RTLD_GLOBAL)
    lib.ncclGetUniqueId.argtypes = [C.POINTER(_NcclUniqueId)]
    lib.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, _NcclUniqueId, C.c_int]
    lib.ncclCommDestroy.argtypes = [C.c_void_p]
    return lib


def raw_nccl_comm(device_index):
    """An ncclComm_t over all ranks of the torch.distributed job, created with NCCL's own API (the unique id
    travels through torch.distributed). For pico_b200_tree_broadcast, which takes a plain ncclComm_t so that
    callers without torch (the C++ header, other runtimes) can replicate a tree. Returns (lib, comm)."""
    import torch
    import torch.distributed as dist
    lib = _nccl()
    uid = _NcclUniqueId()
    if dist.get_rank() == 0 and lib.ncclGetUniqueId(C.byref(uid)) != 0:
        raise RuntimeError("ncclGetUniqueId failed")
    dev = torch.device("cuda", device_index) if dist.get_backend() == "nccl" else torch.device("cpu")
    buf = torch.frombuffer(bytearray(C.string_at(C.byref(uid), 128)), dtype=torch.uint8).to(dev)  # all 128 bytes
    dist.broadcast(buf, 0)
    C.memmove(C.byref(uid), buf.cpu().numpy().tobytes(), 128)
    comm = C.c_void_p()
    torch.cuda.set_device(device_index)
    # NCCL announces its version on the C stdout of the process at the first communicator it creates itself; callers
    # such as bench.py promise exactly one JSON line there, so the announcement is sent to stderr
    import os
    import sys
    sys.stdout.flush()
    keep = os.dup(1)
    os.dup2(2, 1)
    try:
        rc = lib.ncclCommInitRank(C.byref(comm), dist.get_world_size(), uid, dist.get_rank())
    finally:
        os.dup2(keep, 1)
        os.close(keep)
    if rc != 0:
        raise RuntimeError("ncclCommInitRank failed")
    return lib, comm


def replicate_tree_raw_nccl(tree_handle, comm, root, device_index):
    """pico_b200_tree_broadcast: `tree_handle` is the C handle on `root` (None elsewhere); returns a handle on
    every rank (the root's own on the root)."""
    import torch.distributed as dist
    L = _lib.lib()
    h = C.c_void_p(tree_handle.value if tree_handle is not None and dist.get_rank() == root else None)
    _lib.check(L.pico_b200_tree_broadcast(C.byref(h), comm, dist.get_rank(), root, device_index))
    return h
