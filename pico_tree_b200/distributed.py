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
