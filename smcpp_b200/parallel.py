"""Multi-GPU plumbing of the E-step: contig sharding and the single all-reduce of the packed statistics.

Contigs are independent HMMs (reference src/inference_manager.cpp:89-94: `#pragma omp parallel for` over
`hmms`), and the M-step reads the E-step only through sums over contigs (reference
src/inference_manager.cpp:121-125).  So one process per GPU takes a shard of the contigs
and the only exchange is one SUM all-reduce of [ll | gamma0 | xisum | gamma_sums] per E-step.
torch.distributed (NCCL on GPUs, gloo in the CPU tests) is the transport; nothing here computes.
"""
from __future__ import annotations

import numpy as np


def shard_contigs(lengths, world_size: int) -> list[list[int]]:
    """Longest-processing-time assignment of contigs to ranks; deterministic (ties -> lower contig index,
    lower rank).  Returns, per rank, the ascending list of contig indices it owns."""
    lengths = [int(x) for x in lengths]
    order = sorted(range(len(lengths)), key=lambda i: (-lengths[i], i))
    load = [0] * world_size
    owned = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda q: (load[q], q))
        owned[r].append(i)
        load[r] += lengths[i]
    return [sorted(o) for o in owned]


def sort_keys(keys: np.ndarray) -> np.ndarray:
    """Unique rows in lexicographic order = the reference's std::map<block_key,...> order
    (reference include/block_key.h:51-60)."""
    keys = np.asarray(keys, np.int32)
    if keys.size == 0:
        return keys.reshape(0, keys.shape[-1] if keys.ndim == 2 else 0)
    return np.unique(keys, axis=0)


def local_keys(contigs) -> np.ndarray:
    return sort_keys(np.concatenate([np.asarray(c)[:, 1:] for c in contigs], axis=0))


def union_keys(local: np.ndarray, group=None) -> np.ndarray:
    """Global key table = union over ranks (so every rank packs gamma_sums identically)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return sort_keys(local)
    gathered = [None] * dist.get_world_size(group)
    dist.all_gather_object(gathered, np.asarray(local, np.int32), group=group)
    gathered = [g for g in gathered if g is not None and len(g)]
    return sort_keys(np.concatenate(gathered, axis=0))


def unpack_reduced(vec: np.ndarray, M: int, K: int) -> dict:
    vec = np.asarray(vec)
    o = 1 + M
    return {"ll": float(vec[0]), "gamma0": vec[1:o].copy(), "xisum": vec[o:o + M * M].reshape(M, M).copy(),
            "gamma_sums": vec[o + M * M:].reshape(K, M).copy()}


def pack_reduced(ll, gamma0, xisum, gamma_sums) -> np.ndarray:
    return np.concatenate([[np.sum(ll)], np.sum(gamma0, axis=0).ravel(), np.sum(xisum, axis=0).ravel(),
                           np.sum(gamma_sums, axis=0).ravel()])


def allreduce_sum_(tensor, group=None):
    """In-place SUM all-reduce of the packed statistics (one collective per E-step)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)
    return tensor
