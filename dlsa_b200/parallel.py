"""Multi-GPU plumbing of the one-round combine: partitions shard across ranks with no data-path collective; the only
exchange is ONE allreduce(sum, float64) of the packed sufficient statistic.

    packed = [ sum_k Sig_inv_k (p*p) | sum_k Sig_inv_k coef_k (p) | sum_k coef_k (p) | K_eff | K_total ]

Backend-agnostic (NCCL on the GPUs, gloo in the CPU tests): these helpers only move / slice the packed buffer.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def packed_len(p: int) -> int:
    return p * p + 2 * p + 2


def shard_partitions(K: int, world: int, rank: int):
    """Contiguous block of partitions owned by ``rank``: [lo, hi) with ceil(K/world) partitions per rank."""
    per = (K + world - 1) // world
    lo = min(rank * per, K)
    return lo, min(lo + per, K)


def world_size(group=None) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def allreduce_packed(packed: torch.Tensor, group=None) -> torch.Tensor:
    """In-place sum over ranks, issued on the current stream right after the local reduce kernel."""
    if world_size(group) > 1:
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    return packed


def unpack(packed: torch.Tensor, p: int):
    """-> (S [p,p], v [p], c [p], K_eff, K_total) views/values of a packed buffer."""
    if packed.numel() != packed_len(p):
        raise ValueError("packed buffer has the wrong length")
    S = packed[: p * p].view(p, p)
    v = packed[p * p: p * p + p]
    c = packed[p * p + p: p * p + 2 * p]
    tail = packed[p * p + 2 * p:].tolist()
    return S, v, c, int(round(tail[0])), int(round(tail[1]))


def pack_host(Sig_inv: torch.Tensor, Sig_invMcoef: torch.Tensor, coef: torch.Tensor, ok=None) -> torch.Tensor:
    """Reference packing in plain tensor ops (fixed k order) — used by the CPU/gloo tests to build a rank's local
    buffer from oracle outputs; on the GPU the buffer comes from kernel K4 (dlsa_partition_reduce_f64)."""
    K, p = coef.shape
    ok = torch.ones(K, dtype=torch.bool) if ok is None else ok
    out = torch.zeros(packed_len(p), dtype=torch.float64)
    for k in range(K):
        if bool(ok[k]):
            out[: p * p] += Sig_inv[k].reshape(-1)
            out[p * p: p * p + p] += Sig_invMcoef[k]
            out[p * p + p: p * p + 2 * p] += coef[k]
            out[p * p + 2 * p] += 1.0
    out[p * p + 2 * p + 1] = float(K)
    return out
