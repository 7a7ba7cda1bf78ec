"""Tile-sharding of one large fused ABFT-SGEMM over the GPUs of a box (BASELINE.json config 5; new work, the
reference is single-GPU: sgemm.cu:34).

Output tiles are independent in the reference (every CTA owns one C tile and its checksums,
ft_sgemm_huge.cuh:37-41,573-574), so C is split into a P x Q grid of blocks: rank (p, q) keeps the A row-panel
A[I_p, :] (M/P x K), the B row-panel B[J_q, :] (N/Q x K) and its C block resident; no operand moves during the
product.  The one exchange step is a small all-reduce (NCCL on GPUs, gloo in the CPU tests) of the fault verdict:
[tiles, rows_checked, detected, corrected, uncorrectable, checksum_faults] summed and the residual maxima max-ed,
so that every rank agrees whether the distributed product is clean.
"""
from __future__ import annotations


def shard_grid(world: int) -> tuple[int, int]:
    """P x Q with P <= Q, P*Q == world, as square as powers of two allow: 1x1, 1x2, 2x2, 2x4."""
    if world < 1:
        raise ValueError("world must be >= 1")
    p = 1
    while (p * 2) * (p * 2) <= world and world % (p * 2) == 0:
        p *= 2
    if world % p:
        p = 1
    return p, world // p


def block_range(total: int, parts: int, idx: int, align: int = 128) -> tuple[int, int]:
    """[lo, hi) of part idx when `total` is cut into `parts` contiguous pieces on `align`-row boundaries."""
    units = -(-total // align)
    base, extra = divmod(units, parts)
    lo_u = idx * base + min(idx, extra)
    hi_u = lo_u + base + (1 if idx < extra else 0)
    return min(lo_u * align, total), min(hi_u * align, total)


def shard_extents(rank: int, world: int, M: int, N: int, align: int = 128):
    """Rows of A / C and rows of B (= columns of C) owned by `rank`: dict(p, q, m_lo, m_hi, n_lo, n_hi)."""
    P, Q = shard_grid(world)
    p, q = divmod(rank, Q)
    m_lo, m_hi = block_range(M, P, p, align)
    n_lo, n_hi = block_range(N, Q, q, align)
    return {"P": P, "Q": Q, "p": p, "q": q, "m_lo": m_lo, "m_hi": m_hi, "n_lo": n_lo, "n_hi": n_hi}


STAT_KEYS_SUM = ("tiles", "rows_checked", "detected", "corrected", "uncorrectable", "checksum_faults")
STAT_KEYS_MAX = ("max_abs_residual", "max_rel_residual")


def allreduce_verdict(stats: dict, dist, device=None) -> dict:
    """All-reduce a per-rank FtSgemm.stats() dict over the initialised torch.distributed group."""
    import torch
    s = torch.tensor([float(stats[k]) for k in STAT_KEYS_SUM], dtype=torch.float64, device=device)
    m = torch.tensor([float(stats[k]) for k in STAT_KEYS_MAX], dtype=torch.float64, device=device)
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    dist.all_reduce(m, op=dist.ReduceOp.MAX)
    out = {k: int(v) for k, v in zip(STAT_KEYS_SUM, s.tolist())}
    out.update({k: float(v) for k, v in zip(STAT_KEYS_MAX, m.tolist())})
    out["clean"] = out["detected"] == out["corrected"] and out["uncorrectable"] == 0
    return out
