"""Tile-sharding of one large fused ABFT-SGEMM over the GPUs of a box (BASELINE.json config 5; new work, the
reference is single-GPU: sgemm.cu:34).

Output tiles are independent in the reference (every CTA owns one C tile and its checksums,
ft_sgemm_huge.cuh:37-41,573-574), so C is split into a P x Q grid of blocks: rank (p, q) keeps the A row-panel
A[I_p, :] (M/P x K), the B row-panel B[J_q, :] (N/Q x K) and its C block resident; no operand moves during the
product.  The one exchange step is the fault verdict: [tiles, rows_checked, detected, corrected, uncorrectable,
checksum_faults] summed and the residual maxima max-ed over the ranks, so that every rank agrees whether the
distributed product is clean.  Three forms: `PeerVerdict` -- the exchange FUSED into the GEMM kernel: the last CTA of every
fault-tolerant launch stores the rank's verdict vector into every rank's mailbox over NVLink peer memory (CUDA IPC), no
collective launch at all (the default of bench.py at N > 1); `VerdictExchange` (the device-side vectors of
ftsgemm_stats_device, one asynchronous all-gather per step: NCCL on GPUs, gloo in the CPU tests); `allreduce_verdict`
(host dicts, synchronous).
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
    out["clean"] = out["uncorrectable"] == 0  # (everything else that was detected was corrected or recomputed)
    return out


class VerdictExchange:
    """Per-step exchange of every rank's device-side verdict vector (8 doubles, include/ftsgemm.h:
    ftsgemm_stats_device): `step()` snapshots the local counters into a slot (on the launching stream, right behind the
    GEMM it reports on) and starts ONE all-gather of that slot, asynchronously -- the collective runs on the backend's
    own stream behind the snapshot, the next GEMM is not ordered after it; `join()` makes the launching stream wait for
    every outstanding exchange (called before a timed region ends); `verdict()` reduces the last gathered vectors on the
    host: counters summed, residual maxima max-ed.

    fill(buf) writes the local vector into `buf` (a float64 tensor of 8): FtSgemm.stats_device on GPUs, any callable in
    the CPU tests."""

    SLOTS = 4

    def __init__(self, fill, dist, device=None):
        import torch
        self.fill, self.dist = fill, dist
        self.world = dist.get_world_size()
        self.local = [torch.zeros(8, dtype=torch.float64, device=device) for _ in range(self.SLOTS)]
        self.gathered = [torch.zeros(self.world * 8, dtype=torch.float64, device=device) for _ in range(self.SLOTS)]
        self.work = [None] * self.SLOTS
        self.n = 0
        self.last = None

    def step(self):
        s = self.n % self.SLOTS
        if self.work[s] is not None:  # the slot's previous exchange (SLOTS steps ago) must be over before it is rewritten
            self.work[s].wait()
        self.fill(self.local[s])
        self.work[s] = self.dist.all_gather_into_tensor(self.gathered[s], self.local[s], async_op=True)
        self.last = s
        self.n += 1

    def join(self):
        for w in self.work:
            if w is not None:
                w.wait()
        self.work = [None] * self.SLOTS

    def verdict(self) -> dict:
        """Reduced verdict of the most recent exchange (join() first)."""
        self.join()
        v = self.gathered[self.last].reshape(self.world, 8).cpu()
        out = {k: int(v[:, i].sum().item()) for i, k in enumerate(STAT_KEYS_SUM)}
        out.update({k: float(v[:, 6 + i].max().item()) for i, k in enumerate(STAT_KEYS_MAX)})
        out["clean"] = out["uncorrectable"] == 0  # (everything else that was detected was corrected or recomputed)
        out["per_rank_rows_checked"] = [int(x) for x in v[:, 1].tolist()]
        return out


class PeerVerdict:
    """The verdict exchange fused into the GEMM kernel (include/ftsgemm.h: ftsgemm_peer_*).  Construction connects the
    ranks' mailboxes once: every rank exports the CUDA-IPC handle of its mailbox, the handles travel through
    torch.distributed (all_gather_object -- plumbing only), every rank maps all of them.  From then on each fault-tolerant
    launch on `ft` publishes to all ranks from inside the kernel; `verdict()` synchronises the device, meets the other
    ranks at a barrier (every rank's last push has then landed everywhere) and reduces the mailbox: counters summed,
    residual maxima max-ed."""

    def __init__(self, ft, dist):
        import torch
        self.ft, self.dist = ft, dist
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        err = None
        try:
            mine = ft.peer_export()
        except Exception as e:  # noqa: BLE001
            mine, err = None, e
        handles = [None] * self.world
        dist.all_gather_object(handles, mine)
        if err is None and all(h is not None for h in handles):
            try:
                ft.peer_connect(self.rank, self.world, handles)
            except Exception as e:  # noqa: BLE001
                err = e
        elif err is None:
            err = RuntimeError("a peer could not export its mailbox")
        # every rank takes the same decision (an exception on one rank only would leave the others in the barrier)
        ok = torch.tensor([0.0 if err is not None else 1.0], device="cuda" if torch.cuda.is_available() else "cpu")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok.item()) < 1.0:
            raise RuntimeError(f"fused verdict exchange unavailable on at least one rank ({err})")
        dist.barrier()  # every mailbox is mapped everywhere before the first publishing launch

    def verdict(self) -> dict:
        import torch
        torch.cuda.synchronize()
        self.dist.barrier()
        tot, per = self.ft.peer_verdict(self.world, -1)
        out = {k: int(tot[i]) for i, k in enumerate(STAT_KEYS_SUM)}
        out.update({k: float(tot[6 + i]) for i, k in enumerate(STAT_KEYS_MAX)})
        out["clean"] = out["uncorrectable"] == 0
        out["per_rank_rows_checked"] = [int(v[1]) for v in per]
        return out
