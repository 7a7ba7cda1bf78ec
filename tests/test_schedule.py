"""CPU checks of the kernel's work decomposition (stream-K head + data-parallel body + checksum tile-columns), using
the host-side enumeration of the SAME inline code the device runs (ftsgemm_debug_schedule):
  * every (tile, k-block) is computed exactly once; every tile has exactly one finishing segment;
  * a finisher's contributors are exactly the units the kernel's lookup rule visits;
  * no circular wait: with in-order roles, 2 (or 1) TMEM accumulator stages per unit, finishers waiting for
    contributors and ABFT data tiles waiting for their checksum tile-columns, every segment completes."""
import itertools

import pytest

TILE_N = {1: 32, 2: 64, 6: 128, 5: 256, 21: 256, 22: 128, 11: 32, 12: 64, 16: 128, 15: 256, 31: 256, 32: 128}


def _simulate(hdr, segs, acc_stages):
    units = hdr["units"]
    per_unit = [[] for _ in range(units)]
    for s in segs:
        per_unit[s["unit"]].append(s)
    tiles_c = 0
    chk_done = {}     # (m_blk, c) -> bool
    contrib = {}      # tile -> list of (unit, idx)
    for u, lst in enumerate(per_unit):
        for i, s in enumerate(lst):
            if s["is_chk"]:
                chk_done[(s["m_blk"], s["n_blk"])] = False
                tiles_c = max(tiles_c, s["n_blk"] + 1)
            if s["kind"] == 1:
                contrib.setdefault(s["tile"], []).append((u, i))
    mma_done = [0] * units   # number of segments whose MMA finished
    epi_done = [0] * units
    done_epi = set()         # (unit, idx)
    progress = True
    while progress:
        progress = False
        for u, lst in enumerate(per_unit):
            # MMA of segment i needs the accumulator stage used by segment i - acc_stages to be drained
            while mma_done[u] < len(lst) and mma_done[u] - epi_done[u] < acc_stages:
                mma_done[u] += 1
                progress = True
            while epi_done[u] < mma_done[u]:
                i = epi_done[u]
                s = lst[i]
                ok = True
                if s["kind"] == 2:
                    ok = all(c in done_epi for c in contrib.get(s["tile"], []))
                if ok and hdr["n_chk_tiles"] and not s["is_chk"] and s["kind"] != 1:
                    ok = all(chk_done[(s["m_blk"], c)] for c in range(tiles_c))
                if not ok:
                    break
                epi_done[u] += 1
                done_epi.add((u, i))
                if s["is_chk"] and s["kind"] != 1:
                    chk_done[(s["m_blk"], s["n_blk"])] = True
                progress = True
    return all(epi_done[u] == len(per_unit[u]) for u in range(units))


@pytest.mark.parametrize("kid", [1, 6, 5, 21, 22, 12, 16, 15, 31, 32])
@pytest.mark.parametrize("shape", [(4096, 4096, 4096), (2048, 2048, 2048), (1024, 1024, 1024), (8192, 8192, 512),
                                   (1536, 2560, 1000), (256, 256, 64), (128, 4096, 32), (5120, 384, 4096),
                                   (16384, 16384, 1024), (32768, 1024, 256)])
@pytest.mark.parametrize("num_sms", [148, 16, 6])
def test_decomposition_covers_and_cannot_deadlock(ft, kid, shape, num_sms):
    M, N, K = shape
    hdr, segs = ft.debug_schedule(kid, M, N, K, num_sms)
    units, num_kb, H, S = hdr["units"], hdr["num_kb"], hdr["sk_tiles"], hdr["sk_slices"]
    assert hdr["num_kb"] == -(-K // 32)
    first_tail = hdr["num_tiles"] - H
    cover = {}
    finishers = {}
    for s in segs:
        assert 0 <= s["kb_begin"] < s["kb_end"] <= num_kb
        cover.setdefault(s["tile"], []).append((s["kb_begin"], s["kb_end"], s["unit"], s["kind"]))
        if s["kind"] != 1:
            assert s["kb_end"] == num_kb
            assert s["tile"] not in finishers
            finishers[s["tile"]] = s
        else:
            assert s["kb_end"] < num_kb and s["tile"] >= first_tail and not s["is_chk"]
    assert sorted(cover) == list(range(hdr["num_tiles"]))
    for t, pieces in cover.items():
        pieces.sort()
        assert pieces[0][0] == 0 and pieces[-1][1] == num_kb
        for a, b in zip(pieces, pieces[1:]):
            assert a[1] == b[0]  # contiguous, no overlap
        fin = finishers[t]
        if fin["kind"] == 0:
            assert len(pieces) == 1
        else:
            assert t >= first_tail and sorted(p[3] for p in pieces) == [1] * (S - 1) + [2]
    # global item order: [checksum tiles][whole data tiles][split tiles, slice-major] or, "head first",
    # [checksum tiles][split tiles, slice-major][whole data tiles]; every unit's list is increasing in it, so every
    # wait (finisher -> earlier slices of its tile, data tile -> checksum tiles) points to an earlier item
    whole_first = {}
    for s_ in segs:
        if s_["kind"] == 0 and not s_["is_chk"]:
            whole_first.setdefault(s_["unit"], None)
    head_first = False
    if S > 1:
        # detect the order from any unit that owns both kinds of item
        per_unit_kinds = {}
        for s_ in segs:
            if not s_["is_chk"]:
                per_unit_kinds.setdefault(s_["unit"], []).append(s_["kind"] != 0)
        for kinds in per_unit_kinds.values():
            if True in kinds and False in kinds:
                head_first = kinds[0]
                break
    n_chk = hdr["n_chk_tiles"]
    def gidx(s_):
        if s_["is_chk"]:
            return s_["tile"]
        if S == 1:
            return s_["tile"]
        if s_["tile"] >= first_tail:
            off = n_chk if head_first else first_tail
            return off + s_["slice"] * H + (s_["tile"] - first_tail)
        return s_["tile"] + (H * S if head_first else 0)
    last = {}
    for s_ in segs:
        g = gidx(s_)
        assert last.get(s_["unit"], -1) < g
        last[s_["unit"]] = g
        if s_["is_chk"]:
            assert s_["tile"] < hdr["n_chk_tiles"]
    # checksum tiles: n_chk_tiles of them, 8 columns per N-tile
    chk = {(s["m_blk"], s["n_blk"]) for s in segs if s["is_chk"]}
    assert len(chk) == hdr["n_chk_tiles"]
    if kid in (11, 12, 16, 15, 31, 32):
        tiles_n = -(-N // TILE_N[kid])
        assert hdr["n_chk_tiles"] == -(-M // (128 * hdr["cta_group"])) * -(-(tiles_n * 4) // TILE_N[kid])
    acc_stages = 2 if 2 * TILE_N[kid] <= 512 else 1
    assert _simulate(hdr, segs, acc_stages), "circular wait in the schedule"


def test_planner_levels_the_units(ft):
    """The list scheduler + split-K tail removes the wave-quantisation loss where it pays for the fold-in."""
    def makespan(kid, n):
        hdr, segs = ft.debug_schedule(kid, n, n, n, 148)
        work = [0.0] * hdr["units"]
        for s in segs:
            work[s["unit"]] += (0.5 if s["is_chk"] and n == 4096 else 1.0) * (s["kb_end"] - s["kb_begin"]) / hdr["num_kb"]
        return hdr, max(work)
    hdr, t = makespan(21, 4096)   # 256 tiles on 74 pairs: 3.46 waves -> the 34-tile remainder is split in two
    assert (hdr["sk_tiles"], hdr["sk_slices"]) == (34, 2) and abs(t - 3.5) < 1e-9
    hdr, t = makespan(21, 1024)   # 16 short tiles (K = 1024): the partial-sum round trip costs more than it buys
    assert hdr["sk_tiles"] == 0 and t == 1.0
    hdr, t = makespan(21, 8192)   # 13.84 waves: not worth splitting
    assert hdr["sk_tiles"] == 0 and t == 14.0
    hdr, t = makespan(31, 8192)   # ABFT: 32 checksum tiles on top of 1024 data tiles, never K-split
    assert hdr["sk_slices"] == 1 and hdr["n_chk_tiles"] == 32
