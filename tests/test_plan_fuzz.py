"""Randomised (hypothesis) check of the work planner on the CPU: for arbitrary shapes, variants, SM counts and knobs the plan
covers every (tile, k-block) exactly once, respects the global item order and cannot deadlock (same checks as
tests/test_schedule.py, which enumerates fixed shapes)."""
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from test_schedule import PAIR_IDS, TILE_N, _simulate


@settings(max_examples=120, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(kid=st.sampled_from([1, 2, 3, 4, 6, 5, 21, 22, 11, 12, 13, 14, 16, 15, 31, 32]),
       m4=st.integers(1, 3000), n4=st.integers(1, 3000), K=st.integers(1, 20000),
       sms=st.sampled_from([148, 132, 20, 8, 2]), splitk=st.sampled_from([-1, 0, 2, 3, 4]))
def test_random_plans(ft, kid, m4, n4, K, sms, splitk):
    M, N = 4 * m4, 4 * n4
    bn = TILE_N[kid]
    cg = 2 if kid in PAIR_IDS else 1
    if (-(-M // (128 * cg))) * (-(-N // bn)) > 20000:   # keep one example cheap
        M = min(M, 128 * cg * 100)
        N = min(N, bn * 100)
    try:
        ft.debug_set("splitk", splitk)
        hdr, segs = ft.debug_schedule(kid, M, N, K, sms)
    finally:
        ft.debug_set("splitk", -1)
    num_kb = -(-K // 32)
    assert hdr["num_kb"] == num_kb and hdr["cta_group"] == cg
    cover = {}
    for s in segs:
        assert 0 <= s["kb_begin"] < s["kb_end"] <= num_kb and 0 <= s["unit"] < hdr["units"]
        cover.setdefault(s["tile"], []).append((s["kb_begin"], s["kb_end"], s["kind"]))
    assert sorted(cover) == list(range(hdr["num_tiles"]))
    S_chk = max(1, hdr["chk_slices"])
    for t, pieces in cover.items():
        pieces.sort()
        if t < hdr["n_chk_tiles"] and S_chk > 1:
            sl = t // (hdr["n_chk_tiles"] // S_chk)
            assert [p[:2] for p in pieces] == [(num_kb * sl // S_chk, num_kb * (sl + 1) // S_chk)]
            continue
        assert pieces[0][0] == 0 and pieces[-1][1] == num_kb
        assert all(a[1] == b[0] for a, b in zip(pieces, pieces[1:]))
        kinds = [p[2] for p in pieces]
        assert kinds in ([0], [6]) or kinds == [1] + [3] * (len(kinds) - 2) + [2]
    assert _simulate(hdr, segs, 2), "circular wait in the schedule"
