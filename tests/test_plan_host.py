"""Host logic of the packed E-step, no GPU: the column-group planner (csrc/capi_em.inl make_group_plan / plan_passes) through
bamm_plan_describe. The plan replaces the per-position products of EM::EStep (reference src/refinement/EM.cpp:149-196) by
one table lookup per column group; here its invariants are checked for every width / order / budget, and the bit-field
extraction the kernels do with (kd, shift, shift2, mask4) is replayed in integers against the definition of a group's
table index (bases lo .. last column of the group, last base in the lowest two bits)."""
import numpy as np
import pytest

from bammmotif2_b200 import capi

BUDGETS = [227 * 1024, 48 * 1024, 9000, 5000]


def ctx(j, K, K_bg, reduced):
    c = min(j, K) if reduced else K
    return max(c, K_bg)


def window_word(bases, p, kd):
    """64-bit word of the 32 bases from p-kd (zero outside the sequence, as the packed stream's pad words give)."""
    w = 0
    for i in range(32):
        q = p - kd + i
        b = int(bases[q]) if 0 <= q < len(bases) else 0
        w |= b << (62 - 2 * i)
    return w


def extract(w, g, fast):
    if fast:
        assert g["shift"] <= 31
        return ((w >> g["shift"]) & 0xffffffff) & g["mask4"]
    v = (w >> min(g["shift"], 32)) & 0xffffffff         # funnel shift with a clamped count
    return (v >> g["shift2"]) & g["mask4"]


@pytest.mark.parametrize("budget", BUDGETS)
@pytest.mark.parametrize("reduced", [True, False])
def test_plan_invariants_and_extraction(budget, reduced):
    rng = np.random.default_rng(5)
    nplans = 0
    for K in range(0, 7):
        for K_bg in (0, 2):
            for W in range(1, 33):
                passes = capi.plan_describe(W, K, K_bg, reduced, budget)
                K_bg = min(K_bg, K)                       # the background order an EM object uses (reference EM.cpp:23)
                if not passes:
                    # even one column per group does not fit: the widest single-column table is 4^(max ctx + 1) floats
                    need = 4 * 4 ** (max(ctx(j, K, K_bg, reduced) for j in range(W)) + 1)
                    assert need > budget or K + W > 32, (W, K, K_bg, need)
                    continue
                nplans += 1
                # passes tile [0, W) in order; groups tile the pass
                assert passes[0]["ca"] == 0 and passes[-1]["cb"] == W
                assert passes[0]["first"] and passes[-1]["last"]
                bases = rng.integers(0, 4, size=W + 40)
                for i, ps in enumerate(passes):
                    if i:
                        assert ps["ca"] == passes[i - 1]["cb"] and not ps["first"]
                    if i + 1 < len(passes):
                        assert not ps["last"]
                    assert 1 <= ps["G"] <= 16
                    assert ps["table_bytes"] <= budget
                    assert ps["kd"] >= K - ps["ca"] and ps["kd"] <= 31 - ps["cb"]
                    col, base = ps["ca"], 0
                    for g in ps["groups"]:
                        assert g["col0"] == col and g["ncol"] >= 1
                        hi = g["col0"] + g["ncol"] - 1
                        lo = min(j - ctx(j, K, K_bg, reduced) for j in range(g["col0"], hi + 1))
                        assert g["lo"] == lo
                        nb = hi - lo + 1
                        assert g["base"] == base and g["mask4"] == ((4 ** nb - 1) << 2)
                        assert g["colmask"] == sum(1 << j for j in range(g["col0"], hi + 1))
                        base += 4 * 4 ** nb
                        col = hi + 1
                        # the extracted byte offset is 4 x the tuple index, for windows at the start, inside and at the end
                        for p in (0, 1, K, 7, len(bases) - W):
                            w = window_word(bases, p, ps["kd"])
                            idx = 0
                            for q in range(p + lo, p + hi + 1):
                                idx = idx * 4 + (int(bases[q]) if q >= 0 else 0)
                            assert extract(w, g, ps["fast"]) == 4 * idx, (W, K, K_bg, p, g)
                    assert col == ps["cb"] and base == ps["table_bytes"]
    assert nplans > 100


def test_plan_of_the_bench_configurations():
    """The shapes bench.py runs: c3 (W=20, K=4) needs 8 lookups per window, c2 (W=12, K=2) 3, all one-shift."""
    c3 = capi.plan_describe(20, 4, 2, True, 227 * 1024)
    assert len(c3) == 1 and c3[0]["G"] == 8 and c3[0]["fast"]
    c2 = capi.plan_describe(12, 2, 2, True, 227 * 1024)
    assert len(c2) == 1 and c2[0]["fast"]
    # fewer bytes never gives fewer lookups
    lookups = [sum(p["G"] for p in capi.plan_describe(20, 4, 2, True, b)) for b in BUDGETS]
    assert lookups == sorted(lookups)


def test_plan_rejects_bad_arguments():
    with pytest.raises(capi.BammError):
        capi.plan_describe(0, 2)
    with pytest.raises(capi.BammError):
        capi.plan_describe(33, 2)
