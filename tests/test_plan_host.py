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


# ---- bound plan of the pruned E-step (csrc/capi_em.inl make_bound_plan, csrc/packed.cuh bound_levels_cta / k_make_bound_tables) --------
def bf16_up(x):
    """Smallest bfloat16 >= x (x >= 0), as float32 — what k_make_bound_tables stores."""
    b = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    b = np.where(b & 0xffff, (b | 0xffff) + 1, b) >> 16
    return (b << 16).astype(np.uint32).view(np.float32)


@pytest.mark.parametrize("K,K_bg", [(0, 0), (1, 1), (2, 2), (4, 2), (5, 2)])
def test_bound_plan_invariants_and_bound_property(K, K_bg):
    """For every width: the groups tile the columns, their tables fit the budget, the (kd, shift, mask4) extraction of the kernel
    yields the bases lo .. hi+1 of a group — and, with tables built as the device builds them (maximum of s over the context bases
    left of the group, product over the group's columns, rounded up to bfloat16), the product of the looked-up entries is an
    upper bound of the exact window product (reference: the per-window product of EM::EStep, src/refinement/EM.cpp:167-176) for
    window p (low half) and window p+1 (high half)."""
    rng = np.random.default_rng(17 + K)
    Yn = 4 ** (K + 1)
    seen = 0
    for W in (6, 8, 12, 13, 20, 24, 30):
        plan = capi.bound_plan_describe(W, K, K_bg)
        assert plan is not None, W
        G, kd, groups = plan["G"], plan["kd"], plan["groups"]
        assert sum(4 * 4 ** (g["ncol"] + g["col0"] - g["lo"] + 1) for g in groups) == plan["table_bytes"] <= 232448
        assert groups[0]["col0"] == 0 and groups[-1]["col0"] + groups[-1]["ncol"] == W
        for a, b in zip(groups, groups[1:]):
            assert b["col0"] == a["col0"] + a["ncol"] and a["lo"] <= a["col0"] and b["lo"] <= b["col0"]
        assert -groups[0]["lo"] <= kd <= 31 - (W + 1)
        # a model with strong context dependence, reduced leading columns like Motif::updateV produces (Motif.h:126-128)
        s = rng.lognormal(0.0, 1.0, size=(W, Yn)).astype(np.float32)
        for j in range(min(K, W)):
            period = 4 ** (max(j, K_bg) + 1)
            s[j] = s[j][np.arange(Yn) % period]
        bases = rng.integers(0, 4, size=400)
        def y_at(i, nb):                      # k-mer index of the nb newest bases ending at position i (zero before the sequence)
            return sum((int(bases[i - t]) if i - t >= 0 else 0) << (2 * t) for t in range(nb))
        tabs = []
        for g in groups:
            hi, lo = g["col0"] + g["ncol"] - 1, g["lo"]
            T = hi - lo + 1
            z = np.arange(4 ** T)
            f = np.ones(4 ** T, np.float32)
            for j in range(g["col0"], hi + 1):
                avail = min(j - lo + 1, K + 1)
                U = s[j].reshape(4 ** (K + 1 - avail), 4 ** avail).max(axis=0)          # maximum over the missing (older) context bases
                f = (f * U[(z >> (2 * (hi - j))) & (4 ** avail - 1)]).astype(np.float32)
            z2 = np.arange(4 ** (T + 1))
            tabs.append((bf16_up(f[z2 >> 2]), bf16_up(f[z2 & (4 ** T - 1)])))
        for p in range(0, 300, 7):
            w = window_word(bases, p, kd)
            b0 = b1 = np.float32(1.0)
            for g, (lo_half, hi_half) in zip(groups, tabs):
                hi = g["col0"] + g["ncol"] - 1
                idx = extract(w, g, plan["fast"]) >> 2
                T1 = hi - g["lo"] + 2
                want = sum((int(bases[p + g["lo"] + t]) if 0 <= p + g["lo"] + t else 0) << (2 * (T1 - 1 - t)) for t in range(T1))
                assert idx == want, (W, K, p, g)
                b0, b1 = np.float32(b0 * lo_half[idx]), np.float32(b1 * hi_half[idx])
            for q, b in ((p, b0), (p + 1, b1)):
                exact = np.float64(1.0)
                for j in range(W):
                    exact *= np.float64(s[j][y_at(q + j, K + 1)])
                assert float(b) >= exact * (1 - 1e-4), (W, K, q, float(b), exact)
                seen += 1
    assert seen > 500
