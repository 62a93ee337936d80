"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI (ctypes binding of include/bamm_b200.h),
against (a) the committed golden vectors produced by the reference itself and (b) the CPU oracle on seeded inputs.

Tolerances (BASELINE.json north_star): k-mer indices / site indexing bit-exact; posteriors r, log likelihood and
model probabilities v within 1e-5 relative per iteration; final model within 1e-4 after the reference's stop rule.
"""
import os

import numpy as np
import pytest

from util import CASES, Golden

pytestmark = pytest.mark.gpu

RTOL = 1e-5


@pytest.fixture(scope="module")
def capi():
    from bammmotif2_b200 import capi
    capi.load()
    assert capi.device_count() >= 1, "no CUDA device: the product has no CPU fallback"
    return capi


def make_seqset(capi, g):
    pp, pk = capi.kmer_patches(g["pos_codes"], g["pos_kmer"])
    return capi.SeqSet(g["pos_codes"], g["pos_offsets"], g.A, pp, pk)


def assert_rel(a, b, rtol, atol=0.0, what=""):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    err = np.abs(a - b)
    tol = rtol * np.abs(b) + atol
    bad = err > tol
    assert not bad.any(), "%s: %d / %d outside tolerance, worst rel %.3g" % (
        what, int(bad.sum()), bad.size, float((err / np.maximum(np.abs(b), 1e-300)).max()))


@pytest.mark.parametrize("case", CASES)
def test_kmer_index_bit_exact(capi, case):
    g = Golden(case)
    ss = make_seqset(capi, g)
    for K in sorted({0, 1, 2, g.K, min(g.K + 1, 5)}):
        y = ss.get_index(K)
        assert np.array_equal(y.astype(np.uint64), g["pos_kmer"] % np.uint64(g.A ** (K + 1))), "order %d" % K
    n = ss.count_kmers(g.K_bg_model)
    assert np.array_equal(n, g["bg_n"])


@pytest.mark.parametrize("case", CASES)
def test_first_iteration(capi, case):
    g = Golden(case)
    ss = make_seqset(capi, g)
    em = capi.EM(ss, g.W, g.K, g.K_bg_model)
    em.set_model(g["m1_v_init"], g["bg_v"], g["m1_alpha"], g.q)
    llh = em.estep()
    assert np.array_equal(em.s(), g["m1_s_it1"])            # IEEE division on both sides
    r = em.r()
    gr = g["m1_r_it1"]
    assert np.array_equal(r == 0, gr == 0)                  # zero tail and underflow pattern identical
    assert_rel(r, gr, RTOL, what="r")
    assert abs(llh - g["m1_llh"][0]) <= RTOL * abs(g["m1_llh"][0])
    em.mstep()
    n = em.counts()
    assert_rel(n, g["m1_n_it1"], RTOL, atol=1e-9, what="n")
    assert_rel(em.model(), g["m1_v_it1"], RTOL, what="v")


@pytest.mark.parametrize("case", CASES)
def test_optimize_matches_reference(capi, case):
    g = Golden(case)
    ss = make_seqset(capi, g)
    em = capi.EM(ss, g.W, g.K, g.K_bg_model)
    em.set_model(g["m1_v_init"], g["bg_v"], g["m1_alpha"], g.q)
    res = em.optimize(optimize_q=g.optimize_q)
    n = min(res["iterations"], g.iterations)
    assert_rel(res["llh"][:n], g["m1_llh"][:n], RTOL, what="llh trace")
    assert_rel(res["vdiff"][:n], g["m1_vdiff"][:n], 1e-4, atol=2e-4, what="vdiff trace")   # a sum of differences of nearly equal numbers
    assert_rel(res["qtrace"][:n], g["m1_q"][:n], RTOL, what="q trace")
    if res["iterations"] != g.iterations:
        # The reference's second stop rule (llh dropped, EM.cpp:118) fires on the rounding noise of its sequential
        # float sum once the likelihood has plateaued (syn_k4: the reference stops on a -2 ulp step). Any summation
        # order other than the reference's 1-thread one (its own multi-thread runs included) may leave the plateau at
        # another iteration. Accept that ONLY inside the noise band, then compare models at the reference's count.
        llh = g["m1_llh"].astype(np.float64)
        ulp = float(np.spacing(np.float32(abs(llh[n - 1]))))
        assert n > 10 and abs(llh[n - 1] - llh[n - 2]) <= 4 * ulp, "stopped at %d vs %d outside the llh noise band" % (res["iterations"], g.iterations)
        assert g["m1_vdiff"][n - 1] >= 0.01 and abs(g["m1_llh"][-1] - g["m1_llh"][-2]) <= 4 * ulp
        em.set_model(g["m1_v_init"], g["bg_v"], g["m1_alpha"], g.q)
        em.iterate(g.iterations)
        res["v"] = em.model()
    assert_rel(res["v"], g["m1_v_final"], 1e-4, what="final v")
    assert_rel(em.counts(), g["m1_n_it%d" % g.iterations], 1e-4, atol=1e-8, what="final n")
    if "m1_r_it%d" % g.iterations in g:
        assert_rel(em.r(), g["m1_r_it%d" % g.iterations], 1e-4, atol=1e-30, what="final r")


@pytest.mark.parametrize("case", CASES)
def test_scoring_bit_exact(capi, case):
    g = Golden(case)
    ss = make_seqset(capi, g)
    mops, zoops, z = ss.score(g.W, g.K, g.K_bg_model, g["m1_v_final"], g["bg_v"])
    assert np.array_equal(mops, g["m1_score_mops"])
    assert np.array_equal(zoops, g["m1_score_zoops"])
    assert np.array_equal(z, g["m1_score_z"])
    _, zoops2, z2 = ss.score(g.W, g.K, g.K_bg_model, g["m1_v_final"], g["bg_v"], want_mops=False)
    assert np.array_equal(zoops2, zoops) and np.array_equal(z2, z)


def test_subset_fold_against_oracle(capi, oracle):
    """An FDR training fold is an index subset of the resident set (FDR.cpp:49-57)."""
    g = Golden("syn_k3_fdr")
    ss = make_seqset(capi, g)
    nseq = ss.nseq
    cv = 5
    train = np.array([n for n in range(nseq - nseq % cv) if n % cv != 2], np.uint64)
    off = g["pos_offsets"].astype(np.int64)
    kmer = np.concatenate([g["pos_kmer"][off[n]:off[n + 1]] for n in train])
    soff = np.zeros(len(train) + 1, np.uint64)
    soff[1:] = np.cumsum([off[n + 1] - off[n] for n in train])
    ref = oracle.em_optimize(kmer, soff, g.A, g.K, g.W, g.K_bg_model, g["bg_v"], g["m1_alpha"], g["m1_v_init"], g.q)
    em = capi.EM(ss, g.W, g.K, g.K_bg_model, subset=train)
    em.set_model(g["m1_v_init"], g["bg_v"], g["m1_alpha"], g.q)
    res = em.optimize()
    assert res["iterations"] == ref["iterations"]
    assert_rel(res["llh"], ref["llh"], RTOL, what="llh")
    assert_rel(res["v"], ref["v"], 1e-4, what="v")
    test = np.array([n for n in range(nseq - nseq % cv) if n % cv == 2], np.uint64)
    mops, zoops, z = ss.score(g.W, g.K, g.K_bg_model, ref["v"], g["bg_v"], subset=test)
    s = oracle.log_s(ref["v"], g["bg_v"], g.A, g.K, g.K_bg, g.W)
    tk = np.concatenate([g["pos_kmer"][off[n]:off[n + 1]] for n in test])
    toff = np.zeros(len(test) + 1, np.uint64)
    toff[1:] = np.cumsum([off[n + 1] - off[n] for n in test])
    omops, ozoops, oz = oracle.logodds(tk, toff, g.A, g.K, g.W, s)
    assert np.array_equal(mops, omops) and np.array_equal(zoops, ozoops) and np.array_equal(z, oz)


def test_bit_reproducible_run_to_run(capi):
    """Counts are integer (fixed-point) sums: two runs give identical bits (the reference's OpenMP M-step does not)."""
    g = Golden("jund_k2")
    ss = make_seqset(capi, g)
    outs = []
    for _ in range(2):
        em = capi.EM(ss, g.W, g.K, g.K_bg_model)
        em.set_model(g["m1_v_init"], g["bg_v"], g["m1_alpha"], g.q)
        em.iterate(5)
        outs.append((em.model(), em.counts(), em.r()))
        em.close()
    for a, b in zip(outs[0], outs[1]):
        assert np.array_equal(a, b)


def test_stepwise_equals_fused_loop(capi):
    g = Golden("syn_k4")
    ss = make_seqset(capi, g)
    a = capi.EM(ss, g.W, g.K, g.K_bg_model)
    a.set_model(g["m1_v_init"], g["bg_v"], g["m1_alpha"], g.q)
    for _ in range(3):
        a.estep()
        a.mstep()
    b = capi.EM(ss, g.W, g.K, g.K_bg_model)
    b.set_model(g["m1_v_init"], g["bg_v"], g["m1_alpha"], g.q)
    b.iterate(3)
    assert np.array_equal(a.model(), b.model())
    # the two-half multi-GPU form with a no-op exchange is the same iteration
    c = capi.EM(ss, g.W, g.K, g.K_bg_model)
    c.set_model(g["m1_v_init"], g["bg_v"], g["m1_alpha"], g.q)
    for _ in range(3):
        c.estep_local()
        c.mstep_local()
        c.finish_iteration()
    assert np.array_equal(a.model(), c.model())


def test_error_paths(capi):
    g = Golden("syn_k0")
    ss = make_seqset(capi, g)
    with pytest.raises(capi.BammError):
        capi.EM(ss, 64, 0, 0)                     # W out of range
    with pytest.raises(capi.BammError):
        capi.EM(ss, 7, 0, 0, subset=np.array([10 ** 6], np.uint64))
    em = capi.EM(ss, 7, 0, 0)
    with pytest.raises(capi.BammError):
        em.estep()                                # no model yet
    em.set_model(g["m1_v_init"], g["bg_v"], g["m1_alpha"], g.q)
    with pytest.raises(capi.BammError):
        em.mstep()                                # no r yet
    # a sequence shorter than the motif is rejected like the reference filters it (mainBaMM.cpp:75-83)
    codes = np.array([1, 2, 3, 4, 1, 2, 3], np.uint8)
    short = capi.SeqSet(codes, np.array([0, 3, 7], np.uint64), 4)
    with pytest.raises(capi.BammError):
        capi.EM(short, 4, 0, 0)
    # the patch list must be inside the set and strictly increasing (checked on the device)
    with pytest.raises(capi.BammError, match="out of range"):
        capi.SeqSet(codes, np.array([0, 3, 7], np.uint64), 4, np.array([2, 9], np.uint64), np.array([5, 6], np.uint64))
    with pytest.raises(capi.BammError, match="strictly increasing"):
        capi.SeqSet(codes, np.array([0, 3, 7], np.uint64), 4, np.array([4, 4], np.uint64), np.array([5, 6], np.uint64))
    # a negative set needs at least one draw per record
    with pytest.raises(capi.BammError):
        short.sample_negatives(0)
    # a scoring subset outside the set: caught on the device when the whole set is regular (ZOOPS-only call), on the host otherwise
    reg = capi.SeqSet(np.array([1, 2, 3, 4, 1, 2, 3, 4, 4, 3], np.uint8), np.array([0, 5, 10], np.uint64), 4)
    v1 = np.full(4 * 4, 0.25, np.float32)
    with pytest.raises(capi.BammError, match="out of range"):
        reg.score(4, 0, 0, v1, np.full(4, 0.25, np.float32), subset=np.array([1, 2], np.uint64), want_mops=False)
    with pytest.raises((capi.BammError, IndexError)):         # the ctypes wrapper sizes the MOPS buffer from the subset first
        reg.score(4, 0, 0, v1, np.full(4, 0.25, np.float32), subset=np.array([1, 2], np.uint64), want_mops=True)
    _, zo, zz = reg.score(4, 0, 0, v1, np.full(4, 0.25, np.float32), subset=np.array([1, 0, 1], np.uint64), want_mops=False)
    assert len(zo) == 3 and zo[0] == zo[2] and zz[0] == zz[2]
    # empty subset is legal and a no-op
    e0 = capi.EM(ss, 7, 0, 0, subset=np.zeros(0, np.uint64))
    e0.set_model(g["m1_v_init"], g["bg_v"], g["m1_alpha"], g.q)
    assert e0.estep() == 0.0


# ---------------------------------------------------------------------------------------------------------------------
# kernel variants and size-independent properties

def _run_variant(capi, g, env, iters=3):
    """First `iters` iterations with environment switches that select another kernel path for the same arithmetic."""
    import os
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        ss = make_seqset(capi, g)
        em = capi.EM(ss, g.W, g.K, g.K_bg_model)
        em.set_model(g["m1_v_init"], g["bg_v"], g["m1_alpha"], g.q)
        llh = [em.estep()]
        r1 = em.r()
        em.mstep()
        n1 = em.counts()
        em.iterate(iters - 1)
        return dict(llh=llh, r1=r1, n1=n1, v=em.model(), n=em.counts())
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("case", ["jund_k2", "syn_k2_N", "syn_k4", "syn_k3_fdr", "syn_k0", "syn_ss_k1_q"])
def test_kernel_paths_agree(capi, case):
    """Packed path with the active list (default), with an overflowing list (device-side fall-back to the scan
    M-step), without the list, with the M-step's columns split over CTAs (list and scan kernels), without the reduced
    leading context, with table budgets that force more column groups or several column passes of the E-step, and the
    generic index-array path: the M-step sums
    the same fixed-point integers on every path, so counts are BIT-identical wherever the E-step is; E-step variants
    differ only by the association of the column products (1e-5 tolerance)."""
    g = Golden(case)
    base = _run_variant(capi, g, {})
    for env, same_estep in (({"BAMM_LIST_FRAC": "0"}, True), ({"BAMM_LIST_FRAC": "0.000001"}, True),
                            ({"BAMM_M_COLS": "3"}, True), ({"BAMM_M_COLS": "4", "BAMM_LIST_FRAC": "0"}, True),
                            ({"BAMM_M_REPLICAS": "4"}, True), ({"BAMM_M_REPLICAS": "32", "BAMM_M_COLS": "2", "BAMM_LIST_FRAC": "0"}, True),
                            ({"BAMM_NO_REDUCED": "1"}, False),
                            ({"BAMM_TABLE_BYTES": "40000"}, False), ({"BAMM_TABLE_BYTES": "5000"}, False),
                            ({"BAMM_TABLE_BYTES": "9000", "BAMM_M_COLS": "2"}, False), ({"BAMM_NO_PACKED": "1"}, False)):
        alt = _run_variant(capi, g, env)
        if same_estep:
            assert np.array_equal(alt["r1"], base["r1"]), env
            assert np.array_equal(alt["n1"], base["n1"]), env
            assert np.array_equal(alt["v"], base["v"]) and np.array_equal(alt["n"], base["n"]), env
        else:
            assert_rel(alt["r1"], base["r1"], RTOL, what="r %s" % env)
            assert_rel(alt["n1"], base["n1"], RTOL, atol=1e-9, what="n %s" % env)
            assert_rel(alt["v"], base["v"], 1e-4, what="v %s" % env)
        assert_rel(alt["llh"], base["llh"], RTOL, what="llh %s" % env)


def test_properties_at_scale(capi):
    """Size-independent checks on a set far larger than the fixtures (40k x 500 bp, W=20, K=4, both strands):
    posteriors of a sequence sum to 1 - (1-q)/norm < 1, the zero tail is intact, every count column sums to the
    posterior mass of the windows that reach it (truncation rule of EM.cpp:236), the model stays a probability table."""
    from bammmotif2_b200 import synth, hostmodel
    nseq, L0, W, K, Kbg, A, q = 40000, 500, 20, 4, 2, 4, 0.3
    fwd, sites, _ = synth.planted_sequences(99, nseq, L0, W)
    codes = synth.stored_both_strands(fwd)
    ppos, pkmer = synth.middle_n_patches(codes, 99)
    L = codes.shape[1]
    offsets = np.arange(nseq + 1, dtype=np.uint64) * np.uint64(L)
    ss = capi.SeqSet(codes.ravel(), offsets, A, ppos, pkmer)
    vbg = hostmodel.background_from_counts(ss.count_kmers(Kbg), A, Kbg, hostmodel.default_bg_alpha(Kbg))
    alpha = hostmodel.default_motif_alpha(K, W)
    v0 = hostmodel.motif_from_sites(sites, A, K, alpha, vbg)
    em = capi.EM(ss, W, K, Kbg)
    em.set_model(v0, vbg, alpha, q)
    em.iterate(2)
    llh = em.estep()
    r = em.r().reshape(nseq, L)
    LW1 = L - W + 1
    assert np.all(r[:, LW1:] == 0) and np.all(r >= 0)
    mass = r.sum(axis=1, dtype=np.float64)
    assert np.all(mass < 1.0) and np.all(mass > 0.0)
    norm = (1.0 - q) / (1.0 - mass)                                   # from r0 = (1-q)/norm = 1 - sum_i r_i
    assert abs(np.log(norm).sum() - llh) <= 1e-4 * abs(llh)
    em.mstep()
    n = em.counts()
    off = hostmodel.v_offsets(A, K, W)
    nK = n[off[K]:off[K + 1]].reshape(A ** (K + 1), W).astype(np.float64)
    # column j receives r[L-W-p] from every window start p with j <= min(W-1, L-W-p)  <=>  r index i >= j
    tail_mass = r[:, :LW1].sum(axis=0, dtype=np.float64)             # by r index i
    for j in range(W):
        expect = tail_mass[j:].sum()
        assert abs(nK[:, j].sum() - expect) <= 1e-6 * expect, j
    v = em.model()
    assert np.all(np.isfinite(v)) and np.all(v > 0) and np.all(v <= 1.0)   # Motif.h:116 asserts v <= 1 for order 0


@pytest.mark.parametrize("A,K,W,packed", [(6, 5, 6, False), (6, 5, 13, False), (6, 4, 12, False), (6, 3, 8, False), (4, 6, 8, True), (4, 5, 24, True),
                                          (4, 4, 31, True), (4, 2, 32, True), (4, 6, 28, True), (4, 0, 32, True)])
def test_large_tables_against_oracle(capi, oracle, A, K, W, packed):
    """Orders / alphabets beyond the fixtures: the 6-letter alphabet at order 5 (46 656-row table: one column of low words per CTA in shared
    memory, k_mstep_cols; order 4 takes six columns per CTA), order 6 on ACGT (16 384 rows), and a wide order-5 motif whose table forces one
    column per group, and motifs so wide that window + context exceed one 32-base window word (column passes of the
    E-step, column splits of the M-step, each with its own word alignment) — two full iterations against the CPU oracle on seeded random sequences (both strands, with the
    rand()-patched middle N)."""
    rng = np.random.default_rng(7 + A * 100 + K)
    nseq, L0 = 300, 90
    fwd = rng.integers(1, A + 1, size=(nseq, L0), dtype=np.uint8)
    comp = np.array([0, 4, 3, 2, 1, 3, 3], np.uint8)               # ACGTMH -> TGCAGG (Alphabet.cpp:12-27)
    L = 2 * L0 + 1
    codes = np.zeros((nseq, L), np.uint8)
    codes[:, :L0] = fwd
    codes[:, L0 + 1:] = comp[fwd][:, ::-1]
    # reference-style k-mer hashes with an independent draw per (position, digit) for the N
    kmer = np.zeros((nseq, L), np.uint64)
    d = np.where(codes == 0, 0, codes.astype(np.int64) - 1)
    for t in range(11):
        dig = d[:, :L - t].copy()
        col = np.arange(t, L) - t                                   # position of the digit's base
        isN = (col == L0)
        if isN.any():
            dig[:, isN] = rng.integers(0, A, size=(nseq, int(isN.sum())))
        kmer[:, t:] += (dig * (A ** t)).astype(np.uint64)
    kmer = kmer.ravel()
    offsets = np.arange(nseq + 1, dtype=np.uint64) * np.uint64(L)
    pp, pk = capi.kmer_patches(codes.ravel(), kmer)
    ss = capi.SeqSet(codes.ravel(), offsets, A, pp, pk)
    Kbg = min(K, 2)
    from bammmotif2_b200 import hostmodel
    assert np.array_equal(ss.get_index(K).astype(np.uint64), kmer % np.uint64(A ** (K + 1)))
    nb, vbg = oracle.bg_model(kmer, A, Kbg, hostmodel.default_bg_alpha(Kbg))
    assert np.array_equal(ss.count_kmers(Kbg), nb)
    alpha = hostmodel.default_motif_alpha(K, W)
    sites = rng.integers(1, A + 1, size=(200, W), dtype=np.uint8)
    v0 = oracle.motif_from_sites(["".join("ACGTMH"[c - 1] for c in s) for s in sites], A, K, alpha.ravel(), vbg) \
        if False else hostmodel.motif_from_sites(sites, A, K, alpha, vbg)
    em = capi.EM(ss, W, K, Kbg)
    em.set_model(v0, vbg, alpha, 0.3)
    for it in range(2):
        # per-iteration parity: the oracle starts every iteration from the model the device holds (a product of up to 32
        # table entries amplifies last-bit differences of the previous iteration's model beyond the per-iteration tolerance)
        v = em.model()
        s = oracle.linear_s(v, vbg, A, K, Kbg, W)
        r_ref, llh_ref = oracle.estep(kmer, offsets, A, K, W, s, 0.3)
        n_ref = oracle.mstep(kmer, offsets, A, K, W, r_ref)
        v = oracle.update_v(n_ref, alpha.ravel(), vbg, A, K, W, v)
        llh = em.estep()
        # llh is a sum of nseq terms logf(norm_n) of either sign (it nearly cancels for the K=0 case): 1e-5 relative on
        # the sum plus one fp32 rounding (6e-8 of a norm close to 1) per term
        assert abs(llh - llh_ref) <= RTOL * abs(llh_ref) + 1e-7 * nseq
        assert_rel(em.r(), r_ref, RTOL, atol=1e-37, what="r it%d" % it)
        em.mstep()
        assert_rel(em.counts(), n_ref, RTOL, atol=1e-9, what="n it%d" % it)
        assert_rel(em.model(), v, RTOL, what="v it%d" % it)
    # scoring on the same path
    mops, zoops, z = ss.score(W, K, Kbg, v, vbg)
    omops, ozoops, oz = oracle.logodds(kmer, offsets, A, K, W, oracle.log_s(v, vbg, A, K, Kbg, W))
    assert np.array_equal(mops, omops) and np.array_equal(zoops, ozoops) and np.array_equal(z, oz)
    # ZOOPS-only call: prune-and-verify kernel on the packed path, same bits
    _, zoops2, z2 = ss.score(W, K, Kbg, v, vbg, want_mops=False)
    assert np.array_equal(zoops2, ozoops) and np.array_equal(z2, oz)


def test_zoops_pruned_scoring_equals_full_scoring_at_scale(capi):
    """k_score_zoops_packed (column-group bound + exact re-scoring near the running maximum) against the plain kernel on
    30k x 300 bp with a trained order-3 model, both strands (windows over the N included) and a single-stranded sampled
    negative set: maxima and first arg-max positions must be bit-identical."""
    from bammmotif2_b200 import synth, hostmodel
    nseq, L0, W, K, Kbg, A = 30000, 300, 12, 3, 2, 4
    fwd, sites, _ = synth.planted_sequences(5, nseq, L0, W)
    codes = synth.stored_both_strands(fwd)
    ppos, pkmer = synth.middle_n_patches(codes, 5)
    offsets = np.arange(nseq + 1, dtype=np.uint64) * np.uint64(codes.shape[1])
    ss = capi.SeqSet(codes.ravel(), offsets, A, ppos, pkmer)
    vbg = hostmodel.background_from_counts(ss.count_kmers(Kbg), A, Kbg, hostmodel.default_bg_alpha(Kbg))
    alpha = hostmodel.default_motif_alpha(K, W)
    em = capi.EM(ss, W, K, Kbg)
    em.set_model(hostmodel.motif_from_sites(sites, A, K, alpha, vbg), vbg, alpha, 0.3)
    em.iterate(5)
    v = em.model()
    neg = ss.sample_negatives(2)
    for sset in (ss, neg):
        _, zf, pf = sset.score(W, K, Kbg, v, vbg, want_mops=False)
        os.environ["BAMM_NO_ZOOPS_FAST"] = "1"
        try:
            _, zs, ps = sset.score(W, K, Kbg, v, vbg, want_mops=False)
        finally:
            del os.environ["BAMM_NO_ZOOPS_FAST"]
        assert np.array_equal(zf, zs) and np.array_equal(pf, ps)
        sub = np.arange(0, sset.nseq, 3, dtype=np.uint64)
        _, zsub, psub = sset.score(W, K, Kbg, v, vbg, subset=sub, want_mops=False)
        assert np.array_equal(zsub, zf[::3]) and np.array_equal(psub, pf[::3])


def test_device_negative_sampling_of_a_template_subset(capi):
    """Templates given as an index subset of a resident set (the driver filters short sequences before sampling,
    mainBaMM.cpp:75-83) must give the set that the same sequences give as a set of their own (that path is pinned to the
    reference above)."""
    g = Golden("neg_ss")
    off = g["pos_offsets"].astype(np.int64)
    ss = capi.SeqSet(g["pos_codes"], g["pos_offsets"], g.A)
    sub = np.array([n for n in range(ss.nseq) if n % 3 != 1][::-1], np.uint64)          # a subset in another order
    codes = np.concatenate([g["pos_codes"][off[n]:off[n + 1]] for n in sub])
    soff = np.zeros(len(sub) + 1, np.uint64)
    soff[1:] = np.cumsum([off[n + 1] - off[n] for n in sub])
    own = capi.SeqSet(codes, soff, g.A)
    a, b = ss.sample_negatives(7, subset=sub), own.sample_negatives(7)
    assert np.array_equal(a.offsets, b.offsets) and np.array_equal(a.get_codes(), b.get_codes())


@pytest.mark.parametrize("case", ["neg_ragged_N", "syn_k3_fdr"])
def test_device_negative_sampling_in_shards_equals_the_whole(capi, case):
    """Multi-GPU form of the sampler: two shards of the templates, set-wide k-mer counts summed over the shards, the second
    shard starting at its draw offset — the concatenation is the reference's negative set, bit for bit."""
    g = Golden(case)
    off = g["pos_offsets"].astype(np.int64)
    nseq = len(off) - 1
    cut = nseq // 3
    shards = []
    for lo, hi in ((0, cut), (cut, nseq)):
        codes = g["pos_codes"][off[lo]:off[hi]]
        kmer = g["pos_kmer"][off[lo]:off[hi]]
        pp, pk = capi.kmer_patches(codes, kmer)
        shards.append(capi.SeqSet(codes, (off[lo:hi + 1] - off[lo]).astype(np.uint64), g.A, pp, pk))
    counts = shards[0].negative_kmer_counts() + shards[1].negative_kmer_counts()
    from bammmotif2_b200 import sharding
    whole = sharding.negative_kmer_counts((g["pos_kmer"] % np.uint64(g.A ** 3)).astype(np.int64), off, g.A)
    assert np.array_equal(counts.astype(np.int64), whole)                  # device counters == host restatement (gloo test)
    fold = g.meta["mFold"]
    a = shards[0].sample_negatives_shard(fold, 0, counts)
    b = shards[1].sample_negatives_shard(fold, fold * int(off[cut]), counts)
    assert np.array_equal(np.concatenate([a.get_codes(), b.get_codes()]), g["neg_codes"])


def test_device_rand_stream_is_libc_rand(capi):
    """The device re-creation of glibc's rand() (additive lagged-Fibonacci TYPE_3, jump-ahead by polynomial powers) against
    libc itself: the first draws after srand(42) and a block far into the stream reached by running libc there."""
    import ctypes
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(42)
    n_skip, n_take = 300_000, 4000
    ref = np.array([libc.rand() for _ in range(n_skip + n_take)], np.int32)
    assert np.array_equal(capi.rand_stream(42, 0, 5000), ref[:5000])
    assert np.array_equal(capi.rand_stream(42, n_skip, n_take), ref[n_skip:])
    assert np.array_equal(capi.rand_stream(42, 12345, 77), ref[12345:12345 + 77])
    libc.srand(7)
    assert np.array_equal(capi.rand_stream(7, 0, 64), np.array([libc.rand() for _ in range(64)], np.int32))


NEG_CASES = ["syn_k3_fdr", "neg_ragged_N", "neg_ext", "neg_ss"]


@pytest.mark.parametrize("case", NEG_CASES)
def test_device_negative_sampling_bit_exact(capi, case):
    """bamm_seqset_sample_negatives against the negative set the reference itself sampled (SeqGenerator.cpp:188-206 after
    srand(42)): every base identical."""
    g = Golden(case)
    pp, pk = capi.kmer_patches(g["pos_codes"], g["pos_kmer"])
    ss = capi.SeqSet(g["pos_codes"], g["pos_offsets"], g.A, pp, pk)
    fold = g.meta["mFold"]
    neg = ss.sample_negatives(fold)
    assert np.array_equal(neg.offsets, g["neg_offsets"])
    assert np.array_equal(neg.get_codes(), g["neg_codes"])
    # the sampled set is a first-class seqset: its k-mer index equals the reference's hashes of the sampled records
    if "neg_kmer" in g:
        K = min(g.K, 3)
        assert np.array_equal(neg.get_index(K).astype(np.uint64), g["neg_kmer"] % np.uint64(g.A ** (K + 1)))
    neg.close(); ss.close()


def test_device_sort_and_mops_pvalues(capi):
    """Row f-1: bamm_sort_scores against numpy (bit-identical multiset and order) and bamm_mops_pvalues against a restatement
    of ScoreSeqSet::calcPvalues (ScoreSeqSet.cpp:70-126) in numpy float32 — rank search exact, interpolation branch within
    one rounding, exponential-tail branch within a few ulp of expf."""
    rng = np.random.default_rng(11)
    x = rng.normal(0, 3, 300_001).astype(np.float32)
    x[::1000] = x[5]                                              # ties
    assert np.array_equal(capi.sort_scores(x), np.sort(x))
    assert np.array_equal(capi.sort_scores(x, descending=True), np.sort(x)[::-1])
    neg = rng.normal(-6, 3, 200_000).astype(np.float32)
    pos = np.concatenate([rng.normal(-6, 3, 50_000), rng.normal(4, 3, 500), [neg.max() + 1, neg.min() - 1, neg.max(), neg.min()]]).astype(np.float32)
    p, e = capi.mops_pvalues(neg, pos, 777)
    s = np.sort(neg)
    negN = len(s)
    nTop = min(100, negN // 10)
    S_ntop = s[nTop]
    lam = np.float32(0)
    for n in range(nTop):
        lam = np.float32(lam + np.float32(s[n] - S_ntop))
    lam = np.float32(lam / np.float32(nTop))
    FPl = negN - np.searchsorted(s, pos, side="right")
    eps = np.float32(1e-5)
    ref = np.empty(len(pos), np.float32)
    one = FPl == negN
    tail = (~one) & (FPl < 10) & (abs(lam) > eps)
    lin = ~(one | tail)
    ref[one] = 1.0
    ref[tail] = (np.float32(nTop) / np.float32(negN)) * np.exp((-(pos[tail] - S_ntop) / lam).astype(np.float32)).astype(np.float32)
    hi, lo = s[negN - FPl[lin] - 1], s[negN - FPl[lin]]
    ref[lin] = (FPl[lin].astype(np.float32) + (hi - pos[lin] + eps) / (hi - lo + eps)) / np.float32(negN)
    assert one.sum() >= 1 and tail.sum() >= 1 and lin.sum() > 1000
    assert np.array_equal(p[one], ref[one])
    assert np.all(np.abs(p[lin] - ref[lin]) <= 2e-7 * np.abs(ref[lin]))
    assert np.all(np.abs(p[tail] - ref[tail]) <= 2e-6 * np.abs(ref[tail]))
    assert np.array_equal(e, (p * np.float32(777)).astype(np.float32))


def test_mops_pvalues_against_reference(capi):
    """Row f-1 pinned to the reference: bamm_mops_pvalues on the scores the reference fed its own ScoreSeqSet::calcPvalues
    (src/seq_scoring/ScoreSeqSet.cpp:70-126; golden made by oracle/ref_dump with BAMM_DUMP_PVALUES). Rank search, the p = 1 branch
    and the interpolation branch must be bit-identical (same fp32 operations); the exponential tail differs by the device's expf."""
    g = Golden("syn_pval")
    neg, pos, p_ref, e_ref = g["m1_pval_neg_all"], g["m1_score_mops"], g["m1_pval_p"], g["m1_pval_e"]
    posN = len(g["pos_offsets"]) - 1
    p, e = capi.mops_pvalues(neg, pos, posN)
    s = np.sort(neg)
    negN = len(s)
    nTop = min(100, negN // 10)
    lam = np.float32(0)
    for n in range(nTop):
        lam = np.float32(lam + np.float32(s[n] - s[nTop]))
    lam = np.float32(lam / np.float32(nTop))
    FPl = negN - np.searchsorted(s, pos, side="right")
    tail = (FPl != negN) & (FPl < 10) & (abs(lam) > 1e-5)
    assert (~tail).sum() > 1000
    assert np.array_equal(p[~tail], p_ref[~tail]) and np.array_equal(e[~tail], e_ref[~tail])
    if tail.any():      # the reference's rate parameter is fitted to the LOWEST scores (ascending sort, :85-96), so the tail can overflow to inf
        for ours, ref in ((p[tail], p_ref[tail]), (e[tail], e_ref[tail])):
            fin = np.isfinite(ref)
            assert np.array_equal(ours[~fin], ref[~fin])
            assert np.all(np.abs(ours[fin] - ref[fin]) <= 4e-6 * np.abs(ref[fin]))
    # the sort behind it, also in runs merged on the host (the route for vectors the device cannot hold at once)
    import os
    os.environ["BAMM_SORT_RUN"] = "50000"
    try:
        assert np.array_equal(capi.sort_scores(neg), s) and np.array_equal(capi.sort_scores(neg, descending=True), s[::-1])
    finally:
        del os.environ["BAMM_SORT_RUN"]


@pytest.mark.parametrize("case", ["mask_k2", "mask_k3_ss"])
def test_mask_advanced_em_against_reference(capi, oracle, case):
    """Row f-4: bamm_em_mask against the reference's EM::mask (goldens made by ref_dump, f = 0.05). The first phase and the
    selection are exact (same kept windows, same threshold); the iterations agree like the ordinary EM: same iteration
    count, model within 1e-4, counts / r / llh within 1e-5-level tolerances."""
    g = Golden(case)
    ss = make_seqset(capi, g)
    em = capi.EM(ss, g.W, g.K, g.K_bg_model)
    em.set_model(g["m1_v_init"], g["bg_v"], g["m1_alpha"], g.q)
    res = em.mask(f=0.05)
    ref = oracle.em_mask(g["pos_kmer"], g["pos_offsets"], g.A, g.K, g.W, g.K_bg_model, g["bg_v"], g["m1_alpha"], g["m1_v_init"], float(g.q), f=0.05)
    assert res["nkept"] == ref["nkept"] and np.float32(res["cutoff"]) == np.float32(ref["cutoff"])
    assert res["iterations"] == ref["iterations"]
    assert_rel(res["v"], g["m1_mask_v_final"], 1e-4, what="v")
    assert abs(res["llh"] - float(g["m1_mask_llh"][0])) <= 1e-4 * abs(float(g["m1_mask_llh"][0]))
    assert_rel(em.counts(), g["m1_mask_n"], 1e-4, atol=1e-8, what="n")
    r, r_ref = em.r(), g["m1_mask_r"]
    assert np.array_equal(r == 0, r_ref == 0)
    assert_rel(r, r_ref, 2e-4, atol=1e-30, what="r")
