"""GPU parity on the BASELINE.json shapes (-m gpu): per-iteration comparison of the CUDA path (through the C ABI) with the
CPU oracle (oracle/bamm_oracle.c, pinned bit-exact to the reference) at sizes where every persistent E-step warp walks
SEVERAL sequences and the benchmarked kernel plans are the ones that run:

  c3 down-sample  40 000 x 500 bp, W=20, K=4, K_bg=2   (the G=8 one-shift plan of the headline configuration)
  c2 full         50 000 x 200 bp, W=12, K=2
  K=5, W=20       10 000 x 500 bp                      (column passes: the pruned E-step carries partial products between them)
  A=6, K=5, W=12  10 000 x 100 bp                      (EXTENDED alphabet: generic index-array path)

each on the default path (c3: the PRUNED E-step — bound pass, exact pass over the candidates, list M-step) and with an active
list too small for anything (BAMM_LIST_FRAC=0.000001: dense E-step + scan M-step); c3 also with a candidate list that
overflows (BAMM_CAND_FRAC=0.000001: bound pass gives up on the device, dense E-step + list M-step). The formulations must
agree in every bit of the model (same tables, same multiplication order, integer normaliser and counts). Reference: EM::EStep / MStep, src/refinement/EM.cpp:139-259; Motif::updateV, src/init/Motif.h:95-136.

Tolerance (BASELINE.json north_star): r, llh, n, v within 1e-5 relative per iteration. The reference accumulates counts and
the log likelihood sequentially in fp32; over 10^7 terms that sum itself is only good to ~1e-4, so the counts and the
likelihood are compared with the oracle's double-accumulation variants (`accumulate_double`, `want_double`: same terms,
exact sum) and the deviation of the sequential fp32 sum from them is printed beside ours.

Absolute floor: the device sums posteriors as 2^-40 fixed point, so a window with r < 2^-41 contributes nothing; over the
~3e5 windows that can fall into one bin of the c2 shape that is at most ~1e-7 of a count (the reference's own fp32 running
sum drops every addend below 2^-24 of the bin's current value, i.e. far more). Counts therefore carry atol = 1e-7 and the
probabilities atol = 1e-7 / alpha_K (a count error d moves v = (n + alpha v')/(N + alpha) by at most d / alpha).
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-5


@pytest.fixture(scope="module")
def capi():
    from bammmotif2_b200 import capi
    capi.load()
    assert capi.device_count() >= 1, "no CUDA device: the product has no CPU fallback"
    return capi


def _assert_rel(a, b, rtol, atol, what):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    err = np.abs(a - b)
    bad = err > rtol * np.abs(b) + atol
    assert not bad.any(), "%s: %d / %d outside tolerance, worst rel %.3g at %d (%.9g vs %.9g)" % (
        what, int(bad.sum()), bad.size, float((err / np.maximum(np.abs(b), 1e-300)).max()), int(np.argmax(err / np.maximum(np.abs(b), 1e-300))),
        a.ravel()[int(np.argmax(bad))], b.ravel()[int(np.argmax(bad))])


def _planted(seed, nseq, L0, W):
    from bammmotif2_b200 import synth
    fwd, sites, _ = synth.planted_sequences(seed, nseq, L0, W)
    codes = synth.stored_both_strands(fwd)
    ppos, pkmer = synth.middle_n_patches(codes, seed)
    kmer = synth.full_kmers(codes, ppos, pkmer)
    return codes, kmer, sites


def _random_extended(seed, nseq, L0, W, A=6):
    """ACGTMH sequences, both strands, reference-style hashes with an independent draw per (position, digit) of the N."""
    rng = np.random.default_rng(seed)
    fwd = rng.integers(1, A + 1, size=(nseq, L0), dtype=np.uint8)
    comp = np.array([0, 4, 3, 2, 1, 3, 3], np.uint8)               # ACGTMH -> TGCAGG (Alphabet.cpp:12-27)
    L = 2 * L0 + 1
    codes = np.zeros((nseq, L), np.uint8)
    codes[:, :L0] = fwd
    codes[:, L0 + 1:] = comp[fwd][:, ::-1]
    kmer = np.zeros((nseq, L), np.uint64)
    d = np.where(codes == 0, 0, codes.astype(np.int64) - 1)
    for t in range(11):
        dig = d[:, :L - t].copy()
        isN = (np.arange(t, L) - t) == L0
        if isN.any():
            dig[:, isN] = rng.integers(0, A, size=(nseq, int(isN.sum())))
        kmer[:, t:] += (dig * (A ** t)).astype(np.uint64)
    # a planted site in half of the sequences keeps the posteriors peaked like real data
    sites = rng.integers(1, A + 1, size=(200, W), dtype=np.uint8)
    return codes, kmer.ravel(), sites


SHAPES = {
    # name: (generator, nseq, L0, W, K, K_bg, A)
    "c3_40k": (_planted, 40_000, 500, 20, 4, 2, 4),
    "c2_full": (_planted, 50_000, 200, 12, 2, 2, 4),
    "k5_w20": (_planted, 10_000, 500, 20, 5, 2, 4),
    "a6_k5": (_random_extended, 10_000, 100, 12, 5, 2, 6),
}


@pytest.mark.parametrize("shape", sorted(SHAPES))
def test_per_iteration_parity_on_baseline_shapes(capi, oracle, shape):
    from bammmotif2_b200 import hostmodel
    gen, nseq, L0, W, K, Kbg, A = SHAPES[shape]
    q = 0.3
    codes, kmer, sites = gen(4242, nseq, L0, W)
    L = codes.shape[1]
    offsets = np.arange(nseq + 1, dtype=np.uint64) * np.uint64(L)
    pp, pk = capi.kmer_patches(codes.ravel(), kmer)
    ss = capi.SeqSet(codes.ravel(), offsets, A, pp, pk)
    assert np.array_equal(ss.get_index(K).astype(np.uint64), kmer % np.uint64(A ** (K + 1)))     # indexing bit-exact
    nb, vbg = oracle.bg_model(kmer, A, Kbg, hostmodel.default_bg_alpha(Kbg))
    assert np.array_equal(ss.count_kmers(Kbg), nb)
    alpha = hostmodel.default_motif_alpha(K, W)
    v0 = hostmodel.motif_from_sites(sites, A, K, alpha, vbg)
    variants = [{}, {"BAMM_LIST_FRAC": "0.000001"}]
    if shape == "c3_40k":
        variants.append({"BAMM_CAND_FRAC": "0.000001"})
    if shape == "k5_w20":
        variants.append({"BAMM_M_NO_HI_GLOBAL": "1"})          # M-step with the high-word table in shared memory (3 column splits instead of 2)
    ems = []
    for env in variants:
        old = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        try:
            em = capi.EM(ss, W, K, Kbg)
        finally:
            for k, v in old.items():
                os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)
        em.set_model(v0, vbg, alpha, q)
        ems.append(em)
    v = v0
    for it in range(2):
        # the oracle's iteration from the model the devices hold (both variants hold bit-identical models, checked below)
        s = oracle.linear_s(v, vbg, A, K, Kbg, W)
        r_ref, llh_f32, llh_ref = oracle.estep(kmer, offsets, A, K, W, s, q, want_double=True)
        n_ref = oracle.mstep(kmer, offsets, A, K, W, r_ref, accumulate_double=True)
        n_f32 = oracle.mstep(kmer, offsets, A, K, W, r_ref)
        v_ref = oracle.update_v(n_ref, alpha.ravel(), vbg, A, K, W, v.copy())
        top = slice(hostmodel.v_offsets(A, K, W)[K], None)
        dev_f32 = float(np.max(np.abs(n_f32[top].astype(np.float64) - n_ref[top]) / np.maximum(np.abs(n_ref[top]), 1e-3)))
        print("%s it%d: sequential fp32 sums of the reference deviate from the exact sums by llh %.2e, n %.2e (relative)" % (
            shape, it + 1, abs(llh_f32 - llh_ref) / abs(llh_ref), dev_f32))
        models = []
        for env, em in zip(variants, ems):
            tag = "%s it%d %s" % (shape, it + 1, env or "default")
            llh = em.estep()
            info = em.estep_info()
            if shape == "k5_w20" and not env:
                assert info["pruned"] and info["passes"] > 1 and not info["dense_ran"], (tag, info)     # pruned path with column passes
                print(tag, info)
            if shape == "c3_40k":       # the benchmarked plan: pruned by default, the two fall-backs when a list is too small
                assert info["pruned"] == ("BAMM_LIST_FRAC" not in env), (tag, info)
                assert info["dense_ran"] == bool(env), (tag, info)
                if not env:
                    assert 0 < info["candidates"] < 0.2 * nseq * L and info["G_bound"] + 3 <= info["G"], (tag, info)
                    print(tag, info)
            assert abs(llh - llh_ref) <= RTOL * abs(llh_ref) + 1e-7 * nseq, tag
            r = em.r()
            assert np.all(r.reshape(nseq, L)[:, L - W + 1:] == 0), tag                  # zero tail (EM.cpp:190-192)
            _assert_rel(r, r_ref, RTOL, 1e-37, "r " + tag)
            em.mstep()
            _assert_rel(em.counts(), n_ref, RTOL, 1e-7, "n " + tag)
            m = em.model()
            _assert_rel(m, v_ref, RTOL, 1e-7 / float(alpha[K].min()), "v " + tag)
            models.append(m)
        for m in models[1:]:
            assert np.array_equal(models[0], m), "pruned / dense E-step and list / scan M-step must give the same bits"
        v = models[0]
    # the fused loop (no r read-back between the steps) reaches the same model bits as the step-wise calls
    em = capi.EM(ss, W, K, Kbg)
    em.set_model(v0, vbg, alpha, q)
    em.iterate(2)
    _assert_rel(em.model(), v, RTOL, 0.0, "fused loop vs step-wise " + shape)
    # ... and r after the fused loop is the r of ITS last E-step (EM.h:48-52 contract of getR())
    _assert_rel(em.r(), ems[0].r(), RTOL, 1e-37, "r after the fused loop " + shape)


@pytest.mark.parametrize("W,K", [(20, 4), (8, 2), (12, 5)])
def test_short_and_ragged_sequences_on_the_pruned_path(capi, oracle, W, K):
    """Sequences of every length from just above W to a few hundred bases, both strands and single-stranded, in one set: the windows
    over the N and the truncated windows (EM.cpp:167) overlap or fill the whole sequence, so k_emasked takes its generic route for
    the warps that hold them, the bound pass has nothing to do for some, and the candidate / active regions are ragged. Compared
    with the oracle per iteration (1e-5) and bit for bit with the dense path (BAMM_NO_SPARSE)."""
    from bammmotif2_b200 import hostmodel
    rng = np.random.default_rng(1000 + W)
    A, Kbg, q, nseq = 4, 2, 0.3, 6000
    pwm = rng.dirichlet(np.full(4, 0.3), size=W)
    sites = np.stack([rng.choice(4, size=200, p=pwm[j]) for j in range(W)], axis=1).astype(np.uint8) + 1
    chunks, kmers, offs = [], [], [0]
    lens = np.concatenate([rng.integers(W // 2 + 1, 2 * W + 3, size=nseq // 2), rng.integers(2 * W, 180, size=nseq - nseq // 2)])
    rng.shuffle(lens)
    for n, L0 in enumerate(lens):
        fwd = rng.integers(1, 5, size=int(L0), dtype=np.uint8)
        if L0 >= W and n % 2 == 0:
            s0 = rng.integers(0, L0 - W + 1)
            fwd[s0:s0 + W] = sites[n % 200]
        single = (n % 7 == 3) and L0 >= W                      # some single-stranded records (no N, no reverse strand)
        codes = fwd if single else np.concatenate([fwd, [0], (5 - fwd)[::-1]]).astype(np.uint8)
        d = np.where(codes == 0, 0, codes.astype(np.int64) - 1)
        km = np.zeros(len(codes), np.uint64)
        for t in range(min(11, len(codes))):
            dig = d[:len(codes) - t].copy()
            if not single:
                isn = np.arange(len(codes) - t) == int(L0)
                dig[isn] = rng.integers(0, 4, size=int(isn.sum()))        # an independent draw per (position, digit) of the N, Sequence.cpp:38
            km[t:] += (dig * (4 ** t)).astype(np.uint64)
        chunks.append(codes); kmers.append(km); offs.append(offs[-1] + len(codes))
    codes, kmer, offsets = np.concatenate(chunks), np.concatenate(kmers), np.array(offs, np.uint64)
    assert np.diff(offsets.astype(np.int64)).min() >= W
    pp, pk = capi.kmer_patches(codes, kmer)
    ss = capi.SeqSet(codes, offsets, A, pp, pk)
    nb, vbg = oracle.bg_model(kmer, A, Kbg, hostmodel.default_bg_alpha(Kbg))
    alpha = hostmodel.default_motif_alpha(K, W)
    v0 = hostmodel.motif_from_sites(sites, A, K, alpha, vbg)
    ems = []
    for env in ({}, {"BAMM_NO_SPARSE": "1"}):
        os.environ.update(env)
        try:
            em = capi.EM(ss, W, K, Kbg)
            em.set_model(v0, vbg, alpha, q)
        finally:
            for k in env:
                os.environ.pop(k, None)
        ems.append(em)
    v = v0
    for it in range(2):
        s = oracle.linear_s(v, vbg, A, K, Kbg, W)
        r_ref, _, llh_ref = oracle.estep(kmer, offsets, A, K, W, s, q, want_double=True)
        n_ref = oracle.mstep(kmer, offsets, A, K, W, r_ref, accumulate_double=True)
        v_ref = oracle.update_v(n_ref, alpha.ravel(), vbg, A, K, W, v.copy())
        out = []
        for em, name in zip(ems, ("pruned", "dense")):
            llh = em.estep()
            info = em.estep_info()
            # W = 8 is so unspecific that more than 42 % of a warp's windows can pass the bound: the device then falls back to the
            # dense kernel for the iteration (and must still agree in every bit)
            assert info["pruned"] == (name == "pruned") and (name == "dense" or W < 12 or not info["dense_ran"]), (name, info)
            assert abs(llh - llh_ref) <= RTOL * abs(llh_ref) + 1e-7 * nseq, (name, it)
            r = em.r()
            _assert_rel(r, r_ref, RTOL, 1e-37, "r %s it%d" % (name, it + 1))
            em.mstep()
            _assert_rel(em.counts(), n_ref, RTOL, 1e-7, "n %s it%d" % (name, it + 1))
            m = em.model()
            _assert_rel(m, v_ref, RTOL, 1e-7 / float(alpha[K].min()), "v %s it%d" % (name, it + 1))
            out.append((llh, r, m))
        assert out[0][0] == out[1][0] and np.array_equal(out[0][1], out[1][1]) and np.array_equal(out[0][2], out[1][2]), "pruned and dense path must agree in every bit"
        v = out[0][2]
