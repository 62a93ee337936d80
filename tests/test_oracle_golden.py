"""CPU suite: pins oracle/bamm_oracle.c against vectors produced by the reference itself
(oracle/_ref/ref_dump via oracle/make_golden.py). Integer work and the sequential float arithmetic
must be BIT-EXACT: the oracle restates the reference operation by operation."""
import numpy as np
import pytest

from util import CASES, Golden, encode_text


def _encode(orc, g):
    A, b2c, _ = orc.alphabet_tables(g.alphabet)
    assert A == g.A
    encs = [encode_text(s, b2c) for _, s in g.fasta_records()]
    orc.srand(42)                      # reference: mainBaMM.cpp:22
    return orc.encode_sequences(encs, g.alphabet, g.ss)


@pytest.mark.parametrize("case", CASES)
def test_encoding_bit_exact(oracle, case):
    g = Golden(case)
    codes, kmer, offsets = _encode(oracle, g)
    assert np.array_equal(offsets, g["pos_offsets"])
    assert np.array_equal(codes, g["pos_codes"])
    assert np.array_equal(kmer, g["pos_kmer"])


@pytest.mark.parametrize("case", CASES)
def test_background_and_init(oracle, case):
    g = Golden(case)
    n, v = oracle.bg_model(g["pos_kmer"], g.A, g.K_bg_model, g.bg_alpha())
    assert np.array_equal(n, g["bg_n"])
    assert np.array_equal(v, g["bg_v"])
    _, b2c, _ = oracle.alphabet_tables(g.alphabet)
    sites = np.stack([encode_text(s, b2c) for s in g.sites()])
    alpha = g["m1_alpha"].reshape(g.K + 1, g.W)
    v0 = oracle.motif_from_sites(sites, g.A, g.K, alpha, g["bg_v"])
    assert np.array_equal(v0, g["m1_v_init"])


@pytest.mark.parametrize("case", CASES)
def test_first_iteration_bit_exact(oracle, case):
    g = Golden(case)
    kmer, off = g["pos_kmer"], g["pos_offsets"]
    s = oracle.linear_s(g["m1_v_init"], g["bg_v"], g.A, g.K, g.K_bg, g.W)
    assert np.array_equal(s, g["m1_s_it1"])
    r, llh = oracle.estep(kmer, off, g.A, g.K, g.W, s, g.q)
    assert np.array_equal(r, g["m1_r_it1"])
    assert np.float32(llh) == g["m1_llh"][0]
    n_all = oracle.mstep(kmer, off, g.A, g.K, g.W, r)
    assert np.array_equal(n_all, g["m1_n_it1"])
    v = g["m1_v_init"].copy()
    oracle.update_v(n_all, g["m1_alpha"], g["bg_v"], g.A, g.K, g.W, v)
    assert np.array_equal(v, g["m1_v_it1"])


@pytest.mark.parametrize("case", CASES)
def test_optimize_loop_bit_exact(oracle, case):
    g = Golden(case)
    res = oracle.em_optimize(g["pos_kmer"], g["pos_offsets"], g.A, g.K, g.W, g.K_bg_model, g["bg_v"], g["m1_alpha"],
                             g["m1_v_init"], g.q, g.optimize_q)
    assert res["iterations"] == g.iterations
    assert np.array_equal(res["llh"], g["m1_llh"])
    assert np.array_equal(res["vdiff"], g["m1_vdiff"])
    assert np.array_equal(res["qtrace"], g["m1_q"])
    assert np.array_equal(res["v"], g["m1_v_final"])
    assert np.array_equal(res["v"], g["m1_opt_v_final"])      # reference's own optimize() loop
    assert np.array_equal(res["n"], g["m1_n_it%d" % g.iterations])
    p = oracle.calculate_p(res["v"], g["bg_v"], g.K_bg_model, g.A, g.K, g.W)
    assert np.array_equal(p, g["m1_p_final"])


@pytest.mark.parametrize("case", CASES)
def test_scoring_bit_exact(oracle, case):
    g = Golden(case)
    s = oracle.log_s(g["m1_v_final"], g["bg_v"], g.A, g.K, g.K_bg, g.W)
    assert np.array_equal(s, g["m1_score_logs"])
    mops, zoops, z = oracle.logodds(g["pos_kmer"], g["pos_offsets"], g.A, g.K, g.W, s)
    assert np.array_equal(mops, g["m1_score_mops"])
    assert np.array_equal(zoops, g["m1_score_zoops"])
    assert np.array_equal(z, g["m1_score_z"])


def test_double_accumulation_close_to_float(oracle):
    """The double-accumulating variants (used for the at-scale checks) agree with the float path at small scale."""
    g = Golden("syn_k4")
    s = oracle.linear_s(g["m1_v_init"], g["bg_v"], g.A, g.K, g.K_bg, g.W)
    r, llh, llhd = oracle.estep(g["pos_kmer"], g["pos_offsets"], g.A, g.K, g.W, s, g.q, want_double=True)
    assert abs(llh - llhd) <= 1e-5 * abs(llhd)
    nf = oracle.mstep(g["pos_kmer"], g["pos_offsets"], g.A, g.K, g.W, r)
    nd = oracle.mstep(g["pos_kmer"], g["pos_offsets"], g.A, g.K, g.W, r, accumulate_double=True)
    assert np.allclose(nf, nd, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("case", ["mask_k2", "mask_k3_ss"])
def test_oracle_mask_matches_reference_bit_exact(case):
    """orc_em_mask (restatement of EM::mask, EM.cpp:261-503, --advanceEM with the driver's default f = 0.05) against the
    reference's own mask() run single-threaded: final model, responsibilities, counts and log likelihood, every bit."""
    from oracle import oracle as orc
    g = Golden(case)
    res = orc.em_mask(g["pos_kmer"], g["pos_offsets"], g.A, g.K, g.W, g.K_bg_model, g["bg_v"], g["m1_alpha"], g["m1_v_init"], float(g.q), f=0.05)
    assert np.array_equal(res["v"], g["m1_mask_v_final"])
    assert np.array_equal(res["r"], g["m1_mask_r"])
    assert np.array_equal(res["n"], g["m1_mask_n"])
    assert np.float32(res["llh"]) == g["m1_mask_llh"][0]
