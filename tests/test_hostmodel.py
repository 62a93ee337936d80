"""CPU suite: host-side helpers (bench / test plumbing) against the golden vectors and the oracle."""
import numpy as np
import pytest

from bammmotif2_b200 import hostmodel, synth
from util import CASES, Golden, encode_text


@pytest.mark.parametrize("case", CASES)
def test_background_and_site_init_bit_exact(oracle, case):
    g = Golden(case)
    v = hostmodel.background_from_counts(g["bg_n"], g.A, g.K_bg_model, hostmodel.default_bg_alpha(g.K_bg_model))
    assert np.array_equal(v, g["bg_v"])
    alpha = hostmodel.default_motif_alpha(g.K, g.W)
    assert np.array_equal(alpha.ravel(), g["m1_alpha"])
    _, b2c, _ = oracle.alphabet_tables(g.alphabet)
    sites = np.stack([encode_text(s, b2c) for s in g.sites()])
    v0 = hostmodel.motif_from_sites(sites, g.A, g.K, alpha, g["bg_v"])
    assert np.array_equal(v0, g["m1_v_init"])


def test_synth_layout_matches_reference_encoding(oracle):
    """stored_both_strands / full_kmers reproduce Sequence.cpp's layout and hash everywhere except at the
    positions whose hash depends on rand() (those are exactly the patch list)."""
    fwd, sites, _ = synth.planted_sequences(5, 20, 30, 6)
    codes = synth.stored_both_strands(fwd)
    ppos, pkmer = synth.middle_n_patches(codes, 5)
    kmer = synth.full_kmers(codes, ppos, pkmer)
    oracle.srand(1)
    ocodes, okmer, ooff = oracle.encode_sequences([f for f in fwd], "STANDARD", False)
    assert np.array_equal(ocodes, codes.ravel())
    mask = np.ones(len(kmer), bool)
    mask[ppos.astype(np.int64)] = False
    assert np.array_equal(kmer[mask], okmer[mask])
    # patched positions: identical outside the digit that belongs to the N
    L, L0 = codes.shape[1], fwd.shape[1]
    for a in range(11):
        i = L0 + a
        sel = np.arange(20) * L + i
        assert np.array_equal(kmer[sel] // 4 ** (a + 1), okmer[sel] // 4 ** (a + 1))
        assert np.array_equal(kmer[sel] % 4 ** a, okmer[sel] % 4 ** a)
    assert sites.min() >= 1 and sites.max() <= 4 and fwd.min() >= 1 and fwd.max() <= 4


def test_fasta_roundtrip(tmp_path):
    from util import parse_fasta
    fwd, sites, _ = synth.planted_sequences(9, 7, 25, 5)
    p = tmp_path / "x.fasta"
    synth.write_fasta(str(p), fwd)
    recs = parse_fasta(open(p).read())
    assert len(recs) == 7
    lut = {c: i + 1 for i, c in enumerate("ACGT")}
    assert all(np.array_equal([lut[c] for c in s], fwd[n]) for n, (_, s) in enumerate(recs))
