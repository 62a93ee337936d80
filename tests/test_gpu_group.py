"""Device groups (-m gpu, needs two devices): ONE process, the sequences of an EM object split over the devices of
bamm_set_device_group, counts exchanged over NVLink peer memory. The results must equal the single-device run in every bit
(reference: the OpenMP loops over the sequences in EM::EStep / MStep, src/refinement/EM.cpp:148-149, 230)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from bammmotif2_b200 import capi
    capi.load()
    if capi.device_count() < 2:
        pytest.skip("needs two devices")
    yield capi
    capi.set_device_group([])


def _setup(capi, nseq=12_000, L0=200, W=12, K=2, Kbg=2, seed=77):
    from bammmotif2_b200 import synth, hostmodel
    fwd, sites, _ = synth.planted_sequences(seed, nseq, L0, W)
    codes = synth.stored_both_strands(fwd)
    ppos, pkmer = synth.middle_n_patches(codes, seed)
    L = codes.shape[1]
    offsets = np.arange(nseq + 1, dtype=np.uint64) * np.uint64(L)
    ss = capi.SeqSet(codes.ravel(), offsets, 4, ppos, pkmer)
    vbg = hostmodel.background_from_counts(ss.count_kmers(Kbg), 4, Kbg, hostmodel.default_bg_alpha(Kbg))
    alpha = hostmodel.default_motif_alpha(K, W)
    v0 = hostmodel.motif_from_sites(sites, 4, K, alpha, vbg)
    return ss, (W, K, Kbg), (v0, vbg, alpha)


@pytest.mark.parametrize("shape", [dict(W=12, K=2), dict(W=20, K=4, L0=300, nseq=16_000)])
def test_group_equals_one_device(capi, shape):
    ss, (W, K, Kbg), (v0, vbg, alpha) = _setup(capi, **shape)
    runs = {}
    for name, devs in (("one", []), ("two", [0, 1])):
        capi.set_device_group(devs)
        # whole set and an interleaved subset (what an FDR fold trains on)
        for sub_name, subset in (("all", None), ("fold", np.array([i for i in range(ss.nseq) if i % 5 != 2], np.uint64))):
            em = capi.EM(ss, W, K, Kbg, subset=subset)
            em.set_model(v0, vbg, alpha, 0.3)
            llh1 = em.estep()
            r1 = em.r()
            em.mstep()
            m1 = em.model()
            o = em.optimize(optimize_q=True, epsilon=0.01, max_iter=30)
            runs[(name, sub_name)] = dict(llh1=llh1, r1=r1, m1=m1, it=o["iterations"], llh=o["llh"], vd=o["vdiff"], qt=o["qtrace"], model=o["v"],
                                          counts=em.counts(), r=em.r(), q=o["q"])
            em.close()
    capi.set_device_group([])
    for sub_name in ("all", "fold"):
        a, b = runs[("one", sub_name)], runs[("two", sub_name)]
        assert a["it"] == b["it"] and a["llh1"] == b["llh1"]
        for k in ("r1", "m1", "llh", "vd", "qt", "model", "counts", "r"):
            assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), (sub_name, k)
        assert a["q"] == b["q"]
