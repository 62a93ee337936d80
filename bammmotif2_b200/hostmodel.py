"""Small host-side model arithmetic used by bench.py and the tests to prepare inputs of the device path:
background model from k-mer counts and the initial motif from binding sites. float32 throughout, in the
reference's operation order, so the results are bit-identical to the reference's (checked in
tests/test_hostmodel.py). The drop-in host side for C++ callers is bammmotif2_b200/host/.
"""
import numpy as np

f32 = np.float32


def bg_offsets(A, K):
    off = [0]
    for k in range(K + 1):
        off.append(off[-1] + A ** (k + 1))
    return off


def v_offsets(A, K, W):
    off = [0]
    for k in range(K + 1):
        off.append(off[-1] + A ** (k + 1) * W)
    return off


def default_bg_alpha(K):
    """reference: src/refinement/Global.cpp:48, 274-278 (1 for order 0, 10 above)"""
    a = np.full(K + 1, 10.0, f32)
    a[0] = 1.0
    return a


def default_motif_alpha(K, W, beta=7.0, gamma=3.0):
    """reference: src/refinement/Global.cpp:36-38, 227-232 and Motif.cpp:44-47 (alpha_k = beta * gamma^k, k>0)"""
    a = np.ones((K + 1, W), f32)
    for k in range(1, K + 1):
        a[k, :] = f32(beta) * np.power(f32(gamma), f32(k), dtype=f32)
    return a


def background_from_counts(n_all, A, K, alpha, interpolate=True):
    """reference: BackgroundModel::calculateV, src/init/BackgroundModel.cpp:441-472. n_all: uint64 counts, all orders."""
    off = bg_offsets(A, K)
    v = np.zeros(off[-1], f32)
    alpha = np.asarray(alpha, f32)
    n0 = n_all[:A]
    base = f32(int(n0.sum()))
    v[:A] = (n0.astype(f32) + alpha[0] * f32(0.25)) / (base + alpha[0])
    for k in range(1, K + 1):
        Yk1, Yk = A ** (k + 1), A ** k
        y = np.arange(Yk1)
        nk = n_all[off[k]:off[k + 1]].astype(f32)
        nk1 = n_all[off[k - 1]:off[k]].astype(f32)
        prior = v[off[k - 1]:off[k]][y % Yk] if interpolate else f32(0.25)
        v[off[k]:off[k + 1]] = (nk + alpha[k] * prior) / (nk1[y // A] + alpha[k])
    return v


def motif_from_sites(sites, A, K, alpha, vbg_all):
    """reference: Motif::initFromBindingSites + calculateV, src/init/Motif.cpp:134-189, 403-428 (no flanks).
    sites: [C][W] codes in 1..A."""
    sites = np.asarray(sites, np.int64) - 1
    C, W = sites.shape
    off = v_offsets(A, K, W)
    alpha = np.asarray(alpha, f32).reshape(K + 1, W)
    n = []
    for k in range(K + 1):
        nk = np.zeros((A ** (k + 1), W), np.int64)
        for j in range(k, W):
            y = np.zeros(C, np.int64)
            for a in range(k + 1):
                y += (A ** a) * sites[:, j - a]
            np.add.at(nk[:, j], y, 1)
        n.append(nk)
    v = np.zeros(off[-1], f32)
    v0 = (n[0].astype(f32) + alpha[0][None, :] * np.asarray(vbg_all[:A], f32)[:, None]) / (f32(C) + alpha[0][None, :])
    vs = [v0.astype(f32)]
    for k in range(1, K + 1):
        Yk1, Yk = A ** (k + 1), A ** k
        y = np.arange(Yk1)
        vk = np.zeros((Yk1, W), f32)
        vk[:, :k] = vs[k - 1][y % Yk, :k]
        num = n[k][:, k:].astype(f32) + alpha[k][None, k:] * vs[k - 1][y % Yk, k:]
        den = n[k - 1][y // A, k - 1:W - 1].astype(f32) + alpha[k][None, k:]
        vk[:, k:] = num / den
        vs.append(vk)
    for k in range(K + 1):
        v[off[k]:off[k + 1]] = vs[k].ravel()
    return v
