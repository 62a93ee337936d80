"""ctypes front-end of the CPU oracle (oracle/bamm_oracle.c). TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; nothing under bammmotif2_b200/ does. Each wrapper names the reference
function (file:line) its C counterpart restates.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "liboracle.so")
_lib = None

ALPHABET_TYPES = {"STANDARD": 0, "METHYLC": 1, "HYDROXYMETHYLC": 2, "EXTENDED": 3}


def build(force=False):
    """Compile oracle/bamm_oracle.c -> oracle/_ref/liboracle.so (gcc, a second or two)."""
    src = os.path.join(_HERE, "bamm_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_model_size.restype = C.c_uint64
        _lib.orc_bg_size.restype = C.c_uint64
        _lib.orc_sequence_build.restype = C.c_uint64
        _lib.orc_estep.restype = C.c_float
        _lib.orc_optimize_q.restype = C.c_float
        _lib.orc_em_iteration_omp.restype = C.c_float
        _lib.orc_em_optimize.restype = C.c_int
        _lib.orc_rand.restype = C.c_int
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def model_size(A, K, W):
    return int(lib().orc_model_size(A, K, W))


def bg_size(A, K):
    return int(lib().orc_bg_size(A, K))


def v_offset(A, k, W):
    return sum(A ** (kk + 1) * W for kk in range(k))


def bg_offset(A, k):
    return sum(A ** (kk + 1) for kk in range(k))


def alphabet_tables(alphabet="STANDARD"):
    """reference: src/init/Alphabet.cpp:10-55"""
    b2c = np.zeros(128, np.uint8)
    c2c = np.zeros(128, np.uint8)
    size = lib().orc_alphabet_tables(ALPHABET_TYPES[alphabet], _p(b2c, C.c_uint8), _p(c2c, C.c_uint8))
    assert size > 0
    return size, b2c, c2c


def srand(seed):
    lib().orc_srand(C.c_uint(seed))


def encode_sequences(enc_list, alphabet="STANDARD", single_strand=False):
    """reference: src/init/Sequence.cpp:4-43,91-99. enc_list: list of uint8 code arrays (input strand).
    Consumes libc rand() exactly as the reference does (call srand first). Returns codes, kmer, offsets."""
    A, _, c2c = alphabet_tables(alphabet)
    Ls = [(len(e) if single_strand else 2 * len(e) + 1) for e in enc_list]
    offsets = np.zeros(len(enc_list) + 1, np.uint64)
    offsets[1:] = np.cumsum(Ls)
    codes = np.zeros(int(offsets[-1]), np.uint8)
    kmer = np.zeros(int(offsets[-1]), np.uint64)
    for n, e in enumerate(enc_list):
        e = np.ascontiguousarray(e, np.uint8)
        o = int(offsets[n])
        L = lib().orc_sequence_build(_p(e, C.c_uint8), C.c_uint64(len(e)), A, _p(c2c, C.c_uint8), int(single_strand),
                                     C.cast(codes.ctypes.data + o, C.POINTER(C.c_uint8)),
                                     C.cast(kmer.ctypes.data + 8 * o, C.POINTER(C.c_uint64)))
        assert L == Ls[n]
    return codes, kmer, offsets


def bg_model(kmer, A, K, alpha, interpolate=True):
    """reference: src/init/BackgroundModel.cpp:26-42, 441-472"""
    alpha = np.ascontiguousarray(alpha, np.float32)
    n = np.zeros(bg_size(A, K), np.uint64)
    v = np.zeros(bg_size(A, K), np.float32)
    lib().orc_bg_model(_p(kmer, C.c_uint64), C.c_uint64(len(kmer)), A, K, _p(alpha, C.c_float), int(interpolate),
                       _p(n, C.c_uint64), _p(v, C.c_float))
    return n, v


def motif_from_sites(sites, A, K, alpha, vbg_all):
    """reference: src/init/Motif.cpp:134-189, 403-428. sites: [C][W] codes 1..A; alpha [K+1][W]."""
    sites = np.ascontiguousarray(sites, np.uint8)
    Cn, W = sites.shape
    alpha = np.ascontiguousarray(alpha, np.float32)
    v = np.zeros(model_size(A, K, W), np.float32)
    lib().orc_motif_from_sites(_p(sites, C.c_uint8), C.c_uint64(Cn), W, A, K, _p(alpha, C.c_float),
                               _p(vbg_all, C.c_float), _p(v, C.c_float))
    return v


def update_v(n_all, alpha, vbg_all, A, K, W, v_all):
    """reference: src/init/Motif.h:95-136 (in place on v_all)"""
    lib().orc_update_v(_p(n_all, C.c_float), _p(alpha, C.c_float), _p(vbg_all, C.c_float), A, K, W, _p(v_all, C.c_float))
    return v_all


def calculate_p(v_all, vbg_all, k_bg, A, K, W):
    """reference: src/init/Motif.cpp:430-469"""
    p = np.zeros_like(v_all)
    lib().orc_calculate_p(_p(v_all, C.c_float), _p(vbg_all, C.c_float), k_bg, A, K, W, _p(p, C.c_float))
    return p


def linear_s(v_all, vbg_all, A, K, K_bg, W):
    """reference: src/init/Motif.cpp:485-494"""
    s = np.zeros(A ** (K + 1) * W, np.float32)
    lib().orc_linear_s(_p(v_all, C.c_float), _p(vbg_all, C.c_float), A, K, K_bg, W, _p(s, C.c_float))
    return s


def log_s(v_all, vbg_all, A, K, K_bg, W):
    """reference: src/init/Motif.cpp:471-483"""
    s = np.zeros(A ** (K + 1) * W, np.float32)
    lib().orc_log_s(_p(v_all, C.c_float), _p(vbg_all, C.c_float), A, K, K_bg, W, _p(s, C.c_float))
    return s


def estep(kmer, offsets, A, K, W, s, q, want_double=False):
    """reference: src/refinement/EM.cpp:139-200. Returns (r, llh[, llh_double])."""
    r = np.zeros(len(kmer), np.float32)
    d = C.c_double(0.0)
    llh = lib().orc_estep(_p(kmer, C.c_uint64), _p(offsets, C.c_uint64), C.c_uint64(len(offsets) - 1), A, K, W,
                          _p(s, C.c_float), C.c_float(q), _p(r, C.c_float), C.byref(d))
    return (r, float(llh), d.value) if want_double else (r, float(llh))


def mstep(kmer, offsets, A, K, W, r, accumulate_double=False):
    """reference: src/refinement/EM.cpp:217-259 (counts of all orders, before updateV)"""
    n_all = np.zeros(model_size(A, K, W), np.float32)
    lib().orc_mstep(_p(kmer, C.c_uint64), _p(offsets, C.c_uint64), C.c_uint64(len(offsets) - 1), A, K, W,
                    _p(r, C.c_float), _p(n_all, C.c_float), int(accumulate_double))
    return n_all


def optimize_q(offsets, W, r):
    """reference: src/refinement/EM.cpp:505-519"""
    return float(lib().orc_optimize_q(_p(offsets, C.c_uint64), C.c_uint64(len(offsets) - 1), W, _p(r, C.c_float)))


def em_optimize(kmer, offsets, A, K, W, K_bg_model, vbg_all, alpha, v_all, q, optimize_q_flag=False,
                epsilon=0.01, max_iter=1000):
    """reference: src/refinement/EM.cpp:62-137. Returns dict(v, q, iterations, llh, vdiff, qtrace, n, r)."""
    v = np.array(v_all, np.float32, copy=True)
    r = np.zeros(len(kmer), np.float32)
    llh = np.zeros(max_iter, np.float32)
    vd = np.zeros(max_iter, np.float32)
    qt = np.zeros(max_iter, np.float32)
    n_all = np.zeros(model_size(A, K, W), np.float32)
    qio = C.c_float(q)
    alpha = np.ascontiguousarray(alpha, np.float32)
    it = lib().orc_em_optimize(_p(kmer, C.c_uint64), _p(offsets, C.c_uint64), C.c_uint64(len(offsets) - 1), A, K, W,
                               K_bg_model, _p(vbg_all, C.c_float), _p(alpha, C.c_float), _p(v, C.c_float), C.byref(qio),
                               int(optimize_q_flag), C.c_float(epsilon), max_iter, _p(r, C.c_float),
                               _p(llh, C.c_float), _p(vd, C.c_float), _p(qt, C.c_float), _p(n_all, C.c_float))
    return dict(v=v, q=float(qio.value), iterations=int(it), llh=llh[:it], vdiff=vd[:it], qtrace=qt[:it], n=n_all, r=r)


def em_mask(kmer, offsets, A, K, W, K_bg_model, vbg_all, alpha, v_all, q, f=0.2, epsilon=0.01, max_iter=1000):
    """EM::mask (EM.cpp:261-503, --advanceEM) without optimizeQ. Returns dict(iterations, v, r, llh, cutoff, nkept, n)."""
    L = lib()
    kmer = np.ascontiguousarray(kmer, np.uint64); offsets = np.ascontiguousarray(offsets, np.uint64)
    v = np.ascontiguousarray(v_all, np.float32).copy()
    vbg = np.ascontiguousarray(vbg_all, np.float32); al = np.ascontiguousarray(alpha, np.float32)
    r = np.zeros(int(offsets[-1]), np.float32)
    n_all = np.zeros(model_size(A, K, W), np.float32)
    llh = C.c_float(0); cut = C.c_float(0); nk = C.c_uint64(0)
    L.orc_em_mask.restype = C.c_int
    it = L.orc_em_mask(_p(kmer, C.c_uint64), _p(offsets, C.c_uint64), C.c_uint64(len(offsets) - 1), A, K, W, K_bg_model,
                       _p(vbg, C.c_float), _p(al, C.c_float), _p(v, C.c_float), C.c_float(q), C.c_float(f),
                       C.c_float(epsilon), max_iter, _p(r, C.c_float), C.byref(llh), C.byref(cut), C.byref(nk),
                       _p(n_all, C.c_float))
    return dict(iterations=it, v=v, r=r, llh=llh.value, cutoff=cut.value, nkept=nk.value, n=n_all)


def logodds(kmer, offsets, A, K, W, s, want_mops=True):
    """reference: src/seq_scoring/ScoreSeqSet.cpp:25-67. Returns (mops|None, zoops, z)."""
    nseq = len(offsets) - 1
    L = np.diff(offsets.astype(np.int64))
    mops = np.zeros(int((L - W + 1).sum()), np.float32) if want_mops else None
    zoops = np.zeros(nseq, np.float32)
    z = np.zeros(nseq, np.uint64)
    lib().orc_logodds(_p(kmer, C.c_uint64), _p(offsets, C.c_uint64), C.c_uint64(nseq), A, K, W, _p(s, C.c_float),
                      _p(mops, C.c_float), _p(zoops, C.c_float), _p(z, C.c_uint64))
    return mops, zoops, z


def em_iteration_omp(kmer, offsets, A, K, W, s, q, r, n_all, threads):
    """E-step + M-step with the reference's OpenMP structure (EM.cpp:148-149, 230-243); bench cpu_baseline 'port'."""
    return float(lib().orc_em_iteration_omp(_p(kmer, C.c_uint64), _p(offsets, C.c_uint64), C.c_uint64(len(offsets) - 1),
                                            A, K, W, _p(s, C.c_float), C.c_float(q), _p(r, C.c_float),
                                            _p(n_all, C.c_float), int(threads)))
