// ref_dump — TEST INFRASTRUCTURE (oracle side), never linked into the product.
//
// Links the UNMODIFIED reference sources (compiled where they lie under /root/reference by
// oracle/Makefile) and drives the reference's own public API step by step so that every
// intermediate of the hot path can be written out as raw arrays:
//   Global::init            (reference: src/refinement/Global.cpp:98-118)  -> codes, kmer_
//   BackgroundModel(...)    (reference: src/init/BackgroundModel.cpp:3-46) -> vbg[k][y]
//   MotifSet(...)           (reference: src/init/MotifSet.cpp:3-223)       -> initial v[k][y][j]
//   EM::EStep / EM::MStep   (reference: src/refinement/EM.cpp:139-259)     -> r, n, v, llh per iteration
//   EM::optimize            (reference: src/refinement/EM.cpp:62-137)      -> cross-check of the stepwise loop
//   ScoreSeqSet::calcLogOdds(reference: src/seq_scoring/ScoreSeqSet.cpp:25-67)
//   FDR::evaluateMotif      (reference: src/evaluation/FDR.cpp:28-145)     -> pos/neg ZOOPS scores, PR curves
//
// The only liberty taken is `#define private public` in THIS translation unit so private members
// (llikelihood_, n_, z_, ...) can be read; the reference objects themselves are built from
// untouched sources. Same command line as the reference's BaMMmotif (OUTDIR FASTA [options]);
// extra behaviour is controlled by environment variables so the option parser stays untouched:
//   BAMM_DUMP_R_ITERS="1,2,41"  iterations whose full r[n][i] is dumped (default "1")
//   BAMM_DUMP_MAXITER=N         stop the step-wise loop after N iterations (default: convergence)
// Output: <OUTDIR>/dump/*.npy (+ meta.txt).
// every standard header first: libstdc++ does not survive the access-specifier trick below
#include <algorithm>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <limits>
#include <list>
#include <map>
#include <memory>
#include <numeric>
#include <random>
#include <set>
#include <sstream>
#include <string>
#include <utility>
#include <vector>
#include <sys/stat.h>
#include <stdint.h>
#include <ctype.h>
#include <float.h>
#define private public
#define protected public
#include "refinement/Global.h"
#include "refinement/EM.h"
#include "evaluation/FDR.h"
#undef private
#undef protected

#include <cstdio>
#include <cstdlib>
#include <sstream>
#include <set>

static std::string g_dir;

static void npy_write(const std::string& name, const char* descr, size_t itemsize, const void* data, size_t n) {
    std::string path = g_dir + "/" + name + ".npy";
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { fprintf(stderr, "ref_dump: cannot write %s\n", path.c_str()); exit(2); }
    std::ostringstream hdr;
    hdr << "{'descr': '" << descr << "', 'fortran_order': False, 'shape': (" << n << ",), }";
    std::string h = hdr.str();
    size_t total = 10 + h.size() + 1;
    size_t pad = (64 - total % 64) % 64;
    h.append(pad, ' ');
    h.push_back('\n');
    unsigned char magic[10] = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0, 0, 0};
    magic[8] = (unsigned char)(h.size() & 0xff);
    magic[9] = (unsigned char)(h.size() >> 8);
    fwrite(magic, 1, 10, f);
    fwrite(h.data(), 1, h.size(), f);
    if (n) fwrite(data, itemsize, n, f);
    fclose(f);
}
static void dump_f32(const std::string& name, const std::vector<float>& v) { npy_write(name, "<f4", 4, v.data(), v.size()); }
static void dump_u64(const std::string& name, const std::vector<uint64_t>& v) { npy_write(name, "<u8", 8, v.data(), v.size()); }
static void dump_u8(const std::string& name, const std::vector<uint8_t>& v) { npy_write(name, "|u1", 1, v.data(), v.size()); }

static std::vector<float> flat_v(float*** v, size_t K, size_t W, const std::vector<size_t>& Y) {
    std::vector<float> out;
    for (size_t k = 0; k <= K; k++)
        for (size_t y = 0; y < Y[k + 1]; y++)
            for (size_t j = 0; j < W; j++) out.push_back(v[k][y][j]);
    return out;
}

static void dump_seqs(const std::string& prefix, const std::vector<Sequence*>& seqs) {
    std::vector<uint8_t> codes;
    std::vector<uint64_t> kmer, off;
    off.push_back(0);
    for (size_t n = 0; n < seqs.size(); n++) {
        size_t L = seqs[n]->getL();
        uint8_t* s = seqs[n]->getSequence();
        size_t* km = seqs[n]->getKmer();
        for (size_t i = 0; i < L; i++) { codes.push_back(s[i]); kmer.push_back((uint64_t)km[i]); }
        off.push_back(codes.size());
    }
    dump_u8(prefix + "_codes", codes);
    dump_u64(prefix + "_kmer", kmer);
    dump_u64(prefix + "_offsets", off);
}

int main(int nargs, char* args[]) {
    srand(42);                  // reference: src/refinement/mainBaMM.cpp:22-23
    Global::rngx.seed(42);
    Global::init(nargs, args);
    g_dir = std::string(Global::outputDirectory) + "/dump";
    { std::string cmd = "mkdir -p " + g_dir; if (system(cmd.c_str()) != 0) return 2; }

    std::set<size_t> r_iters;
    {
        const char* e = getenv("BAMM_DUMP_R_ITERS");
        std::string s = e ? e : "1";
        std::stringstream ss(s); std::string tok;
        while (std::getline(ss, tok, ',')) if (!tok.empty()) r_iters.insert((size_t)atol(tok.c_str()));
    }
    size_t max_iter_env = 0;
    if (const char* e = getenv("BAMM_DUMP_MAXITER")) max_iter_env = (size_t)atol(e);

    std::vector<Sequence*> posSet = Global::posSequenceSet->getSequences();
    BackgroundModel* bgModel = new BackgroundModel(posSet, Global::bgModelOrder, Global::bgModelAlpha,
                                                   Global::interpolateBG, Global::outputFileBasename);
    bgModel->write(Global::outputDirectory, Global::outputFileBasename);

    MotifSet motif_set(Global::initialModelFilename, Global::addColumns.at(0), Global::addColumns.at(1),
                       Global::initialModelTag, Global::posSequenceSet, bgModel->getV(), Global::bgModelOrder,
                       Global::modelOrder, Global::modelAlpha, Global::maxPWM, Global::q);

    for (auto it = posSet.begin(); it != posSet.end();) {   // reference: mainBaMM.cpp:75-83
        if ((*it)->getL() < motif_set.getMaxW()) posSet.erase(it); else ++it;
    }
    dump_seqs("pos", posSet);

    // background model tables
    {
        std::vector<size_t> Yb; for (size_t k = 0; k < Global::bgModelOrder + 8; k++) Yb.push_back(ipow(Alphabet::getSize(), k));
        std::vector<float> vb;
        for (size_t k = 0; k <= Global::bgModelOrder; k++)
            for (size_t y = 0; y < Yb[k + 1]; y++) vb.push_back(bgModel->getV()[k][y]);
        dump_f32("bg_v", vb);
        std::vector<uint64_t> nb;
        for (size_t k = 0; k <= Global::bgModelOrder; k++)
            for (size_t y = 0; y < Yb[k + 1]; y++) nb.push_back((uint64_t)bgModel->n_[k][y]);
        dump_u64("bg_n", nb);
    }

    // negative set exactly as the reference samples it (mainBaMM.cpp:97-116)
    std::vector<Sequence*> negSet;
    {
        size_t minSeqN = 5000;
        bool rest = minSeqN % posSet.size();
        if (posSet.size() < minSeqN) Global::mFold = minSeqN / posSet.size() + rest;
        SeqGenerator negseq(posSet, NULL, Global::sOrder, 1.0f, Global::genericNeg);
        std::vector<std::unique_ptr<Sequence>> negSeqs = negseq.sample_bgseqset_by_fold(Global::mFold);
        for (size_t n = 0; n < negSeqs.size(); n++) negSet.push_back(negSeqs[n].release());
    }
    if (getenv("BAMM_DUMP_NEG")) dump_seqs("neg", negSet);

    FILE* meta = fopen((g_dir + "/meta.txt").c_str(), "w");
    fprintf(meta, "A %zu\nK %zu\nK_bg_model %zu\nnmotifs %zu\nq %.9g\nss %d\nmFold %zu\ncvFold %zu\nnpos %zu\nnneg %zu\n",
            Alphabet::getSize(), Global::modelOrder, Global::bgModelOrder, motif_set.getN(), Global::q,
            (int)Global::ss, Global::mFold, Global::cvFold, posSet.size(), negSet.size());

    for (size_t m = 0; m < motif_set.getN(); m++) {
        std::string mp = "m" + std::to_string(m + 1) + "_";
        Motif* motif = new Motif(*motif_set.getMotifs()[m]);
        size_t K = motif->getK(), W = motif->getW();
        std::vector<size_t> Y = motif->getY();
        fprintf(meta, "motif %zu W %zu\n", m + 1, W);
        dump_f32(mp + "v_init", flat_v(motif->getV(), K, W, Y));
        { std::vector<float> a; for (size_t k = 0; k <= K; k++) for (size_t j = 0; j < W; j++) a.push_back(motif->getA()[k][j]); dump_f32(mp + "alpha", a); }

        if (Global::EM) {
            // (1) the reference's own loop on a private copy: iteration count + final model
            Motif* motif_ref = new Motif(*motif_set.getMotifs()[m]);
            size_t ref_iters = 0;
            {
                EM model(motif_ref, bgModel, posSet, Global::optimizeQ, false, Global::f);
                model.optimize();
                dump_f32(mp + "opt_v_final", flat_v(motif_ref->getV(), K, W, Y));
                dump_f32(mp + "opt_p_final", flat_v(motif_ref->p_, K, W, Y));
                std::vector<float> qf(1, model.getQ()); dump_f32(mp + "opt_q_final", qf);
            }
            // (2) the same loop step by step through the public EStep/MStep (EM.h:30-32)
            EM model(motif, bgModel, posSet, Global::optimizeQ, false, Global::f);
            std::vector<float> llh_tr, vdiff_tr, q_tr;
            bool iterate = true; size_t iteration = 0; float llh_prev;
            size_t YK = Y[K + 1];
            std::vector<float> v_before(YK * W);
            while (iterate && iteration < model.maxEMIterations_) {
                iteration++;
                llh_prev = model.llikelihood_;
                for (size_t y = 0; y < YK; y++) for (size_t j = 0; j < W; j++) v_before[y * W + j] = motif->getV()[K][y][j];
                model.EStep();
                if (r_iters.count(iteration)) {
                    std::vector<float> r;
                    for (size_t n = 0; n < posSet.size(); n++) for (size_t i = 0; i < posSet[n]->getL(); i++) r.push_back(model.r_[n][i]);
                    dump_f32(mp + "r_it" + std::to_string(iteration), r);
                    std::vector<float> s; for (size_t y = 0; y < YK; y++) for (size_t j = 0; j < W; j++) s.push_back(model.s_[y][j]);
                    dump_f32(mp + "s_it" + std::to_string(iteration), s);
                }
                model.MStep();
                if (Global::optimizeQ && iteration <= 5) model.optimize_q();
                float v_diff = 0.0f;
                for (size_t y = 0; y < YK; y++) for (size_t j = 0; j < W; j++) v_diff += fabsf(motif->getV()[K][y][j] - v_before[y * W + j]);
                float llh_diff = model.llikelihood_ - llh_prev;
                llh_tr.push_back(model.llikelihood_); vdiff_tr.push_back(v_diff); q_tr.push_back(model.q_);
                dump_f32(mp + "n_it" + std::to_string(iteration), flat_v(model.n_, K, W, Y));
                dump_f32(mp + "v_it" + std::to_string(iteration), flat_v(motif->getV(), K, W, Y));
                if (v_diff < model.epsilon_) iterate = false;
                if (llh_diff < 0 && iteration > 10) iterate = false;
                if (max_iter_env && iteration >= max_iter_env) iterate = false;
            }
            motif->calculateP();
            ref_iters = iteration;
            dump_f32(mp + "llh", llh_tr); dump_f32(mp + "vdiff", vdiff_tr); dump_f32(mp + "q", q_tr);
            dump_f32(mp + "v_final", flat_v(motif->getV(), K, W, Y));
            dump_f32(mp + "p_final", flat_v(motif->p_, K, W, Y));
            fprintf(meta, "motif %zu iterations %zu\n", m + 1, ref_iters);
        }
        // EM::mask (--advanceEM, EM.cpp:261-503) from the same initial motif: final model, r, log likelihood
        if (getenv("BAMM_DUMP_MASK")) {
            Motif* motif_m = new Motif(*motif_set.getMotifs()[m]);
            EM model(motif_m, bgModel, posSet, false, false, Global::f);
            model.mask();
            dump_f32(mp + "mask_v_final", flat_v(motif_m->getV(), K, W, Y));
            std::vector<float> r;
            for (size_t n = 0; n < posSet.size(); n++) for (size_t i = 0; i < posSet[n]->getL(); i++) r.push_back(model.r_[n][i]);
            dump_f32(mp + "mask_r", r);
            std::vector<float> l(1, model.llikelihood_); dump_f32(mp + "mask_llh", l);
            dump_f32(mp + "mask_n", flat_v(model.n_, K, W, Y));
        }
        motif->write(Global::outputDirectory, Global::outputFileBasename + "_motif_" + std::to_string(m + 1));

        // scoring of the positive set with the (learned) motif
        {
            ScoreSeqSet sc(motif, bgModel, posSet);
            sc.calcLogOdds();
            std::vector<float> mops, zo = sc.getZoopsScores();
            std::vector<uint64_t> z;
            std::vector<std::vector<float>> ms = sc.getMopsScores();
            for (size_t n = 0; n < ms.size(); n++) { mops.insert(mops.end(), ms[n].begin(), ms[n].end()); z.push_back((uint64_t)sc.z_[n]); }
            dump_f32(mp + "score_mops", mops); dump_f32(mp + "score_zoops", zo); dump_u64(mp + "score_z", z);
            std::vector<float> s; for (size_t y = 0; y < Y[K + 1]; y++) for (size_t j = 0; j < W; j++) s.push_back(motif->getS()[y][j]);
            dump_f32(mp + "score_logs", s);
            // p-values of every positive window against all negative window scores, as mainBaMM.cpp:203-230 computes them
            // for --scoreSeqset (ScoreSeqSet::calcPvalues, src/seq_scoring/ScoreSeqSet.cpp:70-126)
            if (getenv("BAMM_DUMP_PVALUES")) {
                ScoreSeqSet sn(motif, bgModel, negSet);
                sn.calcLogOdds();
                std::vector<std::vector<float>> na = sn.getMopsScores();
                std::vector<float> negScores;
                for (size_t n = 0; n < negSet.size(); n++) negScores.insert(negScores.end(), na[n].begin(), na[n].end());
                sc.calcPvalues(ms, negScores);
                std::vector<float> pv, evv;
                for (size_t n = 0; n < posSet.size(); n++) {
                    pv.insert(pv.end(), sc.mops_p_values_[n].begin(), sc.mops_p_values_[n].end());
                    evv.insert(evv.end(), sc.mops_e_values_[n].begin(), sc.mops_e_values_[n].end());
                }
                dump_f32(mp + "pval_neg_all", negScores); dump_f32(mp + "pval_p", pv); dump_f32(mp + "pval_e", evv);
            }
        }

        if (Global::FDR) {
            Motif* mf = new Motif(*motif_set.getMotifs()[m]);
            FDR fdr(posSet, negSet, mf, bgModel, Global::cvFold, Global::mops, Global::zoops, Global::savePRs,
                    Global::savePvalues, Global::saveLogOdds);
            fdr.evaluateMotif(Global::EM, Global::CGS, Global::optimizeQ, Global::advanceEM, Global::f, 1);
            dump_f32(mp + "fdr_posScoreMax", fdr.posScoreMax_); dump_f32(mp + "fdr_negScoreMax", fdr.negScoreMax_);
            dump_f32(mp + "fdr_TP", fdr.ZOOPS_TP_); dump_f32(mp + "fdr_FP", fdr.ZOOPS_FP_);
            dump_f32(mp + "fdr_FDR", fdr.ZOOPS_FDR_); dump_f32(mp + "fdr_Rec", fdr.ZOOPS_Rec_);
            dump_f32(mp + "fdr_PNpval", fdr.PN_Pvalue_);
            if (Global::savePvalues) {        // FDR::calculatePvalues (src/evaluation/FDR.cpp:278-330): ranks of the ZOOPS / MOPS scores
                dump_f32(mp + "fdr_zoops_pvalue", fdr.ZOOPS_Pvalue_); dump_f32(mp + "fdr_mops_pvalue", fdr.MOPS_Pvalue_);
                dump_f32(mp + "fdr_posScoreAll", fdr.posScoreAll_); dump_f32(mp + "fdr_negScoreAll", fdr.negScoreAll_);
            }
            std::vector<float> occ(1, fdr.occ_frac_); dump_f32(mp + "fdr_occ_frac", occ);
            fdr.write(Global::outputDirectory, Global::outputFileBasename + "_motif_" + std::to_string(m + 1));
        }
    }
    fclose(meta);
    return 0;
}
