// ref_time — TEST / BENCH INFRASTRUCTURE (oracle side), never linked into the product.
//
// Times the reference's OWN CPU implementation of the hot path: it links the unmodified reference sources
// (compiled where they lie under /root/reference by oracle/Makefile) and calls the public
// EM::EStep() / EM::MStep() (reference: src/refinement/EM.h:30-31, EM.cpp:139-259) a fixed number of times
// on the sequences and initial model given on the reference's own command line (OUTDIR FASTA --bindingSiteFile F
// -k K -K Kbg --threads T ...). bench.py uses it for `cpu_baseline` (kind "reference") and for `--impl reference`.
//   BAMM_TIME_ITERS   timed iterations (default 3)      BAMM_TIME_WARMUP  untimed iterations (default 1)
// Prints one JSON line on stdout (everything else the reference prints goes to stderr).
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
#include <unistd.h>
#define private public
#include "refinement/Global.h"
#include "refinement/EM.h"
#include "seq_scoring/ScoreSeqSet.h"
#include "seq_generator/SeqGenerator.h"
#undef private

int main(int nargs, char* args[]) {
    int saved = dup(1);
    if (!freopen("/dev/null", "w", stdout)) return 2;      // silence the reference's banner chatter
    srand(42);
    Global::rngx.seed(42);
    Global::init(nargs, args);
    int iters = getenv("BAMM_TIME_ITERS") ? atoi(getenv("BAMM_TIME_ITERS")) : 3;
    int warm = getenv("BAMM_TIME_WARMUP") ? atoi(getenv("BAMM_TIME_WARMUP")) : 1;

    std::vector<Sequence*> posSet = Global::posSequenceSet->getSequences();
    BackgroundModel* bgModel = new BackgroundModel(posSet, Global::bgModelOrder, Global::bgModelAlpha,
                                                   Global::interpolateBG, Global::outputFileBasename);
    MotifSet motif_set(Global::initialModelFilename, Global::addColumns.at(0), Global::addColumns.at(1),
                       Global::initialModelTag, Global::posSequenceSet, bgModel->getV(), Global::bgModelOrder,
                       Global::modelOrder, Global::modelAlpha, Global::maxPWM, Global::q);
    Motif* motif = new Motif(*motif_set.getMotifs()[0]);
    size_t positions = 0, bp = 0;
    for (size_t n = 0; n < posSet.size(); n++) {
        size_t L = posSet[n]->getL();
        positions += L;
        bp += Global::ss ? L : (L - 1) / 2;
    }
    // BAMM_TIME_MODE=fdr: the FDR data path of config 4 instead of EM — negative-set sampling (SeqGenerator.cpp:188-206) and
    // ScoreSeqSet::calcLogOdds (ScoreSeqSet.cpp:25-67) on positives + sampled negatives, timed separately
    if (getenv("BAMM_TIME_MODE") && !strcmp(getenv("BAMM_TIME_MODE"), "fdr")) {
        auto t0 = std::chrono::high_resolution_clock::now();
        SeqGenerator negseq(posSet, NULL, Global::sOrder, 1.0f, Global::genericNeg);
        std::vector<std::unique_ptr<Sequence>> negSeqs = negseq.sample_bgseqset_by_fold(Global::mFold);
        auto t1 = std::chrono::high_resolution_clock::now();
        std::vector<Sequence*> all(posSet);
        size_t npos_scored = 0;
        for (size_t n = 0; n < negSeqs.size(); n++) all.push_back(negSeqs[n].get());
        for (size_t n = 0; n < all.size(); n++) npos_scored += all[n]->getL();
        double ts = 0;
        std::vector<double> per;
        for (int i = 0; i < warm + iters; i++) {
            ScoreSeqSet sc(motif, bgModel, all);
            auto a = std::chrono::high_resolution_clock::now();
            sc.calcLogOdds();
            auto b = std::chrono::high_resolution_clock::now();
            if (i >= warm) { per.push_back(std::chrono::duration<double>(b - a).count()); ts += per.back(); }
        }
        fflush(stdout);
        dup2(saved, 1);
        FILE* out = fdopen(saved, "w");
        fprintf(out, "{\"mode\": \"fdr\", \"iters\": %d, \"threads\": %zu, \"npos_seq\": %zu, \"nneg_seq\": %zu, \"mfold\": %zu, "
                     "\"positions_scored\": %zu, \"W\": %zu, \"K\": %zu, \"sample_neg_s\": %.6f, \"score_s\": %.6f, \"per_iter_s\": [",
                iters, Global::threads, posSet.size(), negSeqs.size(), Global::mFold, npos_scored, motif->getW(), motif->getK(),
                std::chrono::duration<double>(t1 - t0).count(), ts);
        for (size_t i = 0; i < per.size(); i++) fprintf(out, "%s%.6f", i ? ", " : "", per[i]);
        fprintf(out, "]}\n");
        fflush(out);
        return 0;
    }
    EM model(motif, bgModel, posSet, false, false, Global::f);
    for (int i = 0; i < warm; i++) { model.EStep(); model.MStep(); }
    double te = 0, tm = 0;
    std::vector<double> per_iter;
    for (int i = 0; i < iters; i++) {
        auto t0 = std::chrono::high_resolution_clock::now();
        model.EStep();
        auto t1 = std::chrono::high_resolution_clock::now();
        model.MStep();
        auto t2 = std::chrono::high_resolution_clock::now();
        double e = std::chrono::duration<double>(t1 - t0).count(), m = std::chrono::duration<double>(t2 - t1).count();
        te += e; tm += m; per_iter.push_back(e + m);
    }
    fflush(stdout);
    dup2(saved, 1);
    FILE* out = fdopen(saved, "w");
    fprintf(out, "{\"iters\": %d, \"warmup\": %d, \"threads\": %zu, \"nseq\": %zu, \"positions\": %zu, \"bp\": %zu, "
                 "\"W\": %zu, \"K\": %zu, \"estep_s\": %.6f, \"mstep_s\": %.6f, \"llh\": %.6f, \"per_iter_s\": [",
            iters, warm, Global::threads, posSet.size(), positions, bp, motif->getW(), motif->getK(), te, tm,
            (double)model.llikelihood_);
    for (size_t i = 0; i < per_iter.size(); i++) fprintf(out, "%s%.6f", i ? ", " : "", per_iter[i]);
    fprintf(out, "]}\n");
    fflush(out);
    return 0;
}
