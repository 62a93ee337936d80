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
