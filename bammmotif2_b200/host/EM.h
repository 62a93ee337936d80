// EM — expectation-maximisation refinement of one motif; host wrapper over the device path.
// Public interface = the reference's (src/refinement/EM.h:18-36); the work behind EStep / MStep / optimize /
// optimize_q runs in the sm_100a kernels behind include/bamm_b200.h. The sequence data are NOT copied per EM object:
// the wrapper resolves its std::vector<Sequence*> to an index subset of the SequenceSet already resident in HBM
// (FDR training folds, filtered sets), so concurrent EM objects over one set share it read-only
// (the reference runs its folds concurrently too, src/evaluation/FDR.cpp:37-38).
#ifndef BAMM_HOST_EM_H_
#define BAMM_HOST_EM_H_

#include <string>
#include <vector>

#include "BackgroundModel.h"
#include "MotifSet.h"

class EM {
public:
    EM( Motif* motif, BackgroundModel* bgModel, std::vector<Sequence*> seqs,
        bool optimizeQ = true, bool verbose = false, float f = 0.2f );
    ~EM();
    EM( const EM& ) = delete;
    EM& operator=( const EM& ) = delete;

    int         optimize();             // EM loop with the reference's stop rule (EM.cpp:62-137)
    int         mask();                 // advanced EM (EM.cpp:261-503) behind bamm_em_mask; not with optimizeQ
    void        print();
    void        write( char* odir, std::string basename, bool ss );

    void        EStep();                // EM.cpp:139-200
    void        MStep();                // EM.cpp:217-259 (+ Motif::updateV)
    void        optimize_q();           // EM.cpp:505-519

    float**     getR();                 // r[n][i], i = L-W-p (reversed, zero tail) — downloaded on demand
    float       getQ()                  { return q_; }
    void        printR();

    float       getLogLikelihood()      { return llikelihood_; }
    size_t      getIterations()         { return iterations_; }

private:
    void        uploadModel();
    void        fetchR();
    void        fetchCounts();

    Motif*                  motif_;
    BackgroundModel*        bgModel_;
    std::vector<Sequence*>  seqs_;
    bamm_em*                dev_ = nullptr;

    size_t                  K_, W_, K_bg_;
    float                   q_;
    float                   f_;
    bool                    optimizeQ_;
    bool                    verbose_;
    float                   llikelihood_ = 0.0f;
    float                   epsilon_ = 0.01f;           // EM.h:62
    size_t                  maxEMIterations_ = 1000;    // EM.h:63
    size_t                  iterations_ = 0;
    std::vector<size_t>     Y_;

    std::vector<float>      r_;                         // flat, reference index order per sequence
    std::vector<float*>     rRows_;
    bool                    rFresh_ = false;
    std::vector<float>      n_;                         // all orders [k][y][j]
    bool                    nFresh_ = false;
};

#endif
