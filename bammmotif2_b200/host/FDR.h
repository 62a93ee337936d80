// FDR — cross-validated motif evaluation: per fold train on (cv-1)/cv of the positives with EM, score the held-out
// positives and the negative subset, then precision / recall / p-value statistics.
// Public interface = the reference's (src/evaluation/FDR.h:27-42). Folds are index subsets of the sequence sets that are
// already resident in HBM — nothing is re-uploaded per fold. The Gibbs-sampling training option is out of scope.
#ifndef BAMM_HOST_FDR_H_
#define BAMM_HOST_FDR_H_

#include <string>
#include <vector>

#include "BackgroundModel.h"
#include "EM.h"
#include "Motif.h"
#include "ScoreSeqSet.h"

class FDR {
public:
    FDR( std::vector<Sequence*> posSeqs, std::vector<Sequence*> negSeqs, Motif* motif = NULL, BackgroundModel* bgmodel = NULL,
         size_t cvFold = 4, bool mops = false, bool zoops = true, bool savePRs = true, bool savePvalues = false,
         bool saveLogOdds = false );
    ~FDR();

    void    evaluateMotif( bool EMoptimize = false, bool CGSoptimize = false, bool optimizeQ = false, bool advanceEM = false,
                           float f = 0.05f, size_t perLoopThreads = 4 );
    void    print();
    void    write( char* odir, std::string basename );
    void    saveUnsortedLogOdds( std::string opath, std::vector<float> logOdds );

    // score vectors / statistics (tests)
    const std::vector<float>& posScoreMax() const   { return posScoreMax_; }
    const std::vector<float>& negScoreMax() const   { return negScoreMax_; }
    const std::vector<float>& zoopsTP() const       { return ZOOPS_TP_; }
    const std::vector<float>& zoopsFP() const       { return ZOOPS_FP_; }
    const std::vector<float>& zoopsFDR() const      { return ZOOPS_FDR_; }
    const std::vector<float>& zoopsRecall() const   { return ZOOPS_Rec_; }
    const std::vector<float>& pnPvalues() const     { return PN_Pvalue_; }
    const std::vector<float>& zoopsPvalues() const  { return ZOOPS_Pvalue_; }
    float                     occFrac() const       { return occ_frac_; }
    // feeds score vectors directly (tests of the statistics without a device)
    void    setScores( std::vector<float> posMax, std::vector<float> negMax ){ posScoreMax_ = posMax; negScoreMax_ = negMax; }
    void    calculatePR();
    void    calculatePvalues();

private:
    std::vector<Sequence*>  posSeqs_;
    std::vector<Sequence*>  negSeqs_;
    float                   q_;
    Motif*                  motif_;
    BackgroundModel*        bgModel_;
    size_t                  cvFold_;
    bool                    mops_, zoops_, savePRs_, savePvalues_, saveLogOdds_;

    std::vector<float>      posScoreAll_, posScoreMax_, negScoreAll_, negScoreMax_;
    std::vector<float>      ZOOPS_FDR_, ZOOPS_Rec_, ZOOPS_TP_, ZOOPS_FP_;
    std::vector<float>      MOPS_FDR_, MOPS_Rec_, MOPS_TP_, MOPS_FP_;
    float                   occ_frac_ = 0.0f;
    float                   occ_mult_ = 0.0f;
    std::vector<float>      PN_Pvalue_, ZOOPS_Pvalue_, MOPS_Pvalue_;
};

#endif
