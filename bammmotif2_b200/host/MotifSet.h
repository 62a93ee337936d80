// MotifSet — initial models parsed from a binding-site block, a MEME v4 PWM file or a BaMM (.ihbcp) file.
// Interface mirrors the reference (src/init/MotifSet.h:9-22; parsing rules src/init/MotifSet.cpp:3-223).
#ifndef BAMM_HOST_MOTIFSET_H_
#define BAMM_HOST_MOTIFSET_H_

#include "Motif.h"

class MotifSet {
public:
    MotifSet( char* indir, size_t l_flank, size_t r_flank, std::string tag,
              SequenceSet* posSet = NULL, float** v_bg = NULL, size_t k_bg = 2,
              size_t order = 2, std::vector<float> alphas = {}, size_t maxPWM = 10, float glob_q = 0.9f );
    ~MotifSet();

    std::vector<Motif*> getMotifs()     { return motifs_; }
    size_t              getN()          { return N_; }
    size_t              getMaxW()       { return maxW_; }
    void                print();
    void                write( char* outdir );

private:
    std::vector<Motif*> motifs_;
    size_t              N_ = 0;
    size_t              maxW_ = 0;
};

#endif
