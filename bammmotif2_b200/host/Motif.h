// Motif — inhomogeneous Bayesian Markov model tables of one motif.
// Interface mirrors the reference (src/init/Motif.h:9-45). Storage is flat and in the C-ABI order
// (for k, for y, for j — the reference's v_[k][y][j]), so a model goes to / comes from the device with one copy;
// getV()/getS()/getA() hand out reference-style pointer views (v[k][y][j], s[y][j], A[k][j]) into the same memory.
#ifndef BAMM_HOST_MOTIF_H_
#define BAMM_HOST_MOTIF_H_

#include <cassert>
#include <string>
#include <vector>

#include "BackgroundModel.h"

class Motif {
public:
    Motif( size_t length, size_t order, std::vector<float> alpha, float** v_bg, size_t k_bg, float glob_q );
    Motif( const Motif& other );
    ~Motif();

    void initFromBindingSites( char* indir, size_t l_flank, size_t r_flank );
    void initFromPWM( float** PWM, size_t asize, SequenceSet* posSet, float q );
    void initFromBaMM( char* indir, size_t l_flank, size_t r_flank );

    size_t              getW()      { return W_; }
    size_t              getK()      { return K_; }
    float               getQ()      { return q_; }
    float**             getA()      { return aRows_.data(); }
    float***            getV()      { return vOrders_.data(); }
    float**             getS()      { return sRows_.data(); }
    std::vector<size_t> getY()      { return Y_; }

    void                updateV( float*** n, float** alpha, size_t k );
    void                calculateP();
    void                calculateLogS( float** Vbg, size_t K_bg );
    void                calculateLinearS( float** Vbg, size_t K_bg );

    void                print();
    void                write( char* odir, std::string basename );

    // ---- flat access for the device wrappers ----
    std::vector<float>&         flatV()             { return v_; }      // all orders, [k][y][j]
    const std::vector<float>&   flatP() const       { return p_; }
    const std::vector<float>&   flatAlpha() const   { return a_; }      // [k][j]
    std::vector<float>&         flatS()             { return s_; }      // [y][j] of the top order
    size_t                      offsetOfOrder( size_t k ) const { return off_[k]; }
    void                        setQ( float q )     { q_ = q; }
    void                        markInitialized()   { isInitialized_ = true; }
    size_t                      backgroundOrder() const { return k_bg_; }

private:
    void                allocate();
    void                bindViews();
    void                calculateV( const std::vector<int>& n );

    bool                isInitialized_ = false;
    size_t              C_ = 0;             // number of binding sites read
    size_t              W_;
    size_t              K_;
    float               q_;
    float**             v_bg_;              // shared with the BackgroundModel (never freed here), v_bg[k][y]
    size_t              k_bg_;
    std::vector<float>  ownBg_;             // uniform background when none was given
    std::vector<float*> ownBgRows_;
    std::vector<size_t> Y_;
    std::vector<size_t> off_;               // off_[k] = start of order k in v_/p_/n_
    std::vector<float>  v_, p_, a_, s_;
    std::vector<int>    n_;
    std::vector<float*>  vRows_, sRows_, aRows_;
    std::vector<float**> vOrders_;
};

#endif
