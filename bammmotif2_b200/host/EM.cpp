#include "EM.h"

#include <chrono>
#include <cmath>
#include <fstream>
#include <iomanip>
#include <iostream>

EM::EM( Motif* motif, BackgroundModel* bgModel, std::vector<Sequence*> seqs, bool optimizeQ, bool verbose, float f )
    : motif_( motif ), bgModel_( bgModel ), seqs_( seqs ), q_( motif->getQ() ), f_( f ), optimizeQ_( optimizeQ ),
      verbose_( verbose ){
    K_ = motif_->getK();
    W_ = motif_->getW();
    Y_ = motif_->getY();
    K_bg_ = ( bgModel_->getOrder() < K_ ) ? bgModel_->getOrder() : K_;

    std::vector<uint64_t> indices;
    bool whole = false;
    SequenceSet* set = SequenceSet::commonSet( seqs_, indices, &whole );
    BAMM_CHECK( bamm_em_create( set->device(), whole ? nullptr : indices.data(), indices.size(),
                                static_cast<int>( W_ ), static_cast<int>( K_ ), static_cast<int>( bgModel_->getOrder() ), &dev_ ) );
}

EM::~EM(){
    if( dev_ ) bamm_em_destroy( dev_ );
}

void EM::uploadModel(){
    BAMM_CHECK( bamm_em_set_model( dev_, motif_->flatV().data(), bgModel_->flatV().data(), motif_->flatAlpha().data(), q_ ) );
}

void EM::EStep(){
    // the Motif is the source of truth between calls (callers may have changed it), as in the reference where
    // EStep starts from motif_->calculateLinearS (EM.cpp:142)
    uploadModel();
    motif_->calculateLinearS( bgModel_->getV(), K_bg_ );
    BAMM_CHECK( bamm_em_estep( dev_, &llikelihood_ ) );
    rFresh_ = false;
}

void EM::MStep(){
    BAMM_CHECK( bamm_em_mstep( dev_ ) );
    BAMM_CHECK( bamm_em_get_model( dev_, motif_->flatV().data() ) );
    nFresh_ = false;
}

void EM::optimize_q(){
    BAMM_CHECK( bamm_em_optimize_q( dev_, &q_ ) );
    if( verbose_ ) std::cout << "optimized q=" << q_ << std::endl;
}

int EM::optimize(){
    auto t0_wall = std::chrono::high_resolution_clock::now();

    uploadModel();
    std::vector<float> llh( maxEMIterations_ ), vdiff( maxEMIterations_ ), qtrace( maxEMIterations_ );
    int iterations = 0;
    BAMM_CHECK( bamm_em_optimize( dev_, optimizeQ_ ? 1 : 0, epsilon_, static_cast<int>( maxEMIterations_ ), &iterations,
                                  llh.data(), vdiff.data(), qtrace.data() ) );
    iterations_ = static_cast<size_t>( iterations );
    if( verbose_ ){
        // same lines, same order as the reference prints them inside its loop (EM.cpp:99, 112-115)
        float prev = llikelihood_;  // the member starts at 0 (EM.h:61) and carries over between calls, as in the reference
        for( int it = 0; it < iterations; it++ ){
            if( optimizeQ_ && it < 5 ) std::cout << "optimized q=" << qtrace[it] << std::endl;
            std::cout << it + 1 << " iter, llh=" << llh[it] << ", diff_llh=" << llh[it] - prev << ", v_diff=" << vdiff[it] << std::endl;
            prev = llh[it];
        }
    }
    if( iterations > 0 ){
        llikelihood_ = llh[iterations - 1];
        q_ = qtrace[iterations - 1];
    }
    // bring the learned model home: v of every order, the odds table of the last E-step (what Motif::getS() held in
    // the reference at this point), then the probabilities p (EM.cpp:131)
    BAMM_CHECK( bamm_em_get_model( dev_, motif_->flatV().data() ) );
    BAMM_CHECK( bamm_em_get_s( dev_, motif_->flatS().data() ) );
    motif_->calculateP();
    rFresh_ = false;
    nFresh_ = false;

    auto t1_wall = std::chrono::high_resolution_clock::now();
    auto t_diff = std::chrono::duration_cast<std::chrono::duration<double>>( t1_wall - t0_wall );
    std::cout << "\n--- Runtime for EM: " << t_diff.count() << " seconds ---\n";
    return 0;
}

// reference: EM::mask, src/refinement/EM.cpp:261-503 (the "advanced EM" of --advanceEM): order-0 E-step over all windows,
// threshold that keeps the fraction f_ of them, EM over the kept windows. All three phases run behind bamm_em_mask.
int EM::mask(){
    auto t0_wall = std::chrono::high_resolution_clock::now();
    if( optimizeQ_ ){
        // the reference calls optimize_q() inside its per-sequence loop (EM.cpp:316), i.e. on responsibilities that are
        // partly stale and partly not yet computed; that order dependence is not reproduced
        std::cerr << "Error: --advanceEM cannot be combined with --optimizeQ on the B200 path." << std::endl;
        exit( 1 );
    }
    uploadModel();
    int iterations = 0;
    float llh = 0.0f, cutoff = 0.0f;
    uint64_t kept = 0;
    BAMM_CHECK( bamm_em_mask( dev_, f_, epsilon_, static_cast<int>( maxEMIterations_ ), &iterations, &llh, &kept, &cutoff ) );
    iterations_ = static_cast<size_t>( iterations );
    llikelihood_ = llh;
    if( verbose_ ) std::cout << iterations << " iterations on " << kept << " windows (r >= " << cutoff << ")" << std::endl;
    BAMM_CHECK( bamm_em_get_model( dev_, motif_->flatV().data() ) );
    BAMM_CHECK( bamm_em_get_s( dev_, motif_->flatS().data() ) );
    motif_->calculateP();
    rFresh_ = false;
    nFresh_ = false;
    auto t1_wall = std::chrono::high_resolution_clock::now();
    auto t_diff = std::chrono::duration_cast<std::chrono::duration<double>>( t1_wall - t0_wall );
    std::cout << "\n--- Runtime for EM: " << t_diff.count() << " seconds ---\n";
    return 0;
}

void EM::fetchR(){
    if( rFresh_ ) return;
    const size_t N = seqs_.size();
    r_.resize( bamm_em_r_size( dev_ ) );
    BAMM_CHECK( bamm_em_get_r( dev_, 0, N, r_.data() ) );
    rRows_.resize( N );
    size_t off = 0;
    for( size_t n = 0; n < N; n++ ){ rRows_[n] = r_.data() + off; off += seqs_[n]->getL(); }
    rFresh_ = true;
}

void EM::fetchCounts(){
    if( nFresh_ ) return;
    n_.resize( motif_->flatV().size() );
    BAMM_CHECK( bamm_em_get_counts( dev_, n_.data() ) );
    nFresh_ = true;
}

float** EM::getR(){
    fetchR();
    return rRows_.data();
}

void EM::print(){
    fetchCounts();
    const float* nK = n_.data() + motif_->offsetOfOrder( K_ );
    for( size_t j = 0; j < W_; j++ ){
        for( size_t y = 0; y < Y_[K_ + 1]; y++ ) std::cout << std::setprecision( 3 ) << nK[y * W_ + j] << '\t';
        std::cout << std::endl;
    }
}

void EM::printR(){
    fetchR();
    for( size_t n = 0; n < seqs_.size(); n++ ){
        std::cout << "seq " << n << ":" << std::endl;
        const size_t L = seqs_[n]->getL();
        for( size_t i = 0; i + W_ <= L; i++ ) std::cout << rRows_[n][L - W_ - i] << '\t';
        std::cout << std::endl;
    }
}

// .counts: fractional counts truncated to int, per position one line per order; .positions: every window start whose
// posterior reaches 0.3 (reference: EM::write, src/refinement/EM.cpp:553-615)
void EM::write( char* odir, std::string basename, bool ss ){
    const std::string opath = std::string( odir ) + '/' + basename;
    fetchCounts();
    std::ofstream ofile_n( ( opath + ".counts" ).c_str() );
    for( size_t j = 0; j < W_; j++ ){
        for( size_t k = 0; k <= K_; k++ ){
            const float* nk = n_.data() + motif_->offsetOfOrder( k );
            for( size_t y = 0; y < Y_[k + 1]; y++ ) ofile_n << static_cast<int>( nk[y * W_ + j] ) << '\t';
            ofile_n << '\n';
        }
        ofile_n << '\n';
    }

    fetchR();
    std::ofstream ofile_pos( ( opath + ".positions" ).c_str() );
    ofile_pos << "seq\tlength\tstrand\tstart..end\tpattern" << std::endl;
    const float cutoff = 0.3f;
    for( size_t n = 0; n < seqs_.size(); n++ ){
        const size_t Lstored = seqs_[n]->getL();
        const size_t L = ss ? Lstored : ( Lstored - 1 ) / 2;
        const uint8_t* codes = seqs_[n]->getSequence();
        for( size_t i = 0; i + W_ <= Lstored; i++ ){
            if( rRows_[n][Lstored - W_ - i] >= cutoff ){
                ofile_pos << seqs_[n]->getHeader() << '\t' << L << '\t' << ( ( i < L ) ? '+' : '-' ) << '\t'
                          << i + 1 << ".." << i + W_ << '\t';
                for( size_t b = i; b < i + W_; b++ ) ofile_pos << Alphabet::getBase( codes[b] );
                ofile_pos << '\n';
            }
        }
    }
}
