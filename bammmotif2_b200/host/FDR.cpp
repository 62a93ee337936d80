#include "FDR.h"
#include "Util.h"

#include <chrono>
#include <algorithm>
#include <cassert>
#include <cmath>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <limits>
#include <cstdio>
#include <string>
#include <thread>

// Score vectors of cross-validation runs reach 10^7 (ZOOPS) to 10^10 (MOPS) entries: those are sorted by the device radix
// sort behind bamm_sort_scores; a few thousand values are not worth the PCIe round trip. Same result either way.
static void sortScores( std::vector<float>& v, bool descending ){
    if( v.size() >= ( 1u << 16 ) ){
        BAMM_CHECK( bamm_sort_scores( v.data(), v.size(), descending ? 1 : 0 ) );
    } else if( descending ){
        std::sort( v.begin(), v.end(), std::greater<float>() );
    } else {
        std::sort( v.begin(), v.end(), std::less<float>() );
    }
}

FDR::FDR( std::vector<Sequence*> posSeqs, std::vector<Sequence*> negSeqs, Motif* motif, BackgroundModel* bgModel, size_t cvFold,
          bool mops, bool zoops, bool savePRs, bool savePvalues, bool saveLogOdds )
    : posSeqs_( posSeqs ), negSeqs_( negSeqs ), q_( motif ? motif->getQ() : 0.f ), motif_( motif ), bgModel_( bgModel ), cvFold_( cvFold ),
      mops_( mops ), zoops_( zoops ), savePRs_( savePRs ), savePvalues_( savePvalues ), saveLogOdds_( saveLogOdds ){}

FDR::~FDR(){}

// reference: FDR::evaluateMotif, src/evaluation/FDR.cpp:28-145.
// Split: positives are dealt round-robin — of every complete block of cvFold consecutive sequences, member `fold` is
// held out and the others train; a trailing incomplete block is dropped. Every fold scores the same negative subset,
// every cvFold-th negative. The reference runs the folds on OpenMP threads and appends scores in completion order; the
// statistics sort the scores first, so fold order does not matter. Folds run one after the other here: each already
// fills the whole GPU, and the device objects they share are read-only.
void FDR::evaluateMotif( bool EMoptimize, bool CGSoptimize, bool optimizeQ, bool advanceEM, float frac, size_t ){
    if( CGSoptimize && !EMoptimize ){
        std::cerr << "Error: collapsed Gibbs sampling (--CGS) is not part of the B200 path." << std::endl;
        exit( 1 );
    }
    float updatedQ = q_;
    std::vector<Sequence*> negSet;
    for( size_t n = 0; n + cvFold_ <= negSeqs_.size(); n += cvFold_ ) negSet.push_back( negSeqs_[n] );

    // BAMM_TRACE: wall time of the phases of every fold on stderr
    const bool trace = getenv( "BAMM_TRACE" ) != NULL;
    auto t_mark = std::chrono::steady_clock::now();
    auto mark = [&]( size_t fold, const char* what ){
        if( !trace ) return;
        const auto t = std::chrono::steady_clock::now();
        std::cerr << "[bamm host] fold " << fold << ": " << what << " " << std::chrono::duration<double, std::milli>( t - t_mark ).count() << " ms" << std::endl;
        t_mark = t;
    };
    for( size_t fold = 0; fold < cvFold_; fold++ ){
        Motif* motif = new Motif( *motif_ );
        std::vector<Sequence*> testSet, trainSet;
        for( size_t n = 0; n + cvFold_ <= posSeqs_.size(); n += cvFold_ ){
            for( size_t f = 0; f < cvFold_; f++ ){
                ( f != fold ? trainSet : testSet ).push_back( posSeqs_[n + f] );
            }
        }
        mark( fold, "train / test split" );
        if( EMoptimize ){
            EM model( motif, bgModel_, trainSet, optimizeQ, false, frac );
            mark( fold, "EM object" );
            if( advanceEM ) model.mask(); else model.optimize();
            updatedQ = model.getQ();
            mark( fold, "EM::optimize" );
        }
        mark( fold, "EM object released" );
        ScoreSeqSet score_testset( motif, bgModel_, testSet );
        score_testset.setKeepMops( mops_ );
        score_testset.calcLogOdds();
        mark( fold, "scores of the held-out positives" );
        ScoreSeqSet score_negset( motif, bgModel_, negSet );
        score_negset.setKeepMops( mops_ );
        score_negset.calcLogOdds();
        mark( fold, "scores of the negatives" );

        if( mops_ ){
            const std::vector<float>& p = score_testset.flatMopsScores();
            posScoreAll_.insert( posScoreAll_.end(), p.begin(), p.end() );
            const std::vector<float>& g = score_negset.flatMopsScores();
            negScoreAll_.insert( negScoreAll_.end(), g.begin(), g.end() );
        }
        if( zoops_ ){
            std::vector<float> z = score_testset.getZoopsScores();
            posScoreMax_.insert( posScoreMax_.end(), z.begin(), z.end() );
            z = score_negset.getZoopsScores();
            negScoreMax_.insert( negScoreMax_.end(), z.begin(), z.end() );
        }
        delete motif;
        mark( fold, "scores appended" );
    }
    q_ = updatedQ;
    calculatePR();
    if( savePvalues_ ){
        fprintf( stderr, " ______________________\n|                      |\n|  calculate P-values  |\n|______________________|\n\n" );
        calculatePvalues();
    }
}

namespace {
// The reference's merge walk reads one element past the end of a score vector once that vector is exhausted
// (FDR.cpp:228-238). Such a read can never decide a branch when the vector holds posN (negN) scores, so it is served
// as NaN (every comparison false) instead of touching foreign memory.
inline float at( const std::vector<float>& v, size_t i ){
    return i < v.size() ? v[i] : std::numeric_limits<float>::quiet_NaN();
}
}

// reference: FDR::calculatePR, src/evaluation/FDR.cpp:147-276
void FDR::calculatePR(){
    const size_t posN = posSeqs_.size();
    const size_t negN = negSeqs_.size();
    const float mFold = ( float )negN / ( float )posN;

    util::srand42( 42 );                                    // tie-breaks below draw from a fresh libc stream (FDR.cpp:153)

    if( mops_ ){
        sortScores( posScoreAll_, true );
        sortScores( negScoreAll_, true );
        size_t idx_posAll = 0, idx_negAll = 0;
        float E_TP_MOPS = 0.0f;
        size_t idx_max = posN + negN;
        const size_t len_all = posScoreAll_.size() + negScoreAll_.size();
        for( size_t i = 0; i < len_all; i++ ){
            if( at( posScoreAll_, idx_posAll ) > at( negScoreAll_, idx_negAll ) || idx_negAll == len_all ) idx_posAll++;
            else idx_negAll++;
            MOPS_TP_.push_back( ( float )idx_posAll - ( float )idx_negAll / mFold );
            MOPS_FP_.push_back( ( float )idx_negAll / mFold );
            if( E_TP_MOPS == MOPS_TP_[i] ) idx_max = i;
            if( E_TP_MOPS < MOPS_TP_[i] ) E_TP_MOPS = MOPS_TP_[i];
        }
        for( size_t i = 0; i < idx_max && i < MOPS_TP_.size(); i++ ){
            MOPS_FDR_.push_back( MOPS_FP_[i] / ( MOPS_TP_[i] + MOPS_FP_[i] ) );
            MOPS_Rec_.push_back( MOPS_TP_[i] / E_TP_MOPS );
        }
        occ_mult_ = E_TP_MOPS / ( float )posN;
    }

    if( zoops_ ){
        PN_Pvalue_.clear();
        sortScores( posScoreMax_, true );
        sortScores( negScoreMax_, true );

        size_t idx_posMax = 0, idx_negMax = 0, min_idx_pos = 0;
        const size_t posN_est = static_cast<size_t>( q_ * ( float )posN );
        const size_t n_top = std::fmin( 100, negN / 10 );

        float lambda = 1e-16f;                      // rate of the exponential tail fitted to the top negative scores
        for( size_t l = 0; l < n_top; l++ ) lambda += negScoreMax_[l] - negScoreMax_[n_top];
        lambda /= n_top;
        assert( lambda > 0.f );

        float Sl = 0.f;
        ZOOPS_TP_.reserve( posN + negN ); ZOOPS_FP_.reserve( posN + negN ); ZOOPS_FDR_.reserve( posN + negN );
        ZOOPS_Rec_.reserve( posN + negN ); PN_Pvalue_.reserve( posN + negN );
        // the two neighbours of Sl among the sorted negatives (std::lower_bound / std::upper_bound with std::greater in the
        // reference, FDR.cpp:245-247): the walk hands out scores in descending order, so both positions only move forward and
        // are advanced instead of searched; a score above the previous one (the forced first positive) searches again
        size_t lb = 0, ub = 0;
        const size_t nNeg = negScoreMax_.size();
        float lastSl = 0.f;
        bool searched = false;
        for( size_t i = 0; i < posN + negN; i++ ){
            const float ps = at( posScoreMax_, idx_posMax ), ns = at( negScoreMax_, idx_negMax );
            if( ( ps > ns || idx_posMax == 0 || idx_negMax == negN ) && idx_posMax < posN ){
                Sl = ps;
                idx_posMax++;
            } else if( ps == ns && util::rand31() % 2 == 0 && idx_posMax < posN ){
                Sl = ps;
                idx_posMax++;
            } else {
                Sl = ns;
                idx_negMax++;
            }
            const float TP = ( float )idx_posMax;
            const float FP = ( float )idx_negMax / mFold;
            ZOOPS_TP_.push_back( TP );
            ZOOPS_FP_.push_back( FP );

            float p_value;
            if( Sl <= negScoreMax_[n_top] ){
                // rank among the negatives, interpolated between the neighbouring negative scores
                if( !searched || Sl > lastSl ){
                    lb = static_cast<size_t>( std::lower_bound( negScoreMax_.begin(), negScoreMax_.end(), Sl, std::greater<float>() ) - negScoreMax_.begin() );
                    ub = static_cast<size_t>( std::upper_bound( negScoreMax_.begin(), negScoreMax_.end(), Sl, std::greater<float>() ) - negScoreMax_.begin() );
                    searched = true;
                } else {
                    while( lb < nNeg && negScoreMax_[lb] > Sl ) lb++;           // first negative that is not above Sl
                    while( ub < nNeg && !( Sl > negScoreMax_[ub] ) ) ub++;      // first negative below Sl
                }
                lastSl = Sl;
                const float Sl_upper = negScoreMax_[lb ? lb - 1 : 0];           // lb > 0 unless the top negatives all equal Sl (the reference reads before the array then)
                // (a score below every negative has no lower neighbour; the reference reads one past the end there —
                //  the score itself is used instead, which puts the p-value at the top of the range)
                const float Sl_lower = ub < nNeg ? negScoreMax_[ub] : Sl;
                p_value = ( idx_negMax + ( Sl_upper - Sl ) / ( Sl_upper - Sl_lower + 1e-5 ) ) / ( float )negN;
            } else {
                p_value = n_top * expf( ( negScoreMax_[n_top] - Sl ) / lambda ) / negN;
            }
            PN_Pvalue_.push_back( p_value );
            if( idx_posMax == posN_est ) min_idx_pos = i;
            ZOOPS_FDR_.push_back( FP / ( TP + FP ) );
            ZOOPS_Rec_.push_back( TP / ( float )posN );
        }
        occ_frac_ = 1.0f - ZOOPS_FP_[min_idx_pos] / ( float )posN;
    }
}

// reference: FDR::calculatePvalues, src/evaluation/FDR.cpp:278-332
void FDR::calculatePvalues(){
    auto rank = []( std::vector<float>& neg, std::vector<float>& pos, std::vector<float>& out ){
        sortScores( neg, false );
        sortScores( pos, false );
        for( size_t i = 0; i < pos.size(); i++ ){
            const size_t low = std::distance( neg.begin(), std::lower_bound( neg.begin(), neg.end(), pos[i] ) );
            const size_t up = std::distance( neg.begin(), std::upper_bound( neg.begin(), neg.end(), pos[i] ) );
            float p = 1.0f - ( float )( up + low ) / ( 2.0f * ( float )neg.size() );
            if( p < 1.e-6 ) p = 1.e-6;
            if( p > 1.0f ) p = 1.0f;
            out.push_back( p );
        }
    };
    if( mops_ ) rank( negScoreAll_, posScoreAll_, MOPS_Pvalue_ );
    if( zoops_ ) rank( negScoreMax_, posScoreMax_, ZOOPS_Pvalue_ );
}

void FDR::print(){}

// Rows of floats, tab-separated with a tab before the newline (the layout of the reference's statistics files), formatted by
// several threads: "%g" prints what operator<<(float) prints with the default precision of 6. The statistics file of a
// cross-validation over 10^6 positives has 1.1 * 10^7 rows; a single formatting thread is most of the run's wall clock there.
static void writeFloatRows( std::ofstream& out, const std::vector<const std::vector<float>*>& cols ){
    size_t rows = cols.empty() ? 0 : cols[0]->size();
    for( const std::vector<float>* c : cols ) rows = std::min( rows, c->size() );
    if( rows == 0 ) return;
    size_t nt = std::min<size_t>( std::max( 1u, std::thread::hardware_concurrency() ), 16 );
    if( rows < 200000 ) nt = 1;
    std::vector<std::string> parts( nt );
    auto work = [&]( size_t t ){
        const size_t a = rows * t / nt, b = rows * ( t + 1 ) / nt;
        std::string& o = parts[t];
        o.reserve( ( b - a ) * ( cols.size() * 12 + 1 ) );
        char buf[32];
        for( size_t i = a; i < b; i++ ){
            for( const std::vector<float>* c : cols ){
                const int n = snprintf( buf, sizeof( buf ), "%g", static_cast<double>( ( *c )[i] ) );
                o.append( buf, static_cast<size_t>( n ) );
                o.push_back( '\t' );
            }
            o.push_back( '\n' );
        }
    };
    std::vector<std::thread> th;
    for( size_t t = 1; t < nt; t++ ) th.emplace_back( work, t );
    work( 0 );
    for( std::thread& t : th ) t.join();
    for( const std::string& o : parts ) out.write( o.data(), static_cast<std::streamsize>( o.size() ) );
}

// file formats: reference FDR::write, src/evaluation/FDR.cpp:338-450 (rows end in '\n', not std::endl: the same bytes without a
// flush per row — the statistics files have one row per scored sequence)
void FDR::write( char* odir, std::string basename ){
    const std::string opath = std::string( odir ) + '/' + basename;
    if( savePRs_ ){
        if( zoops_ ){
            std::ofstream out( opath + ".zoops.stats" );
            out << "TP" << '\t' << "FP" << '\t' << "FDR" << '\t' << "Recall" << '\t' << "p-value" << '\t'
                << ( float )negSeqs_.size() / ( float )posSeqs_.size() << '\t' << occ_frac_ << '\n';
            writeFloatRows( out, { &ZOOPS_TP_, &ZOOPS_FP_, &ZOOPS_FDR_, &ZOOPS_Rec_, &PN_Pvalue_ } );
        }
        if( mops_ ){
            std::ofstream out( opath + ".mops.stats" );
            out << "TP" << '\t' << "FP" << '\t' << "FDR" << '\t' << "Recall" << '\t' << occ_mult_ << '\n';
            writeFloatRows( out, { &MOPS_TP_, &MOPS_FP_, &MOPS_FDR_, &MOPS_Rec_ } );
        }
    }
    if( savePvalues_ ){
        if( zoops_ ){
            std::ofstream out( opath + ".zoops.pvalues" );
            for( float p : ZOOPS_Pvalue_ ) out << std::setprecision( 3 ) << p << '\n';
        }
        if( mops_ ){
            std::ofstream out( opath + ".mops.pvalues" );
            for( float p : MOPS_Pvalue_ ) out << std::setprecision( 3 ) << p << '\n';
        }
    }
    if( saveLogOdds_ ){
        if( zoops_ ){
            std::ofstream out( opath + ".zoops.logOdds" );
            out << "positive" << '\t' << "negative" << '\n';
            for( size_t i = 0; i < posScoreMax_.size(); i++ ){
                out << std::setprecision( 6 ) << posScoreMax_[i] << '\t' << at( negScoreMax_, i * negSeqs_.size() / posSeqs_.size() ) << '\n';
            }
        }
        if( mops_ ){
            std::ofstream out( opath + ".mops.logOdds" );
            out << "positive" << '\t' << "negative" << '\n';
            for( size_t i = 0; i < posScoreAll_.size(); i++ ){
                out << std::setprecision( 6 ) << posScoreAll_[i] << '\t' << at( negScoreAll_, i * negSeqs_.size() / posSeqs_.size() ) << '\n';
            }
        }
    }
}

void FDR::saveUnsortedLogOdds( std::string opath, std::vector<float> logOdds ){
    std::ofstream ofile( opath );
    for( size_t i = 0; i < logOdds.size(); i++ ) ofile << i + 1 << '\t' << std::setprecision( 6 ) << logOdds[i] << '\n';
}
