#include "SeqGenerator.h"

#include <cassert>
#include <cstdlib>
#include <iostream>

#include "Util.h"

SeqGenerator::SeqGenerator( std::vector<Sequence*> seqs, Motif*, size_t sOrder, float, bool genericNeg )
    : seqs_( seqs ), sOrder_( sOrder ), genericNeg_( genericNeg ){
    for( size_t k = 0; k < sOrder_ + 8; k++ ) Y_.push_back( util::ipow( Alphabet::getSize(), k ) );
    v_.resize( sOrder_ + 1 ); v_seq_.resize( sOrder_ + 1 ); range_bar_.resize( sOrder_ + 1 );
    n_.resize( sOrder_ + 1 ); n_seq_.resize( sOrder_ + 1 );
    for( size_t k = 0; k <= sOrder_; k++ ){
        v_[k].assign( Y_[k + 1], 0.f ); v_seq_[k].assign( Y_[k + 1], 0.f ); range_bar_[k].assign( Y_[k + 1], 0.f );
        n_[k].assign( Y_[k + 1], 0 ); n_seq_[k].assign( Y_[k + 1], 0 );
    }
    A_.assign( sOrder_ + 1, 20.f );
    util::srand42( 42 );            // reference: SeqGenerator.cpp:33-34 — the sampler always restarts the libc stream
}

SeqGenerator::~SeqGenerator(){}

// (k+1)-mer counts of one sequence for k <= sOrder, positions j >= k only (reference: SeqGenerator.cpp:76-88, 125-134)
void SeqGenerator::count_kmers( Sequence* seq, std::vector<std::vector<size_t>>& n ){
    const size_t L = seq->getL();
    const uint8_t* c = seq->getSequence();
    size_t lastZero = static_cast<size_t>( -1 );
    for( size_t j = 0; j < L; j++ ){
        if( c[j] == 0 ) lastZero = j;
        const bool patched = lastZero != static_cast<size_t>( -1 ) && j - lastZero <= 10;
        size_t full = 0;
        if( patched ){
            full = seq->kmerAt( j );                       // carries the rand() draws of the reference's kmer_
        } else {
            for( size_t t = 0; t <= sOrder_ && t <= j; t++ ) full += static_cast<size_t>( c[j - t] - 1 ) * Y_[t];
        }
        for( size_t k = 0; k <= sOrder_ && k <= j; k++ ) n[k][full % Y_[k + 1]]++;
    }
}

// reference: SeqGenerator::calculate_kmer_frequency, src/seq_generator/SeqGenerator.cpp:63-112
void SeqGenerator::calculate_kmer_frequency(){
    for( auto& nk : n_ ) std::fill( nk.begin(), nk.end(), 0 );
    for( Sequence* s : seqs_ ) count_kmers( s, n_ );

    size_t normFactor = 0;
    for( size_t y = 0; y < Y_[1]; y++ ) normFactor += n_[0][y];
    float sum = 0.0f;
    for( size_t y = 0; y < Y_[1]; y++ ){
        v_[0][y] = ( ( float )n_[0][y] + A_[0] * 0.25f ) / ( ( float )normFactor + A_[0] );
        sum += v_[0][y];
        range_bar_[0][y] = sum;
    }
    for( size_t k = 1; k <= sOrder_; k++ ){
        sum = 0.f;
        for( size_t y = 0; y < Y_[k + 1]; y++ ){
            const size_t yk = y / Y_[1], y2 = y % Y_[k];
            v_[k][y] = ( ( float )n_[k][y] + A_[k] * v_[k - 1][y2] ) / ( ( float )n_[k - 1][yk] + A_[k] );
            if( y % Y_[1] == 0 ) sum = 0.f;
            sum += v_[k][y];
            range_bar_[k][y] = sum;                         // cumulative distribution within each context
        }
    }
    kmer_freq_is_calculated_ = true;
}

// Conditional probabilities re-scaled towards the composition of one template sequence; orders 0..2 are hard-wired
// (reference: SeqGenerator::rescale_kmer_frequency, src/seq_generator/SeqGenerator.cpp:114-185). Mixed
// size_t / float arithmetic is kept operand by operand.
void SeqGenerator::rescale_kmer_frequency( Sequence* refSeq ){
    const size_t L = refSeq->getL();
    for( auto& nk : n_seq_ ) std::fill( nk.begin(), nk.end(), 0 );
    count_kmers( refSeq, n_seq_ );

    size_t k = 0;
    float sum = 0.f;
    for( size_t y = 0; y < Y_[k + 1]; y++ ){
        v_seq_[k][y] = v_[k][y];
        sum += v_seq_[k][y];
        range_bar_[k][y] = sum;
    }

    k = 1;
    for( size_t y = 0; y < Y_[k + 1]; y++ ){
        const size_t y2 = y % Y_[k];
        v_seq_[k][y] = v_[k][y] * ( n_seq_[k][y] + A_[k - 1] * v_[k - 1][y2] ) / v_[k - 1][y2] / ( L + A_[k - 1] );
    }
    std::vector<float> normFactors( Y_[k], 0.0f );
    for( size_t y = 0; y < Y_[k + 1]; y++ ){
        const size_t yk = y / Y_[1];
        v_seq_[k][y] = ( n_seq_[k][y] + A_[k] * v_seq_[k][y] ) / ( n_seq_[k - 1][yk] + A_[k] );
        normFactors[yk] += v_seq_[k][y];
    }
    for( size_t y = 0; y < Y_[k + 1]; y++ ) v_seq_[k][y] /= normFactors[y / Y_[1]];
    for( size_t y = 0; y < Y_[k + 1]; y++ ){
        if( y % Y_[1] == 0 ) sum = 0.0f;
        sum += v_seq_[k][y];
        range_bar_[k][y] = sum;
    }

    k = 2;
    for( size_t y = 0; y < Y_[k + 1]; y++ ){
        const size_t y2 = y % Y_[k], yk = y / Y_[1];
        v_seq_[k][y] = ( n_seq_[k][y] + A_[k] * v_seq_[k - 1][y2] ) / ( n_seq_[k - 1][yk] + A_[k] );
        if( y % Y_[1] == 0 ) sum = 0.0f;
        sum += v_seq_[k][y];
        range_bar_[k][y] = sum;
    }
    kmer_freq_is_rescaled_ = true;
}

// One Markov-chain sample of length L from range_bar_ (reference: SeqGenerator::bg_sequence /
// bgseq_on_rescaled_v, src/seq_generator/SeqGenerator.cpp:226-348): one util::rand31() per base, inverse-CDF lookup.
void SeqGenerator::sample_into( std::vector<uint8_t>& sequence, size_t L ){
    sequence.assign( L, 0 );
    const size_t A = Y_[1];
    float random = ( float )util::rand31() / ( float )RAND_MAX;
    for( uint8_t y = 0; y < A; y++ ){
        if( random <= range_bar_[0][y] ){ sequence[0] = y + 1; break; }
    }
    for( size_t i = 1; i < L; i++ ){
        const size_t order = i < sOrder_ ? i : sOrder_;            // the first bases use shorter contexts
        size_t yk = 0;
        for( size_t k = order; k > 0; k-- ) yk += ( sequence[i - k] - 1 ) * Y_[k];
        random = ( float )util::rand31() / ( float )RAND_MAX;
        for( size_t y = yk, a = 1; y < yk + A; y++, a++ ){
            sequence[i] = static_cast<uint8_t>( a );
            if( random <= range_bar_[order][y] ) break;
        }
    }
}

std::unique_ptr<SequenceSet> SeqGenerator::sample_bgseqset_by_fold( size_t fold ){
    if( !genericNeg_ && sOrder_ != 2 ){
        // the reference indexes orders 0..2 unconditionally in rescale_kmer_frequency (out of bounds for sOrder < 2,
        // ignores orders above 2)
        std::cerr << "Error: negative-set sampling on re-scaled frequencies supports --sOrder 2 only." << std::endl;
        exit( 1 );
    }
    // The library samples the identical negative set on the GPU (same k-mer models, same libc rand() stream re-created by
    // jump-ahead; bamm_seqset_sample_negatives). It declines (BAMM_E_STATE) only in the one-in-10^7-records case where the
    // reference's own draw order leaves the one-draw-per-base pattern (a first base above the last cumulative bar makes its
    // Sequence constructor consume extra draws); the reference's serial loop below then produces the set, so the result
    // is the reference's in every case. BAMM_HOST_NEGATIVES=1 (tests) forces that loop. No device => the call exits.
    if( !genericNeg_ && !getenv( "BAMM_HOST_NEGATIVES" ) ){
        std::vector<uint64_t> indices;
        bool whole = false;
        SequenceSet* tmpl = SequenceSet::commonSet( seqs_, indices, &whole );
        bamm_seqset* h = nullptr;
        const int rc = bamm_seqset_sample_negatives( tmpl->device(), whole ? NULL : indices.data(), indices.size(), fold, 42, &h );
        if( rc == BAMM_OK ){
            util::srand42( 42 );                                        // the stream position after sampling is never consumed (FDR re-seeds, FDR.cpp:153)
            return std::unique_ptr<SequenceSet>( new SequenceSet( SequenceSet::DeviceBuilt(), h, "> bg_seq" ) );
        }
        if( rc != BAMM_E_STATE ){
            std::cerr << "Error: " << bamm_last_error() << std::endl;
            exit( 1 );
        }
    }
    std::unique_ptr<SequenceSet> negset( new SequenceSet( SequenceSet::Build(), "> bg_seq" ) );
    calculate_kmer_frequency();
    std::vector<uint8_t> sequence;
    for( size_t i = 0; i < seqs_.size(); i++ ){
        for( size_t n = 0; n < fold; n++ ){
            if( !genericNeg_ ) rescale_kmer_frequency( seqs_[i] );
            assert( kmer_freq_is_calculated_ );
            sample_into( sequence, seqs_[i]->getL() );
            negset->appendStoredRecord( sequence.data(), sequence.size() );
        }
    }
    negset->finishBuild();
    return negset;
}

std::unique_ptr<SequenceSet> SeqGenerator::sample_bgseqset_by_num( size_t negN, size_t maxL ){
    std::unique_ptr<SequenceSet> negset( new SequenceSet( SequenceSet::Build(), "> bg_seq" ) );
    calculate_kmer_frequency();
    std::vector<uint8_t> sequence;
    for( size_t n = 0; n < negN; n++ ){
        sample_into( sequence, maxL );
        negset->appendStoredRecord( sequence.data(), sequence.size() );
    }
    negset->finishBuild();
    return negset;
}
