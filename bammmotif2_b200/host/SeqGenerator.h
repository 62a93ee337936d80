// SeqGenerator — negative (background) sequence sampling for FDR / occurrence p-values.
// Covers the part of the reference's SeqGenerator that the BaMMmotif driver uses on this path
// (src/seq_generator/SeqGenerator.h:30-41; sample_bgseqset_by_fold, SeqGenerator.cpp:188-206). The sampled bases
// depend on the libc rand() stream after srand(42) (SeqGenerator.cpp:33-34, 285-348), so the draws are made in the
// same order with the same float comparisons; the result is bit-identical to the reference's negative set.
// Motif embedding / masking (BaMMSimu) is out of scope.
//
// Ownership differs from the reference by design: the sampled records are appended to ONE SequenceSet arena (which is
// what gets uploaded to the device) instead of a vector of individually allocated Sequence objects.
#ifndef BAMM_HOST_SEQGENERATOR_H_
#define BAMM_HOST_SEQGENERATOR_H_

#include <memory>
#include <vector>

#include "Motif.h"

class SeqGenerator {
public:
    SeqGenerator( std::vector<Sequence*> seqs, Motif* motif = NULL, size_t sOrder = 2, float q = 1.0f, bool genericNeg = false );
    ~SeqGenerator();

    // `fold` sampled sequences per input sequence, each as long as its (stored) template, single-stranded
    std::unique_ptr<SequenceSet> sample_bgseqset_by_fold( size_t fold );
    // negN sequences of length maxL from the set-wide k-mer frequencies
    std::unique_ptr<SequenceSet> sample_bgseqset_by_num( size_t negN, size_t maxL );

private:
    void calculate_kmer_frequency();
    void rescale_kmer_frequency( Sequence* refSeq );
    void count_kmers( Sequence* seq, std::vector<std::vector<size_t>>& n );
    void sample_into( std::vector<uint8_t>& sequence, size_t L );

    std::vector<Sequence*>              seqs_;
    size_t                              sOrder_;
    bool                                genericNeg_;
    std::vector<size_t>                 Y_;
    std::vector<float>                  A_;             // pseudo-count weights, 20 for every order
    std::vector<std::vector<float>>     v_, v_seq_, range_bar_;
    std::vector<std::vector<size_t>>    n_, n_seq_;
    bool                                kmer_freq_is_calculated_ = false;
    bool                                kmer_freq_is_rescaled_ = false;
};

#endif
