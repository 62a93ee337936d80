// SequenceSet / Sequence — the sequence store of the host side.
//
// Public interface mirrors the reference (src/init/Sequence.h:13-33, src/init/SequenceSet.h:15-32). The storage
// is different by design: one contiguous arena of stored codes for the whole set (forward | 0 | reverse complement
// per record unless single-stranded, src/init/Sequence.cpp:10-18,91-99), prefix offsets, and the short list of
// positions whose k-mer hash depends on rand() draws for a code-0 base (Sequence.cpp:35-41). That arena is what
// bamm_seqset_create() uploads; the reference's size_t kmer_[] array per sequence is never built unless a host-side
// consumer asks for it through getKmer().
//
// A Sequence is a light handle (set, index). FDR folds and filtered sets are vectors of such handles, which the
// EM / ScoreSeqSet wrappers translate into index subsets of the resident device set.
#ifndef BAMM_HOST_SEQUENCESET_H_
#define BAMM_HOST_SEQUENCESET_H_

#include <cstddef>
#include <cstdint>
#include <fstream>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "Alphabet.h"
#include "Device.h"

class SequenceSet;

class Sequence {
public:
    Sequence( SequenceSet* set, size_t index ) : set_( set ), index_( index ) {}

    uint8_t*        getSequence();          // stored codes (L entries)
    size_t          getL() const;           // stored length: 2*L0+1 on both strands
    std::string     getHeader() const;
    size_t*         getKmer();              // reference kmer_ (11-mer hash per position); materialised on first use
    size_t          kmerAt( size_t i );         // same value for one position, without materialising the array

    float           getIntensity() const    { return intensity_; }
    float           getWeight() const       { return weight_; }
    void            setIntensity( float v ) { intensity_ = v; }
    void            setWeight( float v )    { weight_ = v; }
    void            print();

    SequenceSet*    getSet() const          { return set_; }
    size_t          getIndex() const        { return index_; }

private:
    SequenceSet*    set_;
    size_t          index_;
    float           intensity_ = 0.0f;
    float           weight_ = 0.0f;
};

class SequenceSet {
public:
    // reads a FASTA file (reference: SequenceSet::readFASTA, src/init/SequenceSet.cpp:67-225)
    SequenceSet( std::string sequenceFilepath, bool singleStrand = false, std::string intensityFilepath = "" );
    // adopts already stored codes (e.g. a sampled negative set: single-stranded records, no N). All records share `header`.
    SequenceSet( std::vector<uint8_t> storedCodes, std::vector<uint64_t> offsets, std::string header );
    // incremental construction of a generated set: records are appended already in stored form; N draws (if any) are
    // made right after each record, like the reference's Sequence constructor does
    struct Build {};
    SequenceSet( Build, std::string header );
    // adopts a set that was CREATED ON THE DEVICE (bamm_seqset_sample_negatives): offsets come back now, the stored codes only
    // if a host-side consumer asks for them (getSequence(), codes(), getKmer())
    struct DeviceBuilt {};
    SequenceSet( DeviceBuilt, bamm_seqset* deviceSet, std::string header );
    void appendStoredRecord( const uint8_t* storedCodes, size_t L );
    void finishBuild();
    ~SequenceSet();
    SequenceSet( const SequenceSet& ) = delete;
    SequenceSet& operator=( const SequenceSet& ) = delete;

    std::string             getSequenceFilepath()   { return sequenceFilepath_; }
    std::string             getIntensityFilepath()  { return intensityFilepath_; }
    std::vector<Sequence*>  getSequences();
    size_t                  getMinL()               { return minL_; }
    size_t                  getMaxL()               { return maxL_; }
    float*                  getBaseFrequencies()    { return baseFrequencies_.data(); }
    void                    print();

    // ---- arena access (host wrappers, tests) ----
    size_t                  size() const            { return offsets_.size() - 1; }
    const std::vector<uint8_t>&  codes()            { ensureCodes(); return codes_; }
    const std::vector<uint64_t>& offsets() const    { return offsets_; }
    const std::vector<uint64_t>& patchPositions() const { return patchPos_; }
    const std::vector<uint64_t>& patchKmers() const { return patchKmer_; }
    const std::string&      headerOf( size_t n ) const { return headers_.empty() ? sharedHeader_ : headers_[n]; }
    size_t                  kmerAt( size_t n, size_t i );         // reference kmer_[i] of sequence n
    size_t*                 kmersOf( size_t n );    // materialises the whole set's kmer_ on first call

    // the set resident in HBM (created on first use; shared by every EM / ScoreSeqSet / BackgroundModel that uses it)
    bamm_seqset*            device();

    // maps a vector of handles onto (set, indices); exits with an error if the handles come from different sets
    static SequenceSet*     commonSet( const std::vector<Sequence*>& seqs, std::vector<uint64_t>& indices, bool* isWholeSet = nullptr );

private:
    friend class Sequence;
    void                    appendRecord( const std::string& header, const std::string& bases, bool singleStrand,
                                          std::vector<size_t>& baseCounts );
    bool                    readFastaDevice( std::ifstream& file, size_t bytes, bool singleStrand );   // false: too many undefined bases, use the host loop
    void                    drawPatches( uint64_t begin, uint64_t end );
    void                    finalize();
    void                    ensureCodes()           { if( codesOnDevice_ ) fetchCodes(); }
    void                    fetchCodes();           // device-built set: stored codes device -> host, once

    std::string             sequenceFilepath_;
    std::string             intensityFilepath_;
    std::vector<uint8_t>    codes_;
    std::vector<uint64_t>   offsets_;
    std::vector<std::string> headers_;
    std::string             sharedHeader_;
    std::vector<uint64_t>   patchPos_, patchKmer_;
    std::vector<Sequence>   handles_;
    size_t                  minL_ = 0, maxL_ = 0;
    std::vector<float>      baseFrequencies_;
    std::vector<size_t>     Y_;                     // A^0 .. A^11
    std::vector<size_t>     kmers_;                 // lazily materialised reference kmer_ for all positions
    std::once_flag          kmersOnce_;
    bamm_seqset*            device_ = nullptr;
    bool                    codesOnDevice_ = false; // device-built set whose codes_ has not been fetched yet
    std::mutex              codesMutex_;
    std::mutex              deviceMutex_;
};

inline uint8_t*     Sequence::getSequence()     { set_->ensureCodes(); return const_cast<uint8_t*>( set_->codes_.data() ) + set_->offsets_[index_]; }
inline size_t       Sequence::getL() const      { return static_cast<size_t>( set_->offsets_[index_ + 1] - set_->offsets_[index_] ); }
inline std::string  Sequence::getHeader() const { return set_->headerOf( index_ ); }
inline size_t*      Sequence::getKmer()         { return set_->kmersOf( index_ ); }
inline size_t       Sequence::kmerAt( size_t i )  { return set_->kmerAt( index_, i ); }

#endif
