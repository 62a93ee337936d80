#include "SequenceSet.h"
#include "Util.h"

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>

namespace {
inline size_t ipow( size_t base, size_t exp ){
    size_t r = 1;
    while( exp-- ) r *= base;
    return r;
}
}

SequenceSet::SequenceSet( std::string sequenceFilepath, bool singleStrand, std::string intensityFilepath )
    : sequenceFilepath_( sequenceFilepath ), intensityFilepath_( intensityFilepath ){

    if( !intensityFilepath.empty() ){
        // reference: SequenceSet::readIntensities is a stub that exits (src/init/SequenceSet.cpp:227-231)
        std::cerr << "Error: sequenceSet::readIntensities() is not implemented so far." << std::endl;
        std::exit( 1 );
    }
    for( size_t k = 0; k < 12; k++ ) Y_.push_back( ipow( Alphabet::getSize(), k ) );
    offsets_.push_back( 0 );

    std::ifstream file( sequenceFilepath_.c_str() );
    if( !file.is_open() ){
        std::cerr << "Error: Cannot open FASTA file: " << sequenceFilepath_ << std::endl;
        std::exit( 1 );
    }
    {   // large files: the bases are encoded on the device (readFastaDevice); BAMM_DEVICE_FASTA=1 / 0 forces either reader
        file.seekg( 0, std::ios::end );
        const std::streamoff bytes = file.tellg();
        file.seekg( 0, std::ios::beg );
        const char* force = getenv( "BAMM_DEVICE_FASTA" );
        if( force ? atoi( force ) != 0 : bytes >= ( std::streamoff( 1 ) << 26 ) ){
            if( readFastaDevice( file, static_cast<size_t>( bytes ), singleStrand ) ) return;
            // more than 1/16 of the bases undefined (the device keeps a bounded list of them): this reader handles the file
            file.clear();
            file.seekg( 0, std::ios::beg );
            offsets_.assign( 1, 0 );
            headers_.clear();
        }
    }
    std::vector<size_t> baseCounts( Alphabet::getSize(), 0 );
    size_t maxL = 0, minL = std::numeric_limits<size_t>::max();
    std::string line, header, bases;
    {   // one allocation for the code arena instead of repeated growth: the stored codes are at most the file size
        // (single strand) or twice the file size (both strands: forward | 0 | reverse complement, headers >= 2 bytes)
        file.seekg( 0, std::ios::end );
        const std::streamoff bytes = file.tellg();
        file.seekg( 0, std::ios::beg );
        if( bytes > 0 ) codes_.reserve( static_cast<size_t>( bytes ) * ( singleStrand ? 1 : 2 ) );
    }
    // a record is flushed when the next '>' line (or the end of the file) is reached; a header without bases is dropped
    auto flush = [&](){
        if( header.empty() ) return;
        if( bases.empty() ){
            std::cerr << "Warning: Ignore FASTA entry without sequence: " << sequenceFilepath_ << std::endl;
        } else {
            maxL = std::max( maxL, bases.length() );
            minL = std::min( minL, bases.length() );
            appendRecord( header, bases, singleStrand, baseCounts );
            bases.clear();
        }
        header.clear();
    };
    while( std::getline( file, line ) ){
        if( line.empty() ) continue;
        if( line[0] == '>' ){
            flush();
            if( line.length() == 1 ){
                header = ">";
            } else {
                // header runs to the first tab, and loses a trailing carriage return (reference :137-139)
                header = line.substr( 0, line.find( '\t' ) );
                header = header.substr( 0, header.find( '\r' ) );
            }
        } else if( !header.empty() ){
            if( line.find( ' ' ) != std::string::npos ){
                std::cerr << "Error: FASTA sequence contains space character: " << sequenceFilepath_ << std::endl;
                std::exit( 1 );
            }
            bases += line;
        } else {
            std::cerr << "Error: Wrong FASTA format: " << sequenceFilepath_ << std::endl;
            std::exit( 1 );
        }
    }
    flush();
    minL_ = minL;
    maxL_ = maxL;

    size_t total = 0;
    for( size_t c : baseCounts ) total += c;
    baseFrequencies_.resize( Alphabet::getSize() );
    for( size_t a = 0; a < baseCounts.size(); a++ ){
        baseFrequencies_[a] = static_cast<float>( baseCounts[a] ) / static_cast<float>( total );
    }
    finalize();
}

SequenceSet::SequenceSet( std::vector<uint8_t> storedCodes, std::vector<uint64_t> offsets, std::string header )
    : codes_( std::move( storedCodes ) ), offsets_( std::move( offsets ) ), sharedHeader_( std::move( header ) ){
    for( size_t k = 0; k < 12; k++ ) Y_.push_back( ipow( Alphabet::getSize(), k ) );
    if( offsets_.empty() ) offsets_.push_back( 0 );
    size_t maxL = 0, minL = std::numeric_limits<size_t>::max();
    for( size_t n = 0; n + 1 < offsets_.size(); n++ ){
        const size_t L = static_cast<size_t>( offsets_[n + 1] - offsets_[n] );
        maxL = std::max( maxL, L );
        minL = std::min( minL, L );
        drawPatches( offsets_[n], offsets_[n + 1] );       // no-op (and no rand() call) when the record has no code 0
    }
    minL_ = minL;
    maxL_ = maxL;
    baseFrequencies_.assign( Alphabet::getSize(), 0.0f );
    finalize();
}

SequenceSet::SequenceSet( DeviceBuilt, bamm_seqset* deviceSet, std::string header ) : sharedHeader_( std::move( header ) ){
    for( size_t k = 0; k < 12; k++ ) Y_.push_back( ipow( Alphabet::getSize(), k ) );
    uint64_t nseq = 0, npos = 0; int A = 0;
    BAMM_CHECK( bamm_seqset_info( deviceSet, &nseq, &npos, &A ) );
    offsets_.assign( nseq + 1, 0 );
    BAMM_CHECK( bamm_seqset_get_offsets( deviceSet, offsets_.data() ) );
    size_t maxL = 0, minL = std::numeric_limits<size_t>::max();
    for( size_t n = 0; n < nseq; n++ ){
        const size_t L = static_cast<size_t>( offsets_[n + 1] - offsets_[n] );
        maxL = std::max( maxL, L );
        minL = std::min( minL, L );
    }
    minL_ = minL;
    maxL_ = maxL;
    baseFrequencies_.assign( Alphabet::getSize(), 0.0f );
    device_ = deviceSet;
    codesOnDevice_ = true;
    finalize();
}

void SequenceSet::fetchCodes(){
    std::lock_guard<std::mutex> guard( codesMutex_ );
    if( !codesOnDevice_ ) return;
    codes_.resize( static_cast<size_t>( offsets_.back() ) );
    BAMM_CHECK( bamm_seqset_get_codes( device_, codes_.data() ) );
    codesOnDevice_ = false;
}

SequenceSet::SequenceSet( Build, std::string header ) : sharedHeader_( std::move( header ) ){
    for( size_t k = 0; k < 12; k++ ) Y_.push_back( ipow( Alphabet::getSize(), k ) );
    offsets_.push_back( 0 );
    minL_ = std::numeric_limits<size_t>::max();
    baseFrequencies_.assign( Alphabet::getSize(), 0.0f );
}

void SequenceSet::appendStoredRecord( const uint8_t* storedCodes, size_t L ){
    const uint64_t begin = codes_.size();
    codes_.insert( codes_.end(), storedCodes, storedCodes + L );
    offsets_.push_back( codes_.size() );
    maxL_ = std::max( maxL_, L );
    minL_ = std::min( minL_, L );
    drawPatches( begin, codes_.size() );
}

void SequenceSet::finishBuild(){
    handles_.clear();
    finalize();
}

SequenceSet::~SequenceSet(){
    if( device_ ) bamm_seqset_destroy( device_ );
}

void SequenceSet::appendRecord( const std::string& header, const std::string& bases, bool singleStrand,
                                std::vector<size_t>& baseCounts ){
    const size_t L0 = bases.length();
    const uint64_t begin = codes_.size();
    codes_.resize( begin + ( singleStrand ? L0 : 2 * L0 + 1 ), 0 );
    uint8_t* dst = codes_.data() + begin;
    for( size_t i = 0; i < L0; i++ ){
        const uint8_t c = Alphabet::getCode( bases[i] );
        dst[i] = c;
        if( c ) baseCounts[c - 1]++;                       // undefined bases are not counted (reference :99-108)
    }
    if( !singleStrand ){
        // forward | 0 | reverse complement  (reference: Sequence::appendRevComp, src/init/Sequence.cpp:91-99)
        for( size_t i = 0; i < L0; i++ ) dst[2 * L0 - i] = Alphabet::getComplementCode( dst[i] );
    }
    headers_.push_back( header );
    offsets_.push_back( codes_.size() );
    drawPatches( begin, codes_.size() );
}

// The FASTA reader with the per-base work on the device (bamm_seqset_encode_text): the host finds lines and headers with
// the same rules as the loop above (reference: SequenceSet::readFASTA, src/init/SequenceSet.cpp:67-225), ships the raw
// text, gets the base counts back, and draws the rand()-dependent k-mer hashes around undefined bases from the 21-code
// neighbourhoods the device returns — the same draws in the same order as drawPatches() makes from the host arena.
// The stored codes stay on the device until a host consumer asks for them.
bool SequenceSet::readFastaDevice( std::ifstream& file, size_t bytes, bool singleStrand ){
    // BAMM_TRACE: wall time of the reader's phases on stderr
    const bool trace = getenv( "BAMM_TRACE" ) != NULL;
    auto t_mark = std::chrono::steady_clock::now();
    auto mark = [&]( const char* what ){
        if( !trace ) return;
        const auto t = std::chrono::steady_clock::now();
        std::cerr << "[bamm host] FASTA reader: " << what << " " << std::chrono::duration<double, std::milli>( t - t_mark ).count() << " ms" << std::endl;
        t_mark = t;
    };
    // the text: the file mapped read-only (no copy, no zero-filled buffer); read() into a buffer when it cannot be mapped
    std::string text;
    const char* t = nullptr;
    size_t n = 0;
    struct Mapping {
        void* p = MAP_FAILED; size_t len = 0; int fd = -1;
        ~Mapping(){ if( p != MAP_FAILED ) munmap( p, len ); if( fd >= 0 ) close( fd ); }
    } map;
    map.fd = open( sequenceFilepath_.c_str(), O_RDONLY );
    if( map.fd >= 0 ){
        map.p = mmap( nullptr, bytes, PROT_READ, MAP_PRIVATE | MAP_POPULATE, map.fd, 0 );
        if( map.p != MAP_FAILED ){ map.len = bytes; t = static_cast<const char*>( map.p ); n = bytes; }
    }
    if( !t ){
        text.assign( bytes, '\0' );
        file.read( &text[0], static_cast<std::streamsize>( bytes ) );
        text.resize( static_cast<size_t>( file.gcount() ) );
        t = text.data();
        n = text.size();
    }
    mark( "file read" );

    std::vector<bamm_fasta_seg> segs, pending;
    std::vector<uint32_t> recL0;
    std::string header;
    bool haveHeader = false;
    uint64_t L0 = 0;
    size_t maxL = 0, minL = std::numeric_limits<size_t>::max();
    auto flush = [&](){
        if( !haveHeader ) return;
        if( L0 == 0 ){
            std::cerr << "Warning: Ignore FASTA entry without sequence: " << sequenceFilepath_ << std::endl;
        } else {
            if( L0 >= ( 1ull << 31 ) ){ std::cerr << "Error: sequence too long: " << sequenceFilepath_ << std::endl; std::exit( 1 ); }
            const uint32_t rec = static_cast<uint32_t>( recL0.size() );
            for( bamm_fasta_seg& sg : pending ){ sg.rec = rec; segs.push_back( sg ); }
            recL0.push_back( static_cast<uint32_t>( L0 ) );
            headers_.push_back( header );
            offsets_.push_back( offsets_.back() + ( singleStrand ? L0 : 2 * L0 + 1 ) );
            maxL = std::max<size_t>( maxL, L0 );
            minL = std::min<size_t>( minL, L0 );
        }
        pending.clear();
        L0 = 0;
        haveHeader = false;
    };
    for( size_t pos = 0; pos < n; ){
        const char* nl = static_cast<const char*>( memchr( t + pos, '\n', n - pos ) );
        const size_t end = nl ? static_cast<size_t>( nl - t ) : n;
        const size_t len = end - pos;
        if( len ){
            if( t[pos] == '>' ){
                flush();
                if( len == 1 ){
                    header = ">";
                } else {
                    const char* tab = static_cast<const char*>( memchr( t + pos, '\t', len ) );
                    size_t hl = tab ? static_cast<size_t>( tab - ( t + pos ) ) : len;
                    const char* cr = static_cast<const char*>( memchr( t + pos, '\r', hl ) );
                    if( cr ) hl = static_cast<size_t>( cr - ( t + pos ) );
                    header.assign( t + pos, hl );
                }
                haveHeader = true;
            } else if( haveHeader ){
                if( memchr( t + pos, ' ', len ) ){
                    std::cerr << "Error: FASTA sequence contains space character: " << sequenceFilepath_ << std::endl;
                    std::exit( 1 );
                }
                bamm_fasta_seg sg; sg.text_off = pos; sg.len = static_cast<uint32_t>( len ); sg.rec = 0; sg.dst = L0;
                pending.push_back( sg );
                L0 += len;
            } else {
                std::cerr << "Error: Wrong FASTA format: " << sequenceFilepath_ << std::endl;
                std::exit( 1 );
            }
        }
        pos = end + 1;
    }
    flush();
    mark( "lines + headers" );
    // lengths of the records as read (one strand), as the host reader reports them
    maxL_ = recL0.empty() ? 0 : maxL;
    minL_ = recL0.empty() ? std::numeric_limits<size_t>::max() : minL;

    const size_t A = Alphabet::getSize();
    uint8_t lut[256], comp[256];
    for( int c = 0; c < 256; c++ ){ lut[c] = Alphabet::getCode( static_cast<char>( c ) ); comp[c] = Alphabet::getComplementCode( static_cast<uint8_t>( c ) ); }
    std::vector<uint64_t> counts( A, 0 );
    uint64_t nzeroFwd = 0;
    const size_t N = recL0.size();
    {
        const int rc = bamm_seqset_encode_text( t, n, segs.data(), segs.size(), offsets_.data(), recL0.data(), N, singleStrand ? 1 : 0,
                                                static_cast<int>( A ), lut, comp, counts.data(), &nzeroFwd, &device_ );
        if( rc == BAMM_E_INVALID && std::string( bamm_last_error() ).find( "undefined" ) != std::string::npos ){
            device_ = nullptr;
            return false;
        }
        if( rc != BAMM_OK ){
            std::cerr << "Error: " << bamm_last_error() << std::endl;
            std::exit( 1 );
        }
    }
    mark( "bamm_seqset_encode_text" );
    size_t total = 0;
    for( uint64_t c : counts ) total += c;
    baseFrequencies_.resize( A );
    for( size_t a = 0; a < A; a++ ) baseFrequencies_[a] = static_cast<float>( counts[a] ) / static_cast<float>( total );

    // undefined bases in ascending stored position: the forward ones the device found, plus the structural N of each record
    std::vector<uint64_t> zpos( nzeroFwd );
    BAMM_CHECK( bamm_seqset_forward_zeros( device_, zpos.data() ) );
    std::sort( zpos.begin(), zpos.end() );
    if( !singleStrand ){
        std::vector<uint64_t> merged;
        merged.reserve( zpos.size() + N );
        size_t zi = 0;
        for( size_t r = 0; r < N; r++ ){
            const uint64_t mid = offsets_[r] + recL0[r];
            while( zi < zpos.size() && zpos[zi] < mid ) merged.push_back( zpos[zi++] );
            merged.push_back( mid );
        }
        while( zi < zpos.size() ) merged.push_back( zpos[zi++] );
        zpos.swap( merged );
    }
    std::vector<uint64_t> zbeg( zpos.size() ), zend( zpos.size() );
    {
        size_t r = 0;
        for( size_t k = 0; k < zpos.size(); k++ ){
            while( offsets_[r + 1] <= zpos[k] ) r++;
            zbeg[k] = offsets_[r]; zend[k] = offsets_[r + 1];
        }
    }
    std::vector<uint8_t> win( zpos.size() * 21 );
    BAMM_CHECK( bamm_seqset_code_windows( device_, zpos.data(), zbeg.data(), zend.data(), zpos.size(), win.data() ) );
    mark( "undefined bases + code windows" );
    // the draws of drawPatches(), from the windows: records in order, undefined bases ascending, positions z..z+10 not yet
    // hashed, bases of each 11-mer from the oldest to the newest
    patchPos_.reserve( zpos.size() * 11 ); patchKmer_.reserve( zpos.size() * 11 );
    uint64_t next = 0, curBeg = ~0ull;
    for( size_t k = 0; k < zpos.size(); k++ ){
        if( zbeg[k] != curBeg ){ curBeg = zbeg[k]; next = 0; }
        const uint8_t* w = win.data() + k * 21;                 // w[10 + d] = code at z + d
        const uint64_t z = zpos[k] - curBeg, L = zend[k] - curBeg;
        const uint64_t last = std::min<uint64_t>( L - 1, z + 10 );
        // zeros among w[0..c]: a span with this base as its only undefined one needs one draw, and the rest of its hash rolls
        // from position to position (the usual case: the structural N of a record without other undefined bases)
        int zcum[22]; zcum[0] = 0;
        for( int c = 0; c < 21; c++ ) zcum[c + 1] = zcum[c] + ( w[c] == 0 ? 1 : 0 );
        bool rolling = false;
        size_t hdet = 0;                                        // hash of the span with the undefined base counted as digit 0
        for( uint64_t i = std::max( next, z ); i <= last; i++ ){
            const size_t span = i < 10 ? static_cast<size_t>( i ) + 1 : 11;
            const int cNew = 10 + static_cast<int>( i - z ), cOld = cNew - static_cast<int>( span ) + 1;   // window cells of the span
            size_t h = 0;
            if( zcum[cNew + 1] - zcum[cOld] == 1 ){
                if( rolling && i >= 11 ){                       // the previous span was full: its oldest base leaves
                    hdet = ( hdet - ( w[cOld - 1] ? static_cast<size_t>( w[cOld - 1] - 1 ) : 0 ) * Y_[10] ) * A + static_cast<size_t>( w[cNew] - 1 );
                } else if( rolling ){                           // the span still grows at the start of the record
                    hdet = hdet * A + static_cast<size_t>( w[cNew] - 1 );
                } else {
                    hdet = 0;
                    for( size_t kk = span; kk > 0; kk-- ){
                        const uint8_t code = w[cNew - static_cast<int>( kk ) + 1];
                        if( code ) hdet += static_cast<size_t>( code - 1 ) * Y_[kk - 1];
                    }
                    rolling = true;
                }
                h = hdet + ( static_cast<size_t>( util::rand31() ) % A ) * Y_[i - z];
            } else {
                rolling = false;
                for( size_t kk = span; kk > 0; kk-- ){
                    const uint8_t code = w[cNew - static_cast<int>( kk ) + 1];
                    const size_t digit = ( code == 0 ) ? static_cast<size_t>( util::rand31() ) % A : static_cast<size_t>( code - 1 );
                    h += digit * Y_[kk - 1];
                }
            }
            patchPos_.push_back( curBeg + i );
            patchKmer_.push_back( h );
        }
        next = std::max( next, last + 1 );
    }
    mark( "rand() draws" );
    BAMM_CHECK( bamm_seqset_finish_patches( device_, patchPos_.data(), patchKmer_.data(), patchPos_.size() ) );
    mark( "bamm_seqset_finish_patches" );
    codesOnDevice_ = true;
    finalize();
    mark( "finalize" );
    return true;
}

// Positions whose 11-mer hash contains a code-0 base: the reference replaces the 0 by util::rand31() % A separately for every
// (position, k) pair while it builds kmer_ (src/init/Sequence.cpp:35-41). The draws are made here in the same order
// (records in file order, i ascending, bases of the k-mer from oldest to newest), so the libc rand() stream — and with
// it every later consumer of util::rand31() — stays aligned with the reference.
void SequenceSet::drawPatches( uint64_t begin, uint64_t end ){
    const uint8_t* c = codes_.data() + begin;
    const size_t L = static_cast<size_t>( end - begin );
    const size_t A = Y_[1];
    size_t next = 0;                                        // first position not yet examined
    for( size_t z = 0; z < L; z++ ){
        if( c[z] != 0 ) continue;
        const size_t last = std::min( L - 1, z + 10 );
        for( size_t i = std::max( next, z ); i <= last; i++ ){
            const size_t span = i < 10 ? i + 1 : 11;        // bases i-span+1 .. i
            size_t h = 0;
            for( size_t k = span; k > 0; k-- ){
                const uint8_t code = c[i - k + 1];
                const size_t digit = ( code == 0 ) ? static_cast<size_t>( util::rand31() ) % A : static_cast<size_t>( code - 1 );
                h += digit * Y_[k - 1];
            }
            patchPos_.push_back( begin + i );
            patchKmer_.push_back( h );
        }
        next = std::max( next, last + 1 );
    }
}

void SequenceSet::finalize(){
    const size_t N = offsets_.size() - 1;
    handles_.reserve( N );
    for( size_t n = 0; n < N; n++ ) handles_.emplace_back( this, n );
}

std::vector<Sequence*> SequenceSet::getSequences(){
    std::vector<Sequence*> out( handles_.size() );
    for( size_t n = 0; n < handles_.size(); n++ ) out[n] = &handles_[n];
    return out;
}

size_t SequenceSet::kmerAt( size_t n, size_t i ){
    ensureCodes();
    const uint64_t g = offsets_[n] + i;
    auto it = std::lower_bound( patchPos_.begin(), patchPos_.end(), g );
    if( it != patchPos_.end() && *it == g ) return static_cast<size_t>( patchKmer_[it - patchPos_.begin()] );
    const uint8_t* c = codes_.data() + offsets_[n];
    const size_t span = i < 10 ? i + 1 : 11;
    size_t h = 0;
    for( size_t t = 0; t < span; t++ ) h += static_cast<size_t>( c[i - t] - 1 ) * Y_[t];
    return h;
}

size_t* SequenceSet::kmersOf( size_t n ){
    std::call_once( kmersOnce_, [this](){
        ensureCodes();
        kmers_.assign( codes_.size(), 0 );
        for( size_t s = 0; s + 1 < offsets_.size(); s++ ){
            const uint8_t* c = codes_.data() + offsets_[s];
            size_t* km = kmers_.data() + offsets_[s];
            const size_t L = static_cast<size_t>( offsets_[s + 1] - offsets_[s] );
            for( size_t i = 0; i < L; i++ ){
                const size_t span = i < 10 ? i + 1 : 11;
                size_t h = 0;
                for( size_t t = 0; t < span; t++ ) h += static_cast<size_t>( c[i - t] - 1 ) * Y_[t];
                km[i] = h;
            }
        }
        for( size_t p = 0; p < patchPos_.size(); p++ ) kmers_[patchPos_[p]] = static_cast<size_t>( patchKmer_[p] );
    } );
    return kmers_.data() + offsets_[n];
}

bamm_seqset* SequenceSet::device(){
    std::lock_guard<std::mutex> guard( deviceMutex_ );
    if( !device_ ){
        BAMM_CHECK( bamm_seqset_create( codes_.data(), offsets_.data(), offsets_.size() - 1, static_cast<int>( Alphabet::getSize() ),
                                        patchPos_.data(), patchKmer_.data(), patchPos_.size(), &device_ ) );
    }
    return device_;
}

SequenceSet* SequenceSet::commonSet( const std::vector<Sequence*>& seqs, std::vector<uint64_t>& indices, bool* isWholeSet ){
    indices.clear();
    if( seqs.empty() ){
        std::cerr << "Error: empty sequence list." << std::endl;
        std::exit( 1 );
    }
    SequenceSet* set = seqs[0]->getSet();
    indices.reserve( seqs.size() );
    bool identity = seqs.size() == set->size();
    for( size_t n = 0; n < seqs.size(); n++ ){
        if( seqs[n]->getSet() != set ){
            std::cerr << "Error: sequences of one EM / scoring call must come from one SequenceSet." << std::endl;
            std::exit( 1 );
        }
        indices.push_back( seqs[n]->getIndex() );
        identity = identity && seqs[n]->getIndex() == n;
    }
    if( isWholeSet ) *isWholeSet = identity;
    return set;
}

void Sequence::print(){
    std::cout << ">" << getHeader() << std::endl;
    const uint8_t* c = getSequence();
    for( size_t i = 0; i < getL(); i++ ) std::cout << Alphabet::getBase( c[i] );
    std::cout << std::endl;
}

void SequenceSet::print(){}
