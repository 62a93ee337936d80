// Util.h — the few helpers of the reference's utils.h that the hot path's host side needs
// (reference: src/refinement/utils.h:66-85 baseName, :154-165 createDirectory, :167-179 ipow).
#ifndef BAMM_HOST_UTIL_H_
#define BAMM_HOST_UTIL_H_

#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <string>

#include <sys/stat.h>

namespace util {

inline size_t ipow( size_t base, size_t exp ){
    size_t r = 1;
    for( ; exp; exp >>= 1, base *= base ) if( exp & 1 ) r *= base;
    return r;
}

// File name without directories and without its LAST extension ("dir/JunD.fasta" -> "JunD"). Quirks kept from the
// reference: the dot is searched in the whole path from index 1, and a bare name without a dot keeps one character.
inline std::string baseName( const char* path ){
    const std::string p( path );
    size_t end = 0;
    for( size_t i = 1; i < p.size(); i++ ) if( p[i] == '.' ) end = i - 1;
    size_t start = 0;
    const size_t slash = p.rfind( '/' );
    if( slash != std::string::npos && slash != 0 ) start = slash + 1;
    return p.substr( start, end - start + 1 );          // size_t wrap-around when there is no dot after the last '/': rest of the name
}

inline void createDirectory( const char* dir ){
    struct stat st;
    if( stat( dir, &st ) != 0 ){
        std::cout << "New output directory is created automatically.\n";
        if( std::system( ( "mkdir -p " + std::string( dir ) ).c_str() ) != 0 ){
            std::cerr << "Error: Directory " << dir << " could not be created." << std::endl;
            std::exit( -1 );
        }
    }
}

// The libc random stream the reference consumes everywhere (srand(42), rand(): k-mer hashes around undefined bases, site
// padding, negative sampling, tie-breaks of the PR walk), as a private generator: glibc's TYPE_3 additive feedback
// generator o[n] = o[n-31] + o[n-3] (mod 2^32), seeded by the Lehmer sequence 16807 x mod (2^31 - 1), first 310 values
// dropped, result o[n] >> 1 — bit for bit the values of glibc's srand() / rand() (tests/test_host_cpp.py compares them),
// without the lock and the call into libc: 11 draws per both-strand record add up to 1e7 draws on a 1M-sequence set.
// One stream per process, used from the main thread only, as in the reference.
struct LibcRand {
    uint32_t o[34];
    int at;                                             // ring index of the next value
    LibcRand(){ seed( 1 ); }                            // glibc's state before the first srand()
    void seed( unsigned s ){
        int32_t r[34];
        r[0] = s ? static_cast<int32_t>( s ) : 1;
        for( int i = 1; i < 31; i++ ){
            const long hi = r[i - 1] / 127773, lo = r[i - 1] % 127773;
            long w = 16807 * lo - 2836 * hi;
            if( w < 0 ) w += 2147483647;
            r[i] = static_cast<int32_t>( w );
        }
        for( int i = 31; i < 34; i++ ) r[i] = r[i - 31];
        for( int i = 0; i < 34; i++ ) o[i] = static_cast<uint32_t>( r[i] );
        at = 0;                                         // n = 34 sits in slot 0
        for( int i = 0; i < 310; i++ ) step();
    }
    uint32_t step(){
        // slot of n is n mod 34: n-31 is 3 slots ahead, n-3 is 31 slots ahead
        const int a = at + 3 >= 34 ? at + 3 - 34 : at + 3, b = at + 31 >= 34 ? at + 31 - 34 : at + 31;
        const uint32_t v = o[a] + o[b];
        o[at] = v;
        at = at + 1 == 34 ? 0 : at + 1;
        return v;
    }
    int next(){ return static_cast<int>( step() >> 1 ); }
};
inline LibcRand& libcRand(){ static LibcRand g; return g; }
inline void srand42( unsigned s ){ libcRand().seed( s ); }
inline int rand31(){ return libcRand().next(); }

}  // namespace util

#endif
