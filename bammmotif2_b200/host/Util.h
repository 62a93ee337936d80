// Util.h — the few helpers of the reference's utils.h that the hot path's host side needs
// (reference: src/refinement/utils.h:66-85 baseName, :154-165 createDirectory, :167-179 ipow).
#ifndef BAMM_HOST_UTIL_H_
#define BAMM_HOST_UTIL_H_

#include <cstddef>
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

}  // namespace util

#endif
