#include "Alphabet.h"

#include <cctype>
#include <cstdlib>
#include <cstring>
#include <iostream>

size_t      Alphabet::size_ = 0;
std::string Alphabet::letters_;
std::string Alphabet::complement_;
uint8_t     Alphabet::base2code_[256];
char        Alphabet::code2base_[256];
uint8_t     Alphabet::code2comp_[256];

namespace {
struct Kind { const char* name; const char* letters; const char* complement; };
// reference: src/init/Alphabet.cpp:12-27 (M and H pair with G)
const Kind kKinds[] = {
    { "STANDARD",       "ACGT",   "TGCA"   },
    { "METHYLC",        "ACGTM",  "TGCAG"  },
    { "HYDROXYMETHYLC", "ACGTH",  "TGCAG"  },
    { "EXTENDED",       "ACGTMH", "TGCAGG" },
};
}

void Alphabet::init( const char* alphabetType ){
    const Kind* kind = nullptr;
    for( const Kind& k : kKinds ) if( std::strcmp( alphabetType, k.name ) == 0 ) kind = &k;
    if( !kind ){
        std::cerr << "Error: Correct alphabet type to STANDARD, METHYLC, HYDROXYMETHYLC, or EXTENDED" << std::endl;
        std::exit( 1 );
    }
    letters_ = kind->letters;
    complement_ = kind->complement;
    size_ = letters_.size();
    std::memset( base2code_, 0, sizeof( base2code_ ) );
    for( size_t c = 0; c < 256; c++ ){ code2base_[c] = 'N'; code2comp_[c] = static_cast<uint8_t>( 'N' ); }
    for( size_t i = 0; i < size_; i++ ){
        const uint8_t code = static_cast<uint8_t>( i + 1 );
        base2code_[static_cast<unsigned char>( letters_[i] )] = code;
        base2code_[static_cast<unsigned char>( std::tolower( letters_[i] ) )] = code;
        code2base_[code] = letters_[i];
    }
    for( size_t i = 0; i < size_; i++ ){
        code2comp_[i + 1] = base2code_[static_cast<unsigned char>( complement_[i] )];
    }
}

void Alphabet::destruct(){
    size_ = 0;
    letters_.clear();
    complement_.clear();
}
