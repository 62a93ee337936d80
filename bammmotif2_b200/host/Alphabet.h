// Alphabet — static base <-> code tables of the host side.
// Mirrors the public interface of the reference's Alphabet (src/init/Alphabet.h:13-31); the semantics that matter
// for bit-exact encoding are in src/init/Alphabet.cpp:10-55: N -> 0, letters -> 1..A case-insensitively, every
// other character -> 0, codes above A print as 'N', and the complement of code 0 (and of any code above A) is the
// *character* 'N' (78), not 0.
#ifndef BAMM_HOST_ALPHABET_H_
#define BAMM_HOST_ALPHABET_H_

#include <cstddef>
#include <cstdint>
#include <string>

class Alphabet {
public:
    static void     init( const char* alphabetType );   // STANDARD | METHYLC | HYDROXYMETHYLC | EXTENDED, else exit(1)
    static void     destruct();
    static size_t   getSize()                           { return size_; }
    static const char* getAlphabet()                    { return letters_.c_str(); }
    static uint8_t  getCode( char base )                { return base2code_[static_cast<unsigned char>( base )]; }
    static char     getBase( uint8_t code )             { return code2base_[code]; }
    static uint8_t  getComplementCode( uint8_t code )   { return code2comp_[code]; }

private:
    static size_t       size_;
    static std::string  letters_;
    static std::string  complement_;
    static uint8_t      base2code_[256];
    static char         code2base_[256];
    static uint8_t      code2comp_[256];
};

#endif
