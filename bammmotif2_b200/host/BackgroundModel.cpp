#include "BackgroundModel.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <iomanip>
#include <iostream>

#include <sys/stat.h>

#include "Util.h"

void BackgroundModel::allocate(){
    Y_.clear();
    for( size_t k = 0; k < K_ + 8; k++ ) Y_.push_back( util::ipow( Alphabet::getSize(), k ) );
    off_.assign( K_ + 2, 0 );
    for( size_t k = 0; k <= K_; k++ ) off_[k + 1] = off_[k] + Y_[k + 1];
    v_.assign( off_[K_ + 1], 0.0f );
    vRows_.resize( K_ + 1 );
    for( size_t k = 0; k <= K_; k++ ) vRows_[k] = v_.data() + off_[k];
}

BackgroundModel::BackgroundModel( std::vector<Sequence*> seqs, size_t order, std::vector<float> alpha,
                                  bool interpolate, std::string basename )
    : basename_( basename ), K_( order ), A_( alpha ), interpolate_( interpolate ){
    allocate();
    n_.assign( off_[K_ + 1], 0 );
    // every stored position contributes kmer[i] % A^(k+1) to every order k <= K (reference :26-42)
    std::vector<uint64_t> indices;
    bool whole = false;
    SequenceSet* set = SequenceSet::commonSet( seqs, indices, &whole );
    if( whole ){
        BAMM_CHECK( bamm_seqset_count_kmers( set->device(), static_cast<int>( K_ ), n_.data() ) );
    } else {
        // a subset: count on a temporary device copy of just these records
        std::vector<uint8_t> codes;
        std::vector<uint64_t> offsets( 1, 0 ), ppos, pkmer;
        for( uint64_t n : indices ){
            const uint64_t b = set->offsets()[n], e = set->offsets()[n + 1];
            const uint64_t shift = codes.size() - b;
            codes.insert( codes.end(), set->codes().begin() + b, set->codes().begin() + e );
            offsets.push_back( codes.size() );
            auto lo = std::lower_bound( set->patchPositions().begin(), set->patchPositions().end(), b );
            for( ; lo != set->patchPositions().end() && *lo < e; ++lo ){
                ppos.push_back( *lo + shift );
                pkmer.push_back( set->patchKmers()[lo - set->patchPositions().begin()] );
            }
        }
        bamm_seqset* tmp = nullptr;
        BAMM_CHECK( bamm_seqset_create( codes.data(), offsets.data(), offsets.size() - 1, static_cast<int>( Alphabet::getSize() ),
                                        ppos.data(), pkmer.data(), ppos.size(), &tmp ) );
        BAMM_CHECK( bamm_seqset_count_kmers( tmp, static_cast<int>( K_ ), n_.data() ) );
        bamm_seqset_destroy( tmp );
    }
    hasCounts_ = true;
    calculateV();
}

BackgroundModel::BackgroundModel( const std::vector<uint64_t>& counts, size_t order, std::vector<float> alpha,
                                  bool interpolate, std::string basename )
    : basename_( basename ), K_( order ), A_( alpha ), interpolate_( interpolate ){
    allocate();
    if( counts.size() != off_[K_ + 1] ){
        std::cerr << "Error: background count table has the wrong size." << std::endl;
        std::exit( 1 );
    }
    n_ = counts;
    hasCounts_ = true;
    calculateV();
}

// reads the .hbcp format written by write(): "# K = k", "# A = a0 a1 ...", then one line of A^(k+1) values per order
// (reference: src/init/BackgroundModel.cpp:48-129)
BackgroundModel::BackgroundModel( std::string filePath ){
    basename_ = util::baseName( filePath.c_str() );
    std::ifstream probe( filePath );
    if( !probe.good() ){
        std::cerr << "Error: Input Background Model file does not exist." << std::endl;
        std::exit( 1 );
    }
    probe.close();
    struct stat sb;
    if( !( stat( filePath.c_str(), &sb ) == 0 && S_ISREG( sb.st_mode ) ) ) return;   // the reference silently leaves the model empty
    FILE* file = std::fopen( filePath.c_str(), "r" );
    if( !file ){
        std::cerr << "Error: Cannot open BaMM file: " << filePath << std::endl;
        std::exit( 1 );
    }
    auto bad = [&](){
        std::cerr << "Error: Wrong BaMM format: " << filePath << std::endl;
        std::exit( 1 );
    };
    int K;
    if( std::fscanf( file, "# K = %d\n", &K ) != 1 ) bad();
    K_ = static_cast<size_t>( K );
    A_.resize( K_ + 1 );
    float a;
    if( std::fscanf( file, "# A = %e", &a ) != 1 ) bad();
    A_[0] = a;
    for( size_t k = 1; k <= K_; k++ ){
        if( std::fscanf( file, "%e", &a ) != 1 ) bad();
        A_[k] = a;
    }
    allocate();
    for( size_t i = 0; i < v_.size(); i++ ){
        float value;
        if( std::fscanf( file, "%e", &value ) == EOF ) bad();
        v_[i] = value;
    }
    std::fclose( file );
}

BackgroundModel::~BackgroundModel(){}

void BackgroundModel::expV(){
    for( float& x : v_ ) x = expf( x );
    vIsLog_ = false;
}

void BackgroundModel::logV(){
    for( float& x : v_ ) x = logf( x );
    vIsLog_ = true;
}

// reference: BackgroundModel::calculateV, src/init/BackgroundModel.cpp:441-472 (same operation order, fp32)
void BackgroundModel::calculateV(){
    const size_t A = Y_[1];
    size_t baseCounts = 0;
    for( size_t y = 0; y < A; y++ ) baseCounts += n_[y];
    for( size_t y = 0; y < A; y++ ){
        v_[y] = ( static_cast<float>( n_[y] ) + A_[0] * 0.25f ) / ( static_cast<float>( baseCounts ) + A_[0] );
    }
    for( size_t k = 1; k <= K_; k++ ){
        const uint64_t* nk = n_.data() + off_[k];
        const uint64_t* nk1 = n_.data() + off_[k - 1];
        float* vk = v_.data() + off_[k];
        const float* vk1 = v_.data() + off_[k - 1];
        for( size_t y = 0; y < Y_[k + 1]; y++ ){
            const size_t y2 = y % Y_[k];        // without the oldest base
            const size_t yk = y / A;            // without the newest base
            const float prior = interpolate_ ? vk1[y2] : 0.25f;
            vk[y] = ( static_cast<float>( nk[y] ) + A_[k] * prior ) / ( static_cast<float>( nk1[yk] ) + A_[k] );
        }
    }
}

void BackgroundModel::print(){
    std::cout << ( interpolate_ ? "Homogeneous Bayesian Markov Model" : "Homogeneous Markov Model" ) << std::endl;
    std::cout << "name = " << basename_ << std::endl << std::endl << "K = " << K_ << std::endl << "A =";
    for( size_t k = 0; k <= K_; k++ ) std::cout << " " << A_[k];
    std::cout << std::endl;
    for( size_t k = 0; k <= K_; k++ ){
        for( size_t y = 0; y < Y_[k + 1]; y++ ){
            std::cout << ( y ? " " : "" ) << std::fixed << std::setprecision( 3 ) << v_[off_[k] + y];
        }
        std::cout << std::endl;
    }
}

// .hbcp: conditional probabilities, .hbp: joint probabilities p(y) = v(y) * p(y without its newest base)
// (reference: src/init/BackgroundModel.cpp:353-439; number formats are part of the file contract)
void BackgroundModel::write( char* odir, std::string basename ){
    if( vIsLog_ ) expV();
    const std::string stem = std::string( odir ) + '/' + basename;
    auto fail = [&](){
        std::cerr << "Error: Cannot write into output directory: " << odir << std::endl;
        std::exit( 1 );
    };
    {
        std::ofstream file( stem + ( interpolate_ ? ".hbcp" : ".hnbcp" ) );
        if( !file.is_open() ) fail();
        file << "# K = " << K_ << std::endl << "# A =";
        for( size_t k = 0; k <= K_; k++ ) file << " " << A_[k];
        file << std::endl;
        for( size_t k = 0; k <= K_; k++ ){
            for( size_t y = 0; y < Y_[k + 1]; y++ ) file << std::scientific << std::setprecision( 6 ) << v_[off_[k] + y] << " ";
            file << std::endl;
        }
    }
    std::vector<float> p( v_.size() );
    for( size_t y = 0; y < Y_[1]; y++ ) p[y] = v_[y];
    for( size_t k = 1; k <= K_; k++ ){
        for( size_t y = 0; y < Y_[k + 1]; y++ ) p[off_[k] + y] = v_[off_[k] + y] * p[off_[k - 1] + y / Y_[1]];
    }
    {
        std::ofstream file( stem + ( interpolate_ ? ".hbp" : ".hnbp" ) );
        if( !file.is_open() ) fail();
        file << "# K = " << K_ << std::endl << "# A =";
        for( size_t k = 0; k <= K_; k++ ) file << std::fixed << std::setprecision( 2 ) << " " << A_[k];
        file << std::endl;
        for( size_t k = 0; k <= K_; k++ ){
            for( size_t y = 0; y < Y_[k + 1]; y++ ) file << std::scientific << std::setprecision( 6 ) << p[off_[k] + y] << " ";
            file << std::endl;
        }
    }
}
