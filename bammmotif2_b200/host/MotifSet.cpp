#include "MotifSet.h"

#include <fstream>
#include <iostream>
#include <sstream>

MotifSet::MotifSet( char* indir, size_t l_flank, size_t r_flank, std::string tag, SequenceSet* posSet,
                    float** v_bg, size_t k_bg, size_t K, std::vector<float> alphas, size_t maxPWM, float glob_q ){

    std::ifstream file( indir, std::ifstream::in );

    if( tag == "bindingsites" ){
        if( !file.good() ){
            std::cout << "Error: Cannot open binding sites file: " << indir << std::endl;
            exit( 1 );
        }
        std::string first;
        std::getline( file, first );                            // the first site defines the core width
        Motif* motif = new Motif( first.length() + l_flank + r_flank, K, alphas, v_bg, k_bg, glob_q );
        motif->initFromBindingSites( indir, l_flank, r_flank );
        motifs_.push_back( motif );
        N_ = 1;
        maxW_ = motif->getW();

    } else if( tag == "PWM" ){
        if( !file.good() ){
            std::cout << "Error: Cannot open PWM file: " << indir << std::endl;
            exit( 1 );
        }
        // MEME v4 text: every "letter-probability matrix: alength= A w= W [nsites= ..] [E= ..] [occur= q]" line is
        // followed by W rows of A numbers
        std::string line, row;
        while( std::getline( file, line ) ){
            if( line.find( "letter-probability matrix" ) == std::string::npos ) continue;
            size_t asize = 0, length = 0;
            float q = glob_q;
            { std::stringstream in( line.substr( line.find( "h=" ) + 2 ) ); in >> asize; }
            { std::stringstream in( line.substr( line.find( "w=" ) + 2 ) ); in >> length; }
            length += l_flank + r_flank;
            if( line.find( "occur=" ) != std::string::npos ){
                std::stringstream in( line.substr( line.find( "occur=" ) + 7 ) );
                in >> q;
            }
            Motif* motif = new Motif( length, K, alphas, v_bg, k_bg, q );
            std::vector<std::vector<float>> columns( asize, std::vector<float>( length, 1.0f / ( float )asize ) );   // flanks stay uniform
            for( size_t j = l_flank; j < length - r_flank; j++ ){
                if( !std::getline( file, row ) ){
                    std::cerr << "Error: Cannot find any PWM in the MEME-format file: " << indir
                              << "\nPlease check the content of your input MEME file." << std::endl;
                    exit( 1 );
                }
                std::stringstream number( row );
                for( size_t y = 0; y < asize; y++ ) number >> columns[y][j];
            }
            std::vector<float*> PWM( asize );
            for( size_t y = 0; y < asize; y++ ) PWM[y] = columns[y].data();
            motif->initFromPWM( PWM.data(), asize, posSet, q );
            N_++;
            motifs_.push_back( motif );
            maxW_ = ( motif->getW() > maxW_ ) ? motif->getW() : maxW_;
            if( N_ >= maxPWM ) break;
        }
        if( N_ == 0 ){
            std::cerr << "Error: Cannot find any PWM in the MEME-format file: " << indir
                      << "\nPlease check the version of your input MEME file." << std::endl;
            exit( 1 );
        }

    } else if( tag == "BaMM" ){
        if( !file.good() ){
            std::cerr << "Error: Cannot open BaMM file: " << indir << std::endl;
            exit( 1 );
        }
        // width = number of blank lines, order+1 = number of lines of the first position; every later position must
        // have the same number of lines
        size_t model_length = 0, model_order = 0, check_lines = 0;
        std::string line;
        while( std::getline( file, line ) ){
            if( line.empty() ){
                model_length++;
                if( model_length > 1 && check_lines != model_order ){
                    std::cerr << "This is not a BaMM-format file: " << indir << std::endl;
                    exit( 1 );
                }
                check_lines = 0;
            } else if( model_length == 0 ){
                model_order++;
            } else {
                check_lines++;
            }
        }
        model_order -= 1;
        if( model_order > 8 ){
            std::cerr << "The input BaMM model order is too high: " << indir << std::endl;
            exit( 1 );
        }
        Motif* motif = new Motif( model_length + l_flank + r_flank, K, alphas, v_bg, k_bg, glob_q );
        motif->initFromBaMM( indir, l_flank, r_flank );
        motifs_.push_back( motif );
        N_ = 1;
        maxW_ = motif->getW();
    }
}

MotifSet::~MotifSet(){
    for( Motif* m : motifs_ ) delete m;
}

void MotifSet::print(){
    for( size_t i = 0; i < N_; i++ ){
        fprintf( stderr, "INITIALIZED PROBABILITIES for Motif %d\n", ( int )i + 1 );
        motifs_[i]->print();
    }
}

void MotifSet::write( char* ){}
