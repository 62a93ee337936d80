// host_check — TEST HELPER for the CPU-only parts of the host classes (no CUDA device needed): dumps intermediates as
// raw little-endian arrays that tests/test_host_cpp.py compares with the golden vectors made from the reference.
//   host_check encode  ALPHABET FASTA SS OUTDIR          -> codes.u8 offsets.u64 kmer.u64 basefreq.f32
//   host_check init    ALPHABET K KBG COUNTS.u64 ALPHA_BG.f32 SITES Q OUTDIR   -> bg_v.f32 v_init.f32 p_init.f32 (+ .hbcp/.hbp/.ihbcp/.ihbp files)
//   host_check neg     ALPHABET FASTA SS MFOLD SORDER OUTDIR                   -> neg_codes.u8 neg_offsets.u64
//   host_check pr      POSN NEGN Q POS.f32 NEG.f32 OUTDIR                      -> TP FP FDR Rec PNpval occ_frac (.f32)
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "FDR.h"
#include "Util.h"
#include "MotifSet.h"
#include "SeqGenerator.h"

template <typename T> static void dump( const std::string& path, const T* p, size_t n ){
    std::ofstream f( path, std::ios::binary );
    f.write( reinterpret_cast<const char*>( p ), n * sizeof( T ) );
}
template <typename T> static std::vector<T> slurp( const std::string& path ){
    std::ifstream f( path, std::ios::binary | std::ios::ate );
    if( !f ){ std::cerr << "cannot read " << path << std::endl; exit( 2 ); }
    const size_t bytes = f.tellg();
    std::vector<T> v( bytes / sizeof( T ) );
    f.seekg( 0 );
    f.read( reinterpret_cast<char*>( v.data() ), bytes );
    return v;
}

int main( int argc, char** argv ){
    if( argc < 3 ) return 2;
    const std::string mode = argv[1];
    util::srand42( 42 );
    if( mode == "rand" ){                                    // rand SEED N: N draws of the private libc-compatible stream
        util::srand42( static_cast<unsigned>( strtoul( argv[2], nullptr, 10 ) ) );
        const long n = argc > 3 ? atol( argv[3] ) : 10;
        for( long i = 0; i < n; i++ ) std::cout << util::rand31() << "\n";
        return 0;
    }
    if( mode == "parse" ){                                   // timing of the FASTA reader alone
        Alphabet::init( argv[2] );
        const auto t0 = std::chrono::steady_clock::now();
        SequenceSet set( argv[3], atoi( argv[4] ) != 0 );
        const auto t1 = std::chrono::steady_clock::now();
        double resident = -1.0;
        if( argc > 5 ){                                      // ... and until the set is resident on the device
            set.device();
            resident = std::chrono::duration<double>( std::chrono::steady_clock::now() - t0 ).count();
        }
        std::cout << set.size() << " records, " << set.offsets().back() << " stored codes, reader "
                  << std::chrono::duration<double>( t1 - t0 ).count() << " s";
        if( resident >= 0.0 ) std::cout << ", resident on the device after " << resident << " s";
        std::cout << std::endl;
        return 0;
    }
    if( mode == "encode" ){
        Alphabet::init( argv[2] );
        SequenceSet set( argv[3], atoi( argv[4] ) != 0 );
        const std::string out = argv[5];
        dump( out + "/codes.u8", set.codes().data(), set.codes().size() );
        dump( out + "/offsets.u64", set.offsets().data(), set.offsets().size() );
        std::vector<uint64_t> kmer;
        for( Sequence* s : set.getSequences() ){
            size_t* k = s->getKmer();
            for( size_t i = 0; i < s->getL(); i++ ){
                if( k[i] != s->kmerAt( i ) ){ std::cerr << "kmerAt mismatch" << std::endl; return 3; }
                kmer.push_back( k[i] );
            }
        }
        dump( out + "/kmer.u64", kmer.data(), kmer.size() );
        dump( out + "/basefreq.f32", set.getBaseFrequencies(), Alphabet::getSize() );
        return 0;
    }
    if( mode == "init" ){
        Alphabet::init( argv[2] );
        const size_t K = atoi( argv[3] ), Kbg = atoi( argv[4] );
        std::vector<uint64_t> counts = slurp<uint64_t>( argv[5] );
        std::vector<float> alphaBg = slurp<float>( argv[6] );
        char* sites = argv[7];
        const float q = atof( argv[8] );
        const std::string out = argv[9];
        BackgroundModel bg( counts, Kbg, alphaBg, true, "check" );
        dump( out + "/bg_v.f32", bg.flatV().data(), bg.flatV().size() );
        std::vector<float> alpha( K + 1, 1.f );
        for( size_t k = 1; k <= K; k++ ) alpha[k] = 7.0f * powf( 3.0f, ( float )k );
        MotifSet ms( sites, 0, 0, "bindingsites", NULL, bg.getV(), Kbg, K, alpha, 10, q );
        Motif* m = ms.getMotifs()[0];
        dump( out + "/v_init.f32", m->flatV().data(), m->flatV().size() );
        dump( out + "/p_init.f32", m->flatP().data(), m->flatP().size() );
        std::vector<char> dir( out.begin(), out.end() ); dir.push_back( 0 );
        bg.write( dir.data(), "check" );
        m->write( dir.data(), "check_motif_1" );
        // pointer views must alias the flat tables
        if( m->getV()[K][1] != m->flatV().data() + m->offsetOfOrder( K ) + m->getW() ) return 3;
        BackgroundModel reread( out + "/check.hbcp" );
        if( reread.getOrder() != Kbg ) return 4;
        return 0;
    }
    if( mode == "pwminit" ){
        // host_check pwminit ALPHABET FASTA SS K KBG COUNTS.u64 ALPHA_BG.f32 MEME MAXPWM Q OUTDIR -> v_init_<m>.f32
        Alphabet::init( argv[2] );
        SequenceSet pos( argv[3], atoi( argv[4] ) != 0 );
        const size_t K = atoi( argv[5] ), Kbg = atoi( argv[6] );
        BackgroundModel bg( slurp<uint64_t>( argv[7] ), Kbg, slurp<float>( argv[8] ), true, "check" );
        std::vector<float> alpha( K + 1, 1.f );
        for( size_t k = 1; k <= K; k++ ) alpha[k] = 7.0f * powf( 3.0f, ( float )k );
        MotifSet ms( argv[9], 0, 0, "PWM", &pos, bg.getV(), Kbg, K, alpha, atoi( argv[10] ), atof( argv[11] ) );
        const std::string out = argv[12];
        for( size_t m = 0; m < ms.getN(); m++ ){
            Motif* mo = ms.getMotifs()[m];
            dump( out + "/v_init_" + std::to_string( m + 1 ) + ".f32", mo->flatV().data(), mo->flatV().size() );
        }
        return 0;
    }
    if( mode == "neg" ){
        Alphabet::init( argv[2] );
        SequenceSet pos( argv[3], atoi( argv[4] ) != 0 );
        SequenceSet neg0( argv[3], atoi( argv[4] ) != 0 );        // the driver reads the file twice (Global.cpp:105-106)
        SeqGenerator gen( pos.getSequences(), NULL, atoi( argv[6] ), 1.0f, false );
        std::unique_ptr<SequenceSet> neg = gen.sample_bgseqset_by_fold( atoi( argv[5] ) );
        const std::string out = argv[7];
        dump( out + "/neg_codes.u8", neg->codes().data(), neg->codes().size() );
        dump( out + "/neg_offsets.u64", neg->offsets().data(), neg->offsets().size() );
        return 0;
    }
    if( mode == "pr" ){
        Alphabet::init( "STANDARD" );
        const size_t posN = atol( argv[2] ), negN = atol( argv[3] );
        const float q = atof( argv[4] );
        std::vector<float> alpha( 1, 1.f );
        Motif dummy( 4, 0, alpha, NULL, 0, q );
        std::vector<Sequence*> pos( posN, nullptr ), neg( negN, nullptr );
        FDR fdr( pos, neg, &dummy, NULL, 4, false, true, true, false, false );
        fdr.setScores( slurp<float>( argv[5] ), slurp<float>( argv[6] ) );
        fdr.calculatePR();
        fdr.calculatePvalues();                     // FDR.cpp:278-330 (needs the scores sorted or not: it sorts them itself)
        const std::string out = argv[7];
        dump( out + "/zoops_pvalue.f32", fdr.zoopsPvalues().data(), fdr.zoopsPvalues().size() );
        dump( out + "/TP.f32", fdr.zoopsTP().data(), fdr.zoopsTP().size() );
        dump( out + "/FP.f32", fdr.zoopsFP().data(), fdr.zoopsFP().size() );
        dump( out + "/FDR.f32", fdr.zoopsFDR().data(), fdr.zoopsFDR().size() );
        dump( out + "/Rec.f32", fdr.zoopsRecall().data(), fdr.zoopsRecall().size() );
        dump( out + "/PNpval.f32", fdr.pnPvalues().data(), fdr.pnPvalues().size() );
        const float occ = fdr.occFrac();
        dump( out + "/occ_frac.f32", &occ, 1 );
        std::vector<char> dir( out.begin(), out.end() ); dir.push_back( 0 );
        fdr.write( dir.data(), "check_motif_1" );
        return 0;
    }
    return 2;
}
