// BaMMmotif driver on the B200 path: same command line, same outputs as the reference's
// src/refinement/mainBaMM.cpp (flow: background model -> initial motifs -> EM per motif -> optional occurrence scoring
// -> optional cross-validated FDR). The Gibbs optimiser is not available here.
#include <chrono>
#include <iomanip>
#include <iostream>
#include <memory>
#include <thread>
#include <vector>
#include <cstdlib>

#include "EM.h"
#include "Util.h"
#include "FDR.h"
#include "Global.h"
#include "ScoreSeqSet.h"
#include "SeqGenerator.h"

// devices named by BAMM_DEVICES (list "0,2,3" and / or ranges "0-7"), else BAMM_DEVICE, else device 0
static std::vector<int> deviceList(){
    std::vector<int> out;
    if( const char* e = getenv( "BAMM_DEVICES" ) ){
        const char* p = e;
        while( *p ){
            char* end = nullptr;
            long a = strtol( p, &end, 10 );
            if( end == p ) break;
            long b = a;
            if( *end == '-' ){ p = end + 1; b = strtol( p, &end, 10 ); }
            for( long d = a; d <= b && out.size() < 16; d++ ) out.push_back( static_cast<int>( d ) );
            p = ( *end == ',' ) ? end + 1 : end;
            if( *end != ',' && *end != 0 ) break;
        }
    }
    if( out.empty() ){
        const char* dev = getenv( "BAMM_DEVICE" );
        out.push_back( dev ? atoi( dev ) : 0 );
    }
    return out;
}

int main( int nargs, char* args[] ){
    auto t0_wall = std::chrono::high_resolution_clock::now();
    // BAMM_TRACE: wall time of every stage of the driver on stderr (the library prints its own phases with the same switch)
    const bool trace = getenv( "BAMM_TRACE" ) != NULL;
    auto t_mark = t0_wall;
    auto epoch = [](){ return std::chrono::duration<double>( std::chrono::system_clock::now().time_since_epoch() ).count(); };
    if( trace ) std::cerr << "[bamm host] main entered at " << std::fixed << std::setprecision( 3 ) << epoch() << std::defaultfloat << std::setprecision( 6 ) << std::endl;
    auto mark = [&]( const char* stage ){
        if( !trace ) return;
        const auto t = std::chrono::high_resolution_clock::now();
        std::cerr << "[bamm host] " << stage << " " << std::chrono::duration<double, std::milli>( t - t_mark ).count() << " ms" << std::endl;
        t_mark = t;
    };
    std::cout << std::endl
              << "======================================" << std::endl
              << "=      Welcome to use BaMM!motif     =" << std::endl
              << "=          Version 2.0 (B200 path)   =" << std::endl
              << "======================================" << std::endl;

    util::srand42( 42 );                                    // reference: mainBaMM.cpp:22
    // the first CUDA call creates the device context (0.8 s and more on a multi-GPU box): it runs beside option parsing and
    // the FASTA reader; a failure is not reported here — the first real device call of the main thread reports it
    // BAMM_DEVICES="0,1,2,3" or "0-7": one process drives all of them (EM::optimize splits the sequences over the devices and
    // exchanges the counts over NVLink; everything else stays on the first one). BAMM_DEVICE=n: a single device.
    std::vector<int> devices = deviceList();
    std::thread warm( [devices](){          // one context per device, created side by side
        std::vector<std::thread> more;
        for( size_t i = 1; i < devices.size(); i++ ) more.emplace_back( [d = devices[i]](){ bamm_set_device( d ); } );
        bamm_set_device( devices[0] );
        for( std::thread& t : more ) t.join();
    } );
    if( devices[0] != 0 ) bamm_set_device( devices[0] );      // the FASTA reader may already create the device copy of the set
    Global::init( nargs, args );
    warm.join();
    if( Global::CGS ){
        std::cerr << "Error: collapsed Gibbs sampling (--CGS) is not part of the B200 path; use --EM." << std::endl;
        return 1;
    }
    BAMM_CHECK( bamm_set_device( devices[0] ) );
    // EM::mask (--advanceEM) runs on one device
    if( devices.size() > 1 && !Global::advanceEM ) BAMM_CHECK( bamm_set_device_group( devices.data(), static_cast<int>( devices.size() ) ) );
    mark( "options + FASTA" );

    std::vector<Sequence*> posSet = Global::posSequenceSet->getSequences();

    BackgroundModel* bgModel = Global::bgModelGiven
        ? new BackgroundModel( Global::bgModelFilename )
        : new BackgroundModel( posSet, Global::bgModelOrder, Global::bgModelAlpha, Global::interpolateBG, Global::outputFileBasename );
    bgModel->write( Global::outputDirectory, Global::outputFileBasename );      // always written (mainBaMM.cpp:51)
    mark( "background model" );

    MotifSet motif_set( Global::initialModelFilename, Global::addColumns.at( 0 ), Global::addColumns.at( 1 ), Global::initialModelTag,
                        Global::posSequenceSet, bgModel->getV(), Global::bgModelOrder, Global::modelOrder, Global::modelAlpha,
                        Global::maxPWM, Global::q );
    mark( "initial motifs" );

    // sequences shorter than the widest motif cannot hold a window (mainBaMM.cpp:75-83)
    {
        std::vector<Sequence*> kept;
        for( Sequence* s : posSet ) if( s->getL() >= motif_set.getMaxW() ) kept.push_back( s );
        posSet.swap( kept );
    }
    const size_t posN = posSet.size();
    if( posN < Global::cvFold ){
        std::cerr << "There are " << posN << " sequences longer than input motif. Exit!\n";
        exit( 1 );
    }

    // small sets get more negatives per positive (mainBaMM.cpp:100-106; `rest` is a bool in the reference)
    const size_t minSeqN = 5000;
    const bool rest = minSeqN % posSet.size();
    if( posSet.size() < minSeqN ) Global::mFold = minSeqN / posSet.size() + rest;

    // The reference samples the negative set on every run; nothing reads it (or the util::rand31() state it leaves) unless
    // --FDR / --scoreSeqset is given, so it is sampled only then.
    std::unique_ptr<SequenceSet> negSequences;
    std::vector<Sequence*> negSet;
    if( Global::FDR || Global::scoreSeqset ){
        SeqGenerator negseq( posSet, NULL, Global::sOrder, 1.0f, Global::genericNeg );
        negSequences = negseq.sample_bgseqset_by_fold( Global::mFold );
        negSet = negSequences->getSequences();
    }
    mark( "negative set" );

    for( size_t n = 0; n < motif_set.getN(); n++ ){
        Motif* motif = new Motif( *motif_set.getMotifs()[n] );
        const std::string motifName = Global::outputFileBasename + "_motif_" + std::to_string( n + 1 );
        if( Global::saveInitialBaMMs ){
            motif->write( Global::outputDirectory, Global::outputFileBasename + "_init_motif_" + std::to_string( n + 1 ) );
        }
        if( Global::EM ){
            EM model( motif, bgModel, posSet, Global::optimizeQ, Global::verbose, Global::f );
            if( !Global::advanceEM ) model.optimize(); else model.mask();
            if( Global::saveBaMMs ) model.write( Global::outputDirectory, motifName, Global::ss );
            std::cout << "optimized q = " << model.getQ() << std::endl;
            mark( "EM" );
        } else {
            std::cout << "Note: the model is not optimized!\n";
        }
        motif->write( Global::outputDirectory, motifName );                        // always written (mainBaMM.cpp:169-170)

        if( Global::scoreSeqset ){
            BackgroundModel* bg = bgModel;
            if( !Global::EM ){
                if( Global::initialModelTag == "BaMM" ){
                    if( Global::bgModelFilename == NULL ){
                        std::cout << "No background Model file provided for initial search motif!\n";
                        exit( 1 );
                    }
                    bg = new BackgroundModel( Global::bgModelFilename );
                } else if( Global::initialModelTag == "PWM" ){
                    Global::modelOrder = 0;
                }
            }
            ScoreSeqSet scoreNegSet( motif, bgModel, negSet );
            scoreNegSet.calcLogOdds();
            if( Global::saveLogOdds ) scoreNegSet.writeLogOdds( Global::outputDirectory, Global::outputFileBasename + ".negSet", Global::ss );
            std::vector<float> negScores = scoreNegSet.flatMopsScores();            // all window scores of all negatives

            ScoreSeqSet scorePosSet( motif, bg, posSet );
            scorePosSet.calcLogOdds();
            if( Global::saveLogOdds ) scorePosSet.writeLogOdds( Global::outputDirectory, motifName, Global::ss );
            scorePosSet.calcPvalues( scorePosSet.getMopsScores(), negScores );
            scorePosSet.write( Global::outputDirectory, motifName, Global::pvalCutoff, Global::ss );
            if( bg != bgModel ) delete bg;
        }
        delete motif;
    }

    if( Global::FDR ){
        for( size_t n = 0; n < motif_set.getN(); n++ ){
            Motif* motif = new Motif( *motif_set.getMotifs()[n] );
            FDR fdr( posSet, negSet, motif, bgModel, Global::cvFold, Global::mops, Global::zoops,
                     Global::savePRs, Global::savePvalues, Global::saveLogOdds );
            fdr.evaluateMotif( Global::EM, Global::CGS, Global::optimizeQ, Global::advanceEM, Global::f );
            mark( "FDR folds" );
            fdr.write( Global::outputDirectory, Global::outputFileBasename + "_motif_" + std::to_string( n + 1 ) );
            delete motif;
            mark( "FDR statistics + files" );
        }
    }

    std::cout << std::endl << "******************" << std::endl << "*   Statistics   *" << std::endl << "******************" << std::endl;
    Global::printStat();
    auto t1_wall = std::chrono::high_resolution_clock::now();
    auto t_diff = std::chrono::duration_cast<std::chrono::duration<double>>( t1_wall - t0_wall );
    std::cout << std::endl << "------ Runtime: " << t_diff.count() << " seconds -------" << std::endl;

    negSequences.reset();
    delete bgModel;
    Global::destruct();
    mark( "host teardown" );
    if( trace ) std::cerr << "[bamm host] main left at " << std::fixed << std::setprecision( 3 ) << epoch() << std::endl;
    return 0;
}
