// Device.h — glue between the host classes and the C ABI (include/bamm_b200.h).
// Error convention of the reference on this path: message on stderr, then exit(1)
// (e.g. src/init/SequenceSet.cpp:144-149); BAMM_CHECK turns a C-ABI status into exactly that.
#ifndef BAMM_HOST_DEVICE_H_
#define BAMM_HOST_DEVICE_H_

#include <cstdio>
#include <cstdlib>

#include "../../include/bamm_b200.h"

#define BAMM_CHECK( call )                                                                      \
    do {                                                                                        \
        const int bamm_rc_ = ( call );                                                          \
        if( bamm_rc_ != BAMM_OK ){                                                              \
            std::fprintf( stderr, "Error: %s (%s, status %d)\n", bamm_last_error(), #call, bamm_rc_ ); \
            std::exit( 1 );                                                                     \
        }                                                                                       \
    } while( 0 )

#endif
