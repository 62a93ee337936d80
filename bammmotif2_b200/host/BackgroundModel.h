// BackgroundModel — homogeneous interpolated Markov model of order K.
// Interface mirrors the reference (src/init/BackgroundModel.h:18-54). The k-mer counts over the sequence set
// (reference: src/init/BackgroundModel.cpp:26-42) come from the device histogram bamm_seqset_count_kmers();
// the O(A^(K+1)) arithmetic (calculateV, :441-472) and the .hbcp/.hbp file formats (:48-129, :353-439) stay on the host.
// Tables are contiguous; getV() hands out a float** view (v[k][y]) into them like the reference's accessor.
#ifndef BAMM_HOST_BACKGROUNDMODEL_H_
#define BAMM_HOST_BACKGROUNDMODEL_H_

#include <string>
#include <vector>

#include "SequenceSet.h"

class BackgroundModel {
public:
    BackgroundModel( std::vector<Sequence*> sequenceSet, size_t order, std::vector<float> alpha,
                     bool interpolate = true, std::string basename = "" );
    // from pre-computed counts n[k][y] (all orders concatenated); used by tests and by callers that already hold counts
    BackgroundModel( const std::vector<uint64_t>& countsAllOrders, size_t order, std::vector<float> alpha,
                     bool interpolate = true, std::string basename = "" );
    BackgroundModel( std::string filePath );        // reads a .hbcp file
    ~BackgroundModel();

    std::string getName()       { return basename_; }
    size_t      getOrder()      { return K_; }
    float**     getV()          { return vRows_.data(); }
    const std::vector<float>&       flatV() const       { return v_; }      // [k][y] concatenated (C-ABI layout)
    const std::vector<uint64_t>&    flatCounts() const  { return n_; }
    const std::vector<float>&       getAlpha() const    { return A_; }

    void        expV();
    void        logV();
    bool        vIsLog()        { return vIsLog_; }

    void        print();
    void        write( char* dir, std::string basename );

private:
    void        allocate();
    void        calculateV();

    std::string             basename_;
    size_t                  K_ = 0;
    std::vector<float>      A_;
    bool                    interpolate_ = true;
    bool                    vIsLog_ = false;
    bool                    hasCounts_ = false;
    std::vector<size_t>     Y_;
    std::vector<size_t>     off_;       // offset of order k in the flat tables
    std::vector<uint64_t>   n_;
    std::vector<float>      v_;
    std::vector<float*>     vRows_;
};

#endif
