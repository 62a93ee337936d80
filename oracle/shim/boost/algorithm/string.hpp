// Minimal stand-in for the two Boost.StringAlgo calls the reference's FASTA reader makes
// (reference: src/init/SequenceSet.cpp:138-139). Test infrastructure only: lets the UNMODIFIED
// reference sources compile in an image without Boost. Not part of the product.
#pragma once
#include <string>
#include <vector>
namespace boost {
struct bamm_shim_any_of { std::string set; };
inline bamm_shim_any_of is_any_of(const char* s) { return bamm_shim_any_of{std::string(s)}; }
template <class Seq>
inline Seq& split(Seq& out, const std::string& in, const bamm_shim_any_of& pred) {
    out.clear();
    std::string cur;
    for (char c : in) {
        if (pred.set.find(c) != std::string::npos) { out.push_back(cur); cur.clear(); }
        else cur.push_back(c);
    }
    out.push_back(cur);
    return out;
}
}  // namespace boost
