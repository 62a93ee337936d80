// Stand-in for boost::math::{digamma,lgamma}; only the out-of-scope Gibbs sampler
// (reference: src/refinement/GibbsSampling.cpp:571-859) calls these. The EM / scoring / FDR
// path used as the parity oracle never executes them. Test infrastructure only.
#pragma once
#include <cmath>
namespace boost { namespace math {
template <class T> inline T lgamma(T x) { return static_cast<T>(::lgamma(static_cast<double>(x))); }
template <class T> inline T digamma(T xin) {
    double x = static_cast<double>(xin), r = 0.0;
    while (x < 6.0) { r -= 1.0 / x; x += 1.0; }
    double f = 1.0 / (x * x);
    r += std::log(x) - 0.5 / x - f * (1.0/12 - f * (1.0/120 - f * (1.0/252 - f * (1.0/240 - f / 132))));
    return static_cast<T>(r);
}
}}  // namespace boost::math
