// Stand-in for boost::math::beta_distribution / quantile (reference: GibbsSampling.cpp:425-426,
// out-of-scope Gibbs sampler only). quantile() is a bisection on the regularised incomplete beta
// evaluated by a continued fraction; adequate for compiling and never run on the oracle path.
#pragma once
#include <cmath>
namespace boost { namespace math {
template <class T = double> struct beta_distribution {
    T a, b;
    beta_distribution(T a_, T b_) : a(a_), b(b_) {}
};
namespace bamm_shim {
inline double betacf(double a, double b, double x) {
    const int MAXIT = 300; const double EPS = 3e-14, FPMIN = 1e-300;
    double qab = a + b, qap = a + 1, qam = a - 1, c = 1, d = 1 - qab * x / qap;
    if (std::fabs(d) < FPMIN) d = FPMIN; d = 1 / d; double h = d;
    for (int m = 1; m <= MAXIT; m++) {
        int m2 = 2 * m; double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
        d = 1 + aa * d; if (std::fabs(d) < FPMIN) d = FPMIN; c = 1 + aa / c; if (std::fabs(c) < FPMIN) c = FPMIN;
        d = 1 / d; h *= d * c;
        aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
        d = 1 + aa * d; if (std::fabs(d) < FPMIN) d = FPMIN; c = 1 + aa / c; if (std::fabs(c) < FPMIN) c = FPMIN;
        d = 1 / d; double del = d * c; h *= del; if (std::fabs(del - 1) < EPS) break;
    }
    return h;
}
inline double ibeta(double a, double b, double x) {
    if (x <= 0) return 0; if (x >= 1) return 1;
    double bt = std::exp(::lgamma(a + b) - ::lgamma(a) - ::lgamma(b) + a * std::log(x) + b * std::log(1 - x));
    if (x < (a + 1) / (a + b + 2)) return bt * betacf(a, b, x) / a;
    return 1 - bt * betacf(b, a, 1 - x) / b;
}
}  // namespace bamm_shim
template <class T, class P> inline T quantile(const beta_distribution<T>& d, P p) {
    double lo = 0, hi = 1;
    for (int i = 0; i < 200; i++) { double mid = 0.5 * (lo + hi); if (bamm_shim::ibeta(d.a, d.b, mid) < p) lo = mid; else hi = mid; }
    return static_cast<T>(0.5 * (lo + hi));
}
}}  // namespace boost::math
