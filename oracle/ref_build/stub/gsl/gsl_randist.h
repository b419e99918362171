// TEST INFRASTRUCTURE (oracle build only): closed-form hypergeometric pmf replacing the one GSL
// call on the construction path (reference include/marginalize_key.h:47).
#pragma once
#include <cmath>
static inline double smcb_lchoose(double n, double r) { return std::lgamma(n + 1.) - std::lgamma(r + 1.) - std::lgamma(n - r + 1.); }
// P(k white | t draws without replacement from n1 white + n2 black)
static inline double gsl_ran_hypergeometric_pdf(unsigned int k, unsigned int n1, unsigned int n2, unsigned int t)
{
    if (t > n1 + n2) t = n1 + n2;
    if (k > n1 || k > t) return 0.;
    if (t > n2 && k + n2 < t) return 0.;
    return std::exp(smcb_lchoose(n1, k) + smcb_lchoose(n2, t - k) - smcb_lchoose(n1 + n2, t));
}
