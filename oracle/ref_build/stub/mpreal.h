// TEST INFRASTRUCTURE (oracle build only) -- not part of the product.
// Stand-in for the MPFR C++ wrapper header that the reference includes as "mpreal.h".
// The GPU box / this container have no mpfr.h, so the oracle build backs mpfr::mpreal with
// GCC's __float128 (113-bit significand).  Only the reference's compute_expms() does arithmetic
// in this type; SURVEY.md probe P5 shows 113-bit vs 256-bit is invisible after the cast to double.
#ifndef SMCB_MPREAL_SHIM_H
#define SMCB_MPREAL_SHIM_H
#include <cmath>
#include <iostream>
extern "C" {
__float128 expq(__float128); __float128 sinhq(__float128); __float128 coshq(__float128);
__float128 sqrtq(__float128); __float128 logq(__float128); __float128 fabsq(__float128);
}
namespace mpfr {
class mpreal {
  public:
    __float128 q;
    mpreal() : q(0) {}
    mpreal(double x) : q(x) {}
    mpreal(int x) : q(x) {}
    mpreal(long x) : q(x) {}
    struct raw_tag {};
    mpreal(__float128 x, raw_tag) : q(x) {}
    static void set_default_prec(int) {}
    explicit operator double() const { return (double)q; }
    explicit operator int() const { return (int)q; }
    explicit operator long() const { return (long)q; }
    mpreal &operator+=(const mpreal &o) { q += o.q; return *this; }
    mpreal &operator-=(const mpreal &o) { q -= o.q; return *this; }
    mpreal &operator*=(const mpreal &o) { q *= o.q; return *this; }
    mpreal &operator/=(const mpreal &o) { q /= o.q; return *this; }
    mpreal operator-() const { return mpreal(-q, raw_tag()); }
};
#define SMCB_BINOP(OP) \
    inline mpreal operator OP(const mpreal &a, const mpreal &b) { return mpreal(a.q OP b.q, mpreal::raw_tag()); } \
    inline mpreal operator OP(const mpreal &a, double b) { return mpreal(a.q OP (__float128)b, mpreal::raw_tag()); } \
    inline mpreal operator OP(double a, const mpreal &b) { return mpreal((__float128)a OP b.q, mpreal::raw_tag()); } \
    inline mpreal operator OP(const mpreal &a, int b) { return mpreal(a.q OP (__float128)b, mpreal::raw_tag()); } \
    inline mpreal operator OP(int a, const mpreal &b) { return mpreal((__float128)a OP b.q, mpreal::raw_tag()); }
SMCB_BINOP(+) SMCB_BINOP(-) SMCB_BINOP(*) SMCB_BINOP(/)
#undef SMCB_BINOP
#define SMCB_CMP(OP) \
    inline bool operator OP(const mpreal &a, const mpreal &b) { return a.q OP b.q; } \
    inline bool operator OP(const mpreal &a, double b) { return a.q OP (__float128)b; } \
    inline bool operator OP(double a, const mpreal &b) { return (__float128)a OP b.q; }
SMCB_CMP(<) SMCB_CMP(>) SMCB_CMP(<=) SMCB_CMP(>=) SMCB_CMP(==) SMCB_CMP(!=)
#undef SMCB_CMP
inline mpreal exp(const mpreal &x) { return mpreal(expq(x.q), mpreal::raw_tag()); }
inline mpreal sinh(const mpreal &x) { return mpreal(sinhq(x.q), mpreal::raw_tag()); }
inline mpreal cosh(const mpreal &x) { return mpreal(coshq(x.q), mpreal::raw_tag()); }
inline mpreal sqrt(const mpreal &x) { return mpreal(sqrtq(x.q), mpreal::raw_tag()); }
inline mpreal log(const mpreal &x) { return mpreal(logq(x.q), mpreal::raw_tag()); }
inline mpreal abs(const mpreal &x) { return mpreal(fabsq(x.q), mpreal::raw_tag()); }
inline mpreal fabs(const mpreal &x) { return abs(x); }
inline bool isfinite(const mpreal &x) { return std::isfinite((double)x.q); }
inline bool isnan(const mpreal &x) { return x.q != x.q; }
inline bool isinf(const mpreal &x) { return std::isinf((double)x.q); }
// the reference's common.h says `using mpfr::exp;` etc. and then calls them on doubles
using std::exp; using std::sinh; using std::cosh;
inline std::ostream &operator<<(std::ostream &o, const mpreal &x) { return o << (double)x.q; }
}  // namespace mpfr
#endif
