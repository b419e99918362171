// TEST INFRASTRUCTURE (oracle build only): empty stand-in for gmpxx.h.  Only the reference's
// CSFS / Moran code (off the E-step path, never executed by the oracle) mentions mpq_class.
#pragma once
struct smcb_mpq_stub {};
class mpq_class {
  public:
    mpq_class() {}
    mpq_class(long) {}
    mpq_class(long, long) {}
    const smcb_mpq_stub *get_mpq_t() const { return nullptr; }
};
inline double mpq_get_d(const smcb_mpq_stub *) { return 0.0; }
