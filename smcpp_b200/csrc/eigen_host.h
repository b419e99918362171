// smcpp_b200 -- host-side real non-symmetric eigensystems of diag(e_key) Td^T.
// Replaces the EigenSolver loop of the reference's TransitionBundle::update (src/transition_bundle.cpp:14-25)
// and struct eigensystem (include/transition_bundle.h:9-30): eigenvectors P (unit 2-norm columns),
// Pinv = P^-1 (complex inverse), of which only the REAL parts are kept, d_r = Re(d), scale = max|d|,
// d_r_scaled = d_r / scale, cplx = any Im(d) != 0.
#pragma once
#include <cstdint>
#include <string>

namespace smcb {

// One matrix: A row-major n x n.  Outputs row-major; returns 0 or non-zero (msg filled) if the QR
// iteration does not converge.
int host_eig_real_general(int n, const double *A, double *P_r, double *Pinv_r, double *d_r, double *d_i,
                          std::string *msg);

// All eigen keys of one E-step.  T row-major M x M, E row-major K x M.
int host_eigensystems(int M, int K, int n_eig, const int32_t *eig_keys, const double *T, const double *E, double *P,
                      double *Pinv, double *d, double *d_scaled, double *scale, int32_t *cplx, std::string *msg);

}  // namespace smcb
