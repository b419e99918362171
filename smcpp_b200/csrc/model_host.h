// smcpp_b200 -- host builders of pi / transition / emission table (see model_host.cpp).
#pragma once
#include <cstdint>
#include <string>

namespace smcb {

int host_initial_distribution(int M, const double *hidden_states, int n_pieces, const double *a, const double *s, double *pi);
int host_average_coal_times(int M, const double *hidden_states, int n_pieces, const double *a, const double *s, double *out);
int host_transition(int M, const double *hidden_states, int n_pieces, const double *a, const double *s, double rho, double *T);
// sfs: [M][na[0]+1][sfs_dim] row-major, sfs_dim = prod_{p>=1}(na[p]+1) * prod_p (n[p]+1); keys: [K][3*npop]; E: [K][M]
int host_emission(int npop, const int *n, const int *na, int M, const double *hidden_states, int n_pieces, const double *a,
                  const double *s, double theta, double alpha, double pol_err, const double *sfs, int K, const int32_t *keys,
                  double *E, std::string *msg);

}  // namespace smcb
