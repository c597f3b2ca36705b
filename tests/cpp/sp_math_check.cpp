// Host build of qldpc_b200/csrc/sp_math.cuh for tests/test_host.py: array versions of the two functions.
// -DSPM_EMULATE_RCP replaces the host division by a model of the device sequence (20-bit reciprocal seed + cubic Newton step).
#include <cstddef>
#include "../../qldpc_b200/csrc/sp_math.cuh"
extern "C" {
void sp_tanh_half(const double *x, double *y, size_t n) { for (size_t i = 0; i < n; ++i) y[i] = qldpc::spm_tanh_half(x[i]); }
void sp_2atanh(const double *x, double *y, size_t n) { for (size_t i = 0; i < n; ++i) y[i] = qldpc::spm_2atanh_clipped(x[i]); }
}
