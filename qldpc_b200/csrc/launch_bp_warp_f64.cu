// Launchers of the warp-per-shot float64 min-sum kernels (bp_warp_kernel_f64.cuh, VAR = 0): the bit-exact parity mode.
#define QLDPC_F64_VAR 0
#define QLDPC_F64_ENTRY launch_bp_warp_f64
#include "launch_bp_warp_f64_impl.h"
