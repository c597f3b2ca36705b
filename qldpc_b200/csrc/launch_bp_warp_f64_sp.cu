// Launchers of the warp-per-shot float64 sum-product kernels (bp_warp_kernel_f64.cuh, VAR = 1, 2): G.warp_var 4 / 5.
#define QLDPC_F64_VAR 1
#define QLDPC_F64_ENTRY launch_bp_warp_f64_sp
#include "launch_bp_warp_f64_impl.h"
