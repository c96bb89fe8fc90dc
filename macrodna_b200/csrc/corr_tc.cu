// K2 (throughput mode) -- tcgen05 / TMEM split-precision correlation contraction.  (stub: filled in next)
#include "mcd_internal.cuh"

int mcd_launch_corr_bf16x3(mcd_context* h, const uint16_t*, int64_t, const uint16_t*, int64_t, int64_t, const double*,
                           const double*, double*, int64_t, double*, int64_t) {
  return mcd_fail(h, MCD_ERR_UNSUPPORTED, "bf16x3 correlation kernel not built");
}
