// The correlation kernels compiled with 7-row tiles (namespace irr::corr7): see the note at the top of correlation.cu.
// Only the fused TMA launcher of this translation unit is called (irr_warp_correlation_fwd_ws).
#define IRR_CORR_TH 7
#include "correlation.cu"
