// The correlation kernels compiled with 7-row tiles (namespace irr::corr7): see the note at the top of correlation.cu.
// Only the fused TMA launcher of this translation unit is called (irr_warp_correlation_fwd_ws, IRR_CORR_TH7=1).
#define IRR_CORR_TH 7
#define IRR_CORR_NS corr7
#include "correlation.cu"
