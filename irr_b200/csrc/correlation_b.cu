// The correlation kernels compiled with a 20-row source-footprint window and a 3-deep footprint ring (namespace
// irr::corrb) — an experiment on the fused kernel's copy latency (IRR_CORR_VARB=1).
#define IRR_CORR_TH 8
#define IRR_CORR_FP_H 20
#define IRR_CORR_NFS 3
#define IRR_CORR_NS corrb
#include "correlation.cu"
