// FAST variant: default nvcc floating-point contraction (mul+add -> FMA).
#define RTT_VARIANT fast
#define RTT_APPROX 1      // MUFU rcp/sqrt instead of IEEE div/sqrt (rtt_core.cuh)
#include "rtt_kernels.inl"
