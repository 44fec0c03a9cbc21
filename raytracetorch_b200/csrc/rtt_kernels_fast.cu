// FAST variant: default nvcc floating-point contraction (mul+add -> FMA).
#define RTT_VARIANT fast
#include "rtt_kernels.inl"
