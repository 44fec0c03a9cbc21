// EXACT variant: compiled with -fmad=false so that every fp32 rounding step of the
// reference's eager ops is reproduced (see rtt_core.cuh).  Parity/validation tool.
#define RTT_VARIANT exact
#include "rtt_kernels.inl"
