// FAST variant: default nvcc floating-point contraction (mul+add -> FMA).
#define RTT_VARIANT fast
#define RTT_APPROX 1      // MUFU rcp/sqrt instead of IEEE div/sqrt (rtt_core.cuh)
#define RTT_TILE_LEAN 1   // lean root selection of lens faces in the tile kernel (rtt_tile.cuh)
#include "rtt_kernels.inl"
