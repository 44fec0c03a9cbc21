// rtt_internal.h — helpers shared by the translation units of librtt_b200.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>

int rtt_internal_finish(cudaError_t e);     // launch accounting: 0 on success, the cudaError_t otherwise
int rtt_internal_have_device();
