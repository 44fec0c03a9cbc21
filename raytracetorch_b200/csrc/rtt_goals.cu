// rtt_goals.cu — sensor reductions of the optimisation goals (optim/goals.py:42-96, 99-187;
// elements/sensor.py:67-176) on raw sensor records rec[m] = (x, y, z, w), w == 0 meaning "no hit".
//
// The reference compacts the hit lists with boolean gathers and runs ~40 eager elementwise /
// reduction ops per goal evaluation (and as many again in autograd).  Here a goal is two
// reduction launches forward and one elementwise launch backward, HBM-bound: 16 B per record
// read per pass, 16 B written by the backward.
//
// Sums are deterministic: each block reduces its grid-stride slice (fp32 per thread, fixed
// shuffle tree), writes one partial per sum, and the last block to finish (ticket counter) adds
// the partials in a fixed order in double precision.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rtt_b200.h"
#include "rtt_internal.h"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxBlocks = 1024;                       // RTT_SPOT_WORK = 4 * kMaxBlocks + 4
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// Block-reduce NS running sums and publish them; the last block folds all partials into out[].
template <int NS>
__device__ __forceinline__ void publish(float (&acc)[NS], float* work, float* out) {
    __shared__ float part[NS][kThreads / 32];
    __shared__ double dpart[NS][kThreads / 32];
    __shared__ bool last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        const float v = warp_sum(acc[s]);
        if (lane == 0) part[s][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < NS) {
        float v = 0.0f;
        for (int w = 0; w < kThreads / 32; ++w) v += part[threadIdx.x][w];
        work[4 * blockIdx.x + threadIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    unsigned* ticket = reinterpret_cast<unsigned*>(work + 4 * kMaxBlocks);
    if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!last) return;
    __threadfence();
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        double v = 0.0;
        for (int b = threadIdx.x; b < (int)gridDim.x; b += kThreads) v += (double)__ldcg(work + 4 * b + s);
        v = warp_sum(v);
        if (lane == 0) dpart[s][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < NS) {
        double v = 0.0;
        for (int w = 0; w < kThreads / 32; ++w) v += dpart[threadIdx.x][w];
        out[threadIdx.x] = (float)v;
    }
    if (threadIdx.x == 0) *ticket = 0u;                // leave the workspace ready for the next launch
}

__global__ void __launch_bounds__(kThreads) k_spot_moments(const float4* __restrict__ rec, long long m, int active_only,
                                                           float* out4, float* work) {
    float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) {
        const float4 r = __ldg(rec + i);
        const bool on = active_only ? (r.w > 0.0f) : true;
        if (on) { acc[0] += r.w; acc[1] += r.x * r.w; acc[2] += r.y * r.w; }
        if (r.w > 0.0f) acc[3] += 1.0f;
    }
    publish<4>(acc, work, out4);
}

__global__ void __launch_bounds__(kThreads) k_spot_moments_bwd(const float4* __restrict__ rec, long long m, int active_only,
                                                               const float* __restrict__ g3, float4* __restrict__ g_rec) {
    const float g0 = g3[0], g1 = g3[1], g2 = g3[2];
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) {
        const float4 r = __ldg(rec + i);
        const bool on = active_only ? (r.w > 0.0f) : true;
        g_rec[i] = on ? make_float4(g1 * r.w, g2 * r.w, 0.0f, g0 + g1 * r.x + g2 * r.y) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

struct Centre { float W, iW, cx, cy; bool clamped; };
__device__ __forceinline__ Centre centre_of(const float* mom4, const float* target_xy) {
    Centre c;
    c.clamped = !(mom4[0] >= 1e-12f);
    c.W = c.clamped ? 1e-12f : mom4[0];                // w_sum.clamp(min=1e-12)   optim/goals.py:170
    c.iW = 1.0f / c.W;
    if (target_xy) { c.cx = target_xy[0]; c.cy = target_xy[1]; }
    else { c.cx = mom4[1] / c.W; c.cy = mom4[2] / c.W; }
    return c;
}

__global__ void __launch_bounds__(kThreads) k_spot_size_fwd(const float4* __restrict__ rec, long long m,
                                                            const float* __restrict__ mom4,
                                                            const float* __restrict__ target_xy, float* out3, float* work) {
    const Centre c = centre_of(mom4, target_xy);
    float acc[3] = {0.0f, 0.0f, 0.0f};
    const long long stride = (long long)gridDim.x * blockDim.x;
    // four records per trip, their loads issued together: the divisions and the square root of one record overlap the
    // loads of the next ones (one load in flight per thread ran this pass at half the rate of k_spot_moments)
    constexpr int kUnroll = 4;
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < m; i0 += kUnroll * stride) {
        float4 rr[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const long long i = i0 + u * stride;
            rr[u] = (i < m) ? __ldg(rec + i) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const float4 r = rr[u];
            if (!(r.w > 0.0f)) continue;                                // optim/goals.py:165
            const float dx = r.x - c.cx, dy = r.y - c.cy;
            const float wn = r.w / c.W;
            const float q = (dx * dx + dy * dy) * wn;                   // :176-181
            const float rms = sqrtf(q);
            acc[0] += rms;
            const float a = (q > 0.0f) ? 0.5f / rms : 0.0f;             // d sqrt
            acc[1] += a * (-2.0f * dx * wn);
            acc[2] += a * (-2.0f * dy * wn);
        }
    }
    publish<3>(acc, work, out3);
}

__global__ void __launch_bounds__(kThreads) k_spot_size_bwd(const float4* __restrict__ rec, long long m,
                                                            const float* __restrict__ mom4,
                                                            const float* __restrict__ target_xy,
                                                            const float* __restrict__ out3,
                                                            const float* __restrict__ g_loss, float4* __restrict__ g_rec) {
    const Centre c = centre_of(mom4, target_xy);
    const float gL = g_loss[0];
    const bool free_centre = target_xy == nullptr;
    const float Gcx = free_centre ? out3[1] : 0.0f, Gcy = free_centre ? out3[2] : 0.0f;
    // d loss / d W: through q_i (= -L / 2W) and through the centroid (cx = Mx / W)
    const float gW = c.clamped ? 0.0f : (-0.5f * out3[0] * c.iW - (Gcx * c.cx + Gcy * c.cy) * c.iW);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) {
        const float4 r = __ldg(rec + i);
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r.w > 0.0f) {
            const float dx = r.x - c.cx, dy = r.y - c.cy;
            const float wn = r.w / c.W;
            const float r2 = dx * dx + dy * dy;
            const float q = r2 * wn;
            const float a = (q > 0.0f) ? 0.5f / sqrtf(q) : 0.0f;
            g.x = gL * (a * 2.0f * dx * wn + Gcx * wn);
            g.y = gL * (a * 2.0f * dy * wn + Gcy * wn);
            g.w = gL * (a * r2 * c.iW + (Gcx * r.x + Gcy * r.y) * c.iW + gW);
        }
        g_rec[i] = g;
    }
}

int grid_for(long long m) {
    long long g = (m + kThreads - 1) / kThreads;
    if (g > kMaxBlocks) g = kMaxBlocks;
    if (g < 1) g = 1;
    return (int)g;
}
int bad_rec(const void* p) { return !p || (reinterpret_cast<uintptr_t>(p) & 15); }

}  // namespace

extern "C" {

int rtt_spot_moments(const float* rec, int64_t m, int32_t active_only, float* out4, float* work, void* stream) {
    if (m < 0 || !out4 || !work || (m > 0 && !rec)) return RTT_E_ARG;
    if (m > 0 && bad_rec(rec)) return RTT_E_ALIGN;
    if (!rtt_internal_have_device()) return RTT_E_NO_DEVICE;
    k_spot_moments<<<grid_for(m), kThreads, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(rec), m,
                                                                       active_only, out4, work);
    return rtt_internal_finish(cudaGetLastError());
}

int rtt_spot_moments_bwd(const float* rec, int64_t m, int32_t active_only, const float* g3, float* g_rec, void* stream) {
    if (m == 0) return RTT_OK;
    if (m < 0 || !g3) return RTT_E_ARG;
    if (bad_rec(rec) || bad_rec(g_rec)) return rec && g_rec ? RTT_E_ALIGN : RTT_E_ARG;
    if (!rtt_internal_have_device()) return RTT_E_NO_DEVICE;
    k_spot_moments_bwd<<<grid_for(m) * 4, kThreads, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(rec), m, active_only, g3, reinterpret_cast<float4*>(g_rec));
    return rtt_internal_finish(cudaGetLastError());
}

int rtt_spot_size_fwd(const float* rec, int64_t m, const float* mom4, const float* target_xy, float* out3,
                      float* work, void* stream) {
    if (m < 0 || !mom4 || !out3 || !work || (m > 0 && !rec)) return RTT_E_ARG;
    if (m > 0 && bad_rec(rec)) return RTT_E_ALIGN;
    if (!rtt_internal_have_device()) return RTT_E_NO_DEVICE;
    k_spot_size_fwd<<<grid_for(m), kThreads, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(rec), m, mom4,
                                                                        target_xy, out3, work);
    return rtt_internal_finish(cudaGetLastError());
}

int rtt_spot_size_bwd(const float* rec, int64_t m, const float* mom4, const float* target_xy, const float* out3,
                      const float* g_loss, float* g_rec, void* stream) {
    if (m == 0) return RTT_OK;
    if (m < 0 || !mom4 || !out3 || !g_loss) return RTT_E_ARG;
    if (bad_rec(rec) || bad_rec(g_rec)) return rec && g_rec ? RTT_E_ALIGN : RTT_E_ARG;
    if (!rtt_internal_have_device()) return RTT_E_NO_DEVICE;
    k_spot_size_bwd<<<grid_for(m) * 4, kThreads, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(rec), m, mom4, target_xy, out3, g_loss, reinterpret_cast<float4*>(g_rec));
    return rtt_internal_finish(cudaGetLastError());
}

}  // extern "C"
