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


// ============================================================================================
// Per-id sensor moments (elements/sensor.py:87-176, Sensor.getSpotSizeParallel_xy)
// ============================================================================================
// The reference filters the hit lists with isin, sorts the query ids, maps every hit to its group with searchsorted and
// reduces with three scatter_adds (plus the boolean gathers of getHitsTensors).  Here: one pass over the dense records
// [m,4] + int8 ids per reduction.  A 256-entry table (id + 128 -> group, -1 = not queried) lives in shared memory;
// a thread keeps the running sums of the group it is in (ray ids come in runs: one bundle after the other) and
// touches the block's shared accumulators only when the group changes; blocks publish their partials and the last
// block folds them in a fixed order in double precision (deterministic, like the reductions above).
constexpr int kIdBlocks = 296;                         // RTT_SPOT_ID_WORK = kIdBlocks * 256 * 4 + 4 floats

// MODE 0: v = (w, w x, w y, [w > 0])                                   -> out[k] = per-group sums
// MODE 1: v = (w (|dx|^p + |dy|^p), w p |dx|^(p-1) sgn dx, w p |dy|^(p-1) sgn dy, 0), d = xy - centre[k]
template <int MODE>
__global__ void __launch_bounds__(kThreads) k_spot_id_sums(const float4* __restrict__ rec, const signed char* __restrict__ ids,
                                                           long long m, const int* __restrict__ group_of, int K,
                                                           const float* __restrict__ centres, float p,
                                                           float* __restrict__ out, float* __restrict__ work) {
    __shared__ int lut[256];
    __shared__ float acc[256 * 4];
    __shared__ bool last;
    for (int idx = threadIdx.x; idx < 256; idx += kThreads) lut[idx] = group_of[idx];
    for (int idx = threadIdx.x; idx < 4 * K; idx += kThreads) acc[idx] = 0.0f;
    __syncthreads();
    int cur = -1;
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f, cx = 0.0f, cy = 0.0f;
    auto flush = [&]() {
        if (cur >= 0) {
            if (a0 != 0.0f) atomicAdd(acc + 4 * cur, a0);
            if (a1 != 0.0f) atomicAdd(acc + 4 * cur + 1, a1);
            if (a2 != 0.0f) atomicAdd(acc + 4 * cur + 2, a2);
            if (a3 != 0.0f) atomicAdd(acc + 4 * cur + 3, a3);
        }
        a0 = a1 = a2 = a3 = 0.0f;
    };
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) {
        const int g = lut[(int)ids[i] + 128];
        if (g < 0) continue;
        const float4 r = __ldg(rec + i);
        if (r.w == 0.0f) continue;                                      // every sum is weighted by w
        if (g != cur) {
            flush();
            cur = g;
            if (MODE == 1) { cx = centres[2 * g]; cy = centres[2 * g + 1]; }
        }
        if (MODE == 0) {
            a0 += r.w; a1 += r.w * r.x; a2 += r.w * r.y; a3 += (r.w > 0.0f) ? 1.0f : 0.0f;
        } else {
            const float dx = r.x - cx, dy = r.y - cy;
            if (p == 2.0f) {
                a0 += r.w * (dx * dx + dy * dy); a1 += r.w * 2.0f * dx; a2 += r.w * 2.0f * dy;
            } else {
                const float ax = fabsf(dx), ay = fabsf(dy);
                const float px = powf(ax, p - 1.0f), py = powf(ay, p - 1.0f);
                a0 += r.w * (px * ax + py * ay);
                a1 += r.w * p * px * (dx > 0.0f ? 1.0f : (dx < 0.0f ? -1.0f : 0.0f));
                a2 += r.w * p * py * (dy > 0.0f ? 1.0f : (dy < 0.0f ? -1.0f : 0.0f));
            }
        }
    }
    flush();
    __syncthreads();
    for (int idx = threadIdx.x; idx < 4 * K; idx += kThreads) work[(size_t)blockIdx.x * 4 * K + idx] = acc[idx];
    __threadfence();
    __syncthreads();
    unsigned* ticket = reinterpret_cast<unsigned*>(work + (size_t)kIdBlocks * 256 * 4);
    if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!last) return;
    __threadfence();
    for (int idx = threadIdx.x; idx < 4 * K; idx += kThreads) {
        double v = 0.0;
        for (int b = 0; b < (int)gridDim.x; ++b) v += (double)__ldcg(work + (size_t)b * 4 * K + idx);
        out[idx] = (float)v;
    }
    if (threadIdx.x == 0) *ticket = 0u;                // leave the workspace ready for the next launch
}

// coef[k] = (cx, cy, a, bx, by, sW, -, -):  g_rec = (a w (p |dx|^(p-1) sgn dx - bx), a w (p |dy|^(p-1) sgn dy - by), 0,
//                                                    a ((|dx|^p + |dy|^p) - bx dx - by dy - sW))
__global__ void __launch_bounds__(kThreads) k_spot_id_size_bwd(const float4* __restrict__ rec, const signed char* __restrict__ ids,
                                                               long long m, const int* __restrict__ group_of, int K,
                                                               const float* __restrict__ coef, float p,
                                                               float4* __restrict__ g_rec) {
    __shared__ int lut[256];
    __shared__ float cf[256 * 8];
    for (int idx = threadIdx.x; idx < 256; idx += kThreads) lut[idx] = group_of[idx];
    for (int idx = threadIdx.x; idx < 8 * K; idx += kThreads) cf[idx] = coef[idx];
    __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) {
        const int g = lut[(int)ids[i] + 128];
        float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g >= 0) {
            const float4 r = __ldg(rec + i);
            const float* c = cf + 8 * g;
            const float dx = r.x - c[0], dy = r.y - c[1];
            float sx, sy, pw;
            if (p == 2.0f) { sx = 2.0f * dx; sy = 2.0f * dy; pw = dx * dx + dy * dy; }
            else {
                const float ax = fabsf(dx), ay = fabsf(dy);
                const float px = powf(ax, p - 1.0f), py = powf(ay, p - 1.0f);
                sx = p * px * (dx > 0.0f ? 1.0f : (dx < 0.0f ? -1.0f : 0.0f));
                sy = p * py * (dy > 0.0f ? 1.0f : (dy < 0.0f ? -1.0f : 0.0f));
                pw = px * ax + py * ay;
            }
            out.x = c[2] * r.w * (sx - c[3]);
            out.y = c[2] * r.w * (sy - c[4]);
            out.w = c[2] * (pw - c[3] * dx - c[4] * dy - c[5]);
        }
        g_rec[i] = out;
    }
}

int grid_id(long long m) {
    long long g = (m + kThreads - 1) / kThreads;
    if (g > kIdBlocks) g = kIdBlocks;
    if (g < 1) g = 1;
    return (int)g;
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

int rtt_spot_id_moments(const float* rec, const int8_t* ids, int64_t m, const int32_t* group_of, int32_t n_groups,
                        float* out, float* work, void* stream) {
    if (m < 0 || !group_of || !out || !work || n_groups < 1 || n_groups > 256 || (m > 0 && (!rec || !ids))) return RTT_E_ARG;
    if (m > 0 && bad_rec(rec)) return RTT_E_ALIGN;
    if (!rtt_internal_have_device()) return RTT_E_NO_DEVICE;
    k_spot_id_sums<0><<<grid_id(m), kThreads, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(rec), reinterpret_cast<const signed char*>(ids), m, group_of, n_groups, nullptr,
        2.0f, out, work);
    return rtt_internal_finish(cudaGetLastError());
}

int rtt_spot_id_size(const float* rec, const int8_t* ids, int64_t m, const int32_t* group_of, int32_t n_groups,
                     const float* centres, float norm_ord, float* out, float* work, void* stream) {
    if (m < 0 || !group_of || !centres || !out || !work || n_groups < 1 || n_groups > 256 || (m > 0 && (!rec || !ids)))
        return RTT_E_ARG;
    if (!(norm_ord >= 1.0f)) return RTT_E_ARG;
    if (m > 0 && bad_rec(rec)) return RTT_E_ALIGN;
    if (!rtt_internal_have_device()) return RTT_E_NO_DEVICE;
    k_spot_id_sums<1><<<grid_id(m), kThreads, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(rec), reinterpret_cast<const signed char*>(ids), m, group_of, n_groups, centres,
        norm_ord, out, work);
    return rtt_internal_finish(cudaGetLastError());
}

int rtt_spot_id_size_bwd(const float* rec, const int8_t* ids, int64_t m, const int32_t* group_of, int32_t n_groups,
                         const float* coef, float norm_ord, float* g_rec, void* stream) {
    if (m == 0) return RTT_OK;
    if (m < 0 || !ids || !group_of || !coef || n_groups < 1 || n_groups > 256 || !(norm_ord >= 1.0f)) return RTT_E_ARG;
    if (bad_rec(rec) || bad_rec(g_rec)) return rec && g_rec ? RTT_E_ALIGN : RTT_E_ARG;
    if (!rtt_internal_have_device()) return RTT_E_NO_DEVICE;
    k_spot_id_size_bwd<<<grid_id(m) * 4, kThreads, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(rec), reinterpret_cast<const signed char*>(ids), m, group_of, n_groups, coef,
        norm_ord, reinterpret_cast<float4*>(g_rec));
    return rtt_internal_finish(cudaGetLastError());
}

}  // extern "C"
