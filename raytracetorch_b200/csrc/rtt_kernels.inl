// rtt_kernels.inl — CUDA kernels of the ray-propagation path (sm_100a).
//
// Included twice: rtt_kernels_fast.cu (FMA contraction on) and rtt_kernels_exact.cu
// (-fmad=false; every rounding step of the reference's eager fp32 ops is kept).  The
// including file defines RTT_VARIANT (fast|exact).
//
// Execution model
//   * the surface table ([S,48] f32 + [S,16] i32, a few KB) is read from device memory and
//     staged once per thread block in shared memory together with per-row derived constants;
//     every row access afterwards is a shared-memory broadcast
//   * one thread owns its rays and keeps their state (p, d, I, hit mask) in registers through
//     the whole surface stack; blocks are persistent and walk the bundle with a grid stride, so
//     the staging cost is paid once per block
//   * FAST sequential forward = k_trace_seq_fwd_tile (rtt_tile.cuh): element-frame ray state,
//     2 rays per thread, lens-edge culling; EXACT = k_trace_seq_fwd in the reference's order
//   * sensor hits are binned in the same kernel (no hit lists unless requested)
//   * the adjoint kernels recompute the forward per ray (checkpoints = incoming (p, d) of
//     each interaction), compact each chunk to the rays with non-zero upstream gradients, and
//     reduce parameter gradients warp -> block (shared) -> global
//   * the non-sequential forward refills a lane as soon as its ray has ended
#include <cuda_runtime.h>
#include <type_traits>
#include "rtt_core.cuh"
#include "rtt_tile.cuh"
#include "rtt_pair.cuh"
#include "rtt_lean.cuh"
#include "rtt_kernels_decl.h"

namespace rtt {
namespace RTT_VARIANT {

constexpr int kThreads = 256;
constexpr unsigned kFull = 0xffffffffu;



// Shared-memory image of the table.
struct SmemTable {
    RowDev* rows;        // [S]
    float* lut_ni;       // [L*S]
    float* lut_no;       // [L*S]
    float2* mu;          // [L*S]   (no/ni, ni/no) = (mu_enter, mu_exit): one 8-byte load per lookup
    float* lut_w;        // [L]
};

__host__ __device__ inline size_t smem_table_bytes(int S, int L) {
    return sizeof(RowDev) * (size_t)S + sizeof(float) * ((size_t)4 * L * S + (size_t)L + 4);
}

__device__ __forceinline__ SmemTable carve(unsigned char* base, int S, int L) {
    SmemTable T;
    T.rows = reinterpret_cast<RowDev*>(base);
    float* f = reinterpret_cast<float*>(base + sizeof(RowDev) * (size_t)S);
    T.lut_ni = f; f += (size_t)L * S;
    T.lut_no = f; f += (size_t)L * S;
    T.mu = reinterpret_cast<float2*>(f); f += (size_t)2 * L * S;     // 8-byte aligned: 256 S + 8 L S bytes into the table
    T.lut_w = f;
    return T;
}

// Cooperative staging: coalesced copy of the rows, then one thread per row derives constants.
__device__ __forceinline__ void stage_table(const TableDev& tab, SmemTable& T) {
    const int S = tab.S, L = tab.L;
    for (int idx = threadIdx.x; idx < S * RTT_ROW_F; idx += blockDim.x)
        T.rows[idx / RTT_ROW_F].f[idx % RTT_ROW_F] = tab.f[idx];
    for (int idx = threadIdx.x; idx < S * RTT_ROW_I; idx += blockDim.x)
        T.rows[idx / RTT_ROW_I].i[idx % RTT_ROW_I] = tab.i[idx];
    for (int idx = threadIdx.x; idx < L * S; idx += blockDim.x) {
        const float ni = tab.lut[2 * idx], no = tab.lut[2 * idx + 1];
        T.lut_ni[idx] = ni; T.lut_no[idx] = no;
        T.mu[idx] = make_float2(no / ni, ni / no);
    }
    for (int idx = threadIdx.x; idx < L; idx += blockDim.x) T.lut_w[idx] = tab.lut_w[idx];
    __syncthreads();
    for (int r = threadIdx.x; r < S; r += blockDim.x) prepare_row(T.rows[r]);
    __syncthreads();
}

// nearest sample wavelength, first minimum (oracle: torch.argmin)
__device__ __forceinline__ int wavelength_index(const SmemTable& T, int L, float w) {
    int best = 0;
    float bd = fabsf(w - T.lut_w[0]);
    // straight-line for up to four sample wavelengths (the usual F / d / C triple): the rolled loop unrolls into an
    // eight-wide body plus remainder chains that a three-entry table only jumps through
#pragma unroll
    for (int l = 1; l < 4; ++l) {
        if (l < L) {
            const float dd = fabsf(w - T.lut_w[l]);
            if (dd < bd) { bd = dd; best = l; }
        }
    }
#pragma unroll 1
    for (int l = 4; l < L; ++l) {
        const float dd = fabsf(w - T.lut_w[l]);
        if (dd < bd) { bd = dd; best = l; }
    }
    return best;
}

struct Ior { float ni, no, mu_enter, mu_exit; };

__device__ __forceinline__ Ior row_ior(const SmemTable& T, int S, int L, int r, int lam) {
    Ior q;
    if (L > 0) {
        const int idx = lam * S + r;
        const float2 m = T.mu[idx];
        q.ni = T.lut_ni[idx]; q.no = T.lut_no[idx]; q.mu_enter = m.x; q.mu_exit = m.y;
    } else {
        const RowDev& R = T.rows[r];
        q.ni = R.f[RTT_F_IOR_IN]; q.no = R.f[RTT_F_IOR_OUT]; q.mu_enter = R.f[D_MU_ENTER]; q.mu_exit = R.f[D_MU_EXIT];
    }
    return q;
}

template <class K>
__device__ __forceinline__ bool uses_ior(const RowDev& R) {
    return K::phys(R) == RTT_PHYS_SNELL || K::phys(R) == RTT_PHYS_FRESNEL;
}

__device__ __forceinline__ V3 load3(const float* a, long long i) { return v3(a[3 * i], a[3 * i + 1], a[3 * i + 2]); }
__device__ __forceinline__ void store3(float* a, long long i, V3 v) { a[3 * i] = v.x; a[3 * i + 1] = v.y; a[3 * i + 2] = v.z; }

// Ray i of a launch: read from the caller's arrays, or generated from the ray source (no HBM input).
struct RayIn { V3 p, d; float I, wav; };
template <class Args>
__device__ __forceinline__ RayIn fetch_ray(const Args& a, SourceKey key, long long i, bool want_wav) {
    RayIn r;
    if (a.src.kind >= 0) {
        source_ray(a.src, key, i, r.p, r.d);
        r.I = a.src.intensity; r.wav = a.src.wavelength;
    } else {
        r.p = load3(a.pos, i); r.d = load3(a.dir, i);
        r.I = a.inten[i];
        r.wav = want_wav ? a.wav[i] : 0.0f;
    }
    return r;
}
// GEN fixed at compile time: the kernel build for rays in memory carries no ray-source code (Philox, sincosf, acosf,
// camera pixel math — several hundred instructions per fetch site) and the build for generated rays no ray loads.
// The dead half was never executed, but it changed the register allocation and the layout of the hot loop: a cold edit
// inside source_ray moved the tile kernel by 6 % on C4 (profiles/r2_fwd_pair_ab.md).
template <bool GEN, class Args>
__device__ __forceinline__ RayIn fetch_ray_t(const Args& a, SourceKey key, long long i, bool want_wav) {
    RayIn r;
    if (GEN) {
        source_ray(a.src, key, i, r.p, r.d);
        r.I = a.src.intensity; r.wav = a.src.wavelength;
    } else {
        r.p = load3(a.pos, i); r.d = load3(a.dir, i);
        r.I = a.inten[i];
        r.wav = want_wav ? a.wav[i] : 0.0f;
    }
    return r;
}
template <class Args>
__device__ __forceinline__ SourceKey fetch_key(const Args& a) {
    SourceKey k; k.key = 0ull; k.base = 0ull;
    if (a.src.kind >= 0) k = source_key(a.src);
    return k;
}

// ---- sensor image accumulation ---------------------------------------------------------------
// Focused bundles put ~all rays of a launch into a handful of bins (C1: a 1e8-ray bundle focused
// onto a few pixels), so plain global atomics serialise in L2.  Two levels of privatisation:
//   1. warp: lanes that target the same bin are found with match.any and summed with shuffles;
//      one lane per distinct bin goes on;
//   2. block: a small direct-mapped cache of (bin, partial sum) pairs in shared memory absorbs
//      the hot bins (shared-memory atomics); a lane whose slot is owned by another bin falls
//      through to a global atomic, so spread-out images (low contention anyway) only pay one
//      shared-memory CAS.  The cache is flushed with one global atomic per occupied slot when
//      the persistent block has finished its rays.
// Zero weights (dead rays that still cross the sensor) add nothing and are skipped.
constexpr int kImgLog = 12;                // default: 4096 (bin, partial sum) pairs = 32 KB per block
constexpr int kImgSlots = 1 << kImgLog;
constexpr int kImgProbes = 4;             // linear probes before falling through to a global atomic
constexpr int kImgMaxFails = 256;         // spread-out image (the cache is full of other bins): stop probing, go global directly
constexpr int kImgKeyBits = 28;           // key = sensor slot << 28 | flat bin index

// LOG = log2 of the slot count: the kernels that also stage ray tiles in shared memory carry a smaller cache
template <int LOG>
struct ImgCacheT {
    int* tag;          // [1 << LOG], -1 = empty
    float* val;        // [1 << LOG]
    int* fails;        // [1] probes that found no slot; the cache switches itself off beyond kImgMaxFails
};
typedef ImgCacheT<kImgLog> ImgCache;

template <int LOG = kImgLog>
__host__ __device__ constexpr size_t img_cache_bytes() { return ((size_t)1 << LOG) * 8 + 16; }

template <int LOG = kImgLog>
__device__ __forceinline__ ImgCacheT<LOG> img_cache_carve(unsigned char* base) {
    ImgCacheT<LOG> c;
    c.tag = reinterpret_cast<int*>(base);
    c.val = reinterpret_cast<float*>(c.tag + (1 << LOG));
    c.fails = reinterpret_cast<int*>(c.val + (1 << LOG));
    return c;
}

template <int LOG>
__device__ __forceinline__ void img_cache_init(ImgCacheT<LOG> c) {
    for (int idx = threadIdx.x; idx < (1 << LOG); idx += blockDim.x) { c.tag[idx] = -1; c.val[idx] = 0.0f; }
    if (threadIdx.x == 0) *c.fails = 0;
}

template <int LOG>
__device__ __forceinline__ void img_cache_add(ImgCacheT<LOG> c, float* image, int slot, int bin, float w) {
    if (*reinterpret_cast<volatile int*>(c.fails) >= kImgMaxFails) {     // cache switched off: one RED per deposit
        atomicAdd(image + bin, w);
        return;
    }
    const unsigned conv = __activemask();
    const int key = (slot << kImgKeyBits) | bin;
    const unsigned peers = __match_any_sync(conv, key);
    float sum = w;
    if (peers == kFull) {
        // the whole warp lands in one bin (focused bundle): butterfly sum, lane 0 goes on
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(kFull, sum, o);
        if (threadIdx.x & 31) return;
    } else if (__popc(peers) > 1) {
        // a few lanes share a bin: pairs are folded with one shuffle, larger groups go to the cache lane by
        // lane (shared-memory atomics on one address serialise in the LSU, not in issue slots)
        if (__popc(peers) == 2) {
            const int hi = 31 - __clz((int)peers);
            const float other = __shfl_sync(peers, w, hi);
            if ((int)(threadIdx.x & 31) == hi) return;
            sum = w + other;
        }
    }
    if (*reinterpret_cast<volatile int*>(c.fails) < kImgMaxFails) {
        int s = (int)(((unsigned)key * 2654435761u) >> (32 - LOG));
#pragma unroll 1
        for (int probe = 0; probe < kImgProbes; ++probe) {
            const int old = atomicCAS(c.tag + s, -1, key);
            if (old == -1 || old == key) { atomicAdd(c.val + s, sum); return; }
            s = (s + 1) & ((1 << LOG) - 1);
        }
        atomicAdd(c.fails, 1);
    }
    atomicAdd(image + bin, sum);
}

template <int LOG>
__device__ __forceinline__ void img_cache_flush(ImgCacheT<LOG> c, const SensorDev* sens) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < (1 << LOG); idx += blockDim.x) {
        const int key = c.tag[idx];
        if (key >= 0 && c.val[idx] != 0.0f)
            atomicAdd(sens[key >> kImgKeyBits].image + (key & ((1 << kImgKeyBits) - 1)), c.val[idx]);
    }
}

// `ordinal` = how many times this ray has interacted with this sensor before, `n` = bundle size
template <int LOG>
__device__ __forceinline__ void sensor_deposit(const SensorDev& sd, ImgCacheT<LOG> c, int slot, long long i, V3 hl, float w,
                                               int lam, int ordinal = 0, long long n = 0) {
    if (sd.record && ordinal < sd.K) {
        float4* rec = reinterpret_cast<float4*>(sd.record);
        rec[(long long)ordinal * n + i] = make_float4(hl.x, hl.y, hl.z, w);
    }
    if (sd.image && w != 0.0f) {
        int ix, iy;
        if (sensor_bin(hl.x, hl.y, sd.x0, sd.y0, sd.sx, sd.sy, sd.W, sd.H, &ix, &iy)) {
            const int ch = (sd.C > 1) ? min(lam, sd.C - 1) : 0;
            img_cache_add(c, sd.image, slot, (ch * sd.H + iy) * sd.W + ix, w);
        }
    }
}

// ============================================================================================
// Sequential trace, forward (scene/sequential.py:12-36)
// ============================================================================================

// One row of the sequential walk for one ray (scene/sequential.py:19-34), kinds fixed by K.
template <class K, int LOG = kImgLog>
__device__ __forceinline__ void seq_row(const SmemTable& T, int S, int L, int r, int lam, long long i,
                                        const SeqFwdArgs& a, ImgCacheT<LOG> cache, V3& p, V3& d, float& I,
                                        unsigned long long& mask, unsigned long long bit) {
    Frames F; Roots q; float t; int which;
    if (!intersect<true, K>(T.rows, r, p, d, F, q, t, which)) return;
    const RowDev& R = T.rows[r];
    Ior io;
    io.ni = io.no = 1.0f; io.mu_enter = io.mu_exit = 0.0f;
    if (uses_ior<K>(R)) io = row_ior(T, S, L, r, lam);
    const Step s = interact<K>(R, F, t, p, d, io.mu_enter, io.mu_exit, make_aux<K>(T.rows, R, io.ni, io.no, i, r, 0));
    if (K::sensor(R)) {
        const int slot = R.i[RTT_I_SENSOR];
        if (slot >= 0 && slot < a.n_sens) sensor_deposit(a.sens[slot], cache, slot, i, s.hit_local, I, lam);
    }
    p = s.hit_global; d = s.new_dir; I = I * s.mod;
    mask |= bit;
}

__global__ void __launch_bounds__(kThreads) RTT_NAME(k_trace_seq_fwd)(const __grid_constant__ SeqFwdArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmemTable T = carve(smem_raw, a.tab.S, a.tab.L);
    ImgCache cache = img_cache_carve(smem_raw + ((smem_table_bytes(a.tab.S, a.tab.L) + 15) / 16) * 16);
    img_cache_init(cache);
    stage_table(a.tab, T);
    const int S = a.tab.S, L = a.tab.L;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const SourceKey skey = fetch_key(a);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
        const RayIn ray = fetch_ray(a, skey, i, L > 0);
        V3 p = ray.p, d = ray.d;
        float I = ray.I;
        const int lam = (L > 0) ? wavelength_index(T, L, ray.wav) : 0;
        unsigned long long mask = 0ull, bit = 1ull;
        int op_next = T.rows[0].i[DI_OPCODE];
        const int rows_walked = finite_ray(p, d) ? S : 0;               // NaN / inf rays hit nothing (rtt_tile.cuh)
        for (int r = 0; r < rows_walked; ++r, bit += bit) {
            const int op = op_next;                                     // fetched one row ahead: the
            op_next = T.rows[(r + 1 < S) ? r + 1 : r].i[DI_OPCODE];     // LDS -> BRX latency is hidden
            switch (op) {                                               // warp-uniform
#define RTT_X(OP, SURF, BOUND, SHAPE, PHYS, IDENT, SENSOR)                                            \
                case OP: seq_row<KStatic<SURF, BOUND, SHAPE, PHYS, IDENT, SENSOR>>(T, S, L, r, lam, i, a, cache, p, d, I, mask, bit); break;
                RTT_ROW_SPECS(RTT_X)
#undef RTT_X
                default: seq_row<KDyn>(T, S, L, r, lam, i, a, cache, p, d, I, mask, bit); break;
            }
        }
        if (a.opos) { store3(a.opos, i, p); store3(a.odir, i, d); a.ointen[i] = I; }
        if (a.hitmask) a.hitmask[i] = mask;
    }
    img_cache_flush(cache, a.sens);
}

#if defined(RTT_APPROX)
// ============================================================================================
// Sequential trace, forward, FAST variant: frame-resident ray tiles (rtt_tile.cuh)
// ============================================================================================
// Row loop outside, rays inside: every thread carries RPT rays in registers and walks the table
// once for all of them, so the opcode dispatch, the row constants (shared-memory loads) and the
// frame change of a row are paid once per RPT rays, and the RPT independent dependency chains
// give the scheduler instruction-level parallelism on top of the resident warps.

// Shared-memory layout of the tile kernel: [image cache | Xf[RTT_MAX_ROWS + 1] | table]; xf[r] = frame(r-1) -> frame(r),
// xf[S] = frame(S-1) -> global.  Everything the hot loop touches sits at a compile-time offset, so no pointer has to
// be kept in — or recomputed into — registers.
constexpr size_t kTileOffXf = ((size_t)kImgSlots * 8 + 16 + 15) / 16 * 16;
constexpr size_t kTileOffTable = (kTileOffXf + sizeof(Xf) * (RTT_MAX_ROWS + 1) + 15) / 16 * 16;
__host__ __device__ inline size_t tile_smem_bytes(int S, int L) { return kTileOffTable + smem_table_bytes(S, L); }

__device__ __forceinline__ void stage_tile(SmemTable& T, int S, Xf* xf) {
    for (int r = threadIdx.x; r <= S; r += blockDim.x) {
        xf[r] = make_xf(r > 0 ? &T.rows[r - 1] : nullptr, r < S ? &T.rows[r] : nullptr);
        if (r < S) T.rows[r].i[DI_TILE_OP] = tile_opcode(T.rows[r]);
    }
    __syncthreads();
    // runs of lens-edge rows that can be culled together (same element frame)
    for (int r = threadIdx.x; r < S; r += blockDim.x) {
        edge_run_at(T.rows, S, xf, r);
        // one control word per row for the pair kernel's loop: opcode | frame-change kind << 8 | cull run << 16
        xf[r].ctl = T.rows[r].i[DI_TILE_OP] | (xf[r].kind << 8) | (xf[r].run << 16);
    }
    if (threadIdx.x == 0) xf[S].ctl = -1;                              // ends the tile kernel's walk (no row count in its loop)
    __syncthreads();
}

// Reference-order walk of one ray (rtt_core.cuh arithmetic): irregular rays of a tile launch.  Out of line and
// free of pointer arguments into shared memory: it re-derives the table views from the dynamic shared-memory
// base itself, so the hot loop's table accesses stay provably shared-space (LDS, not generic loads).
struct WalkState { V3 p, d; float I; unsigned long long mask; };
__device__ __noinline__ WalkState seq_walk_generic(const SeqFwdArgs& a, int lam, long long i, WalkState w) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int S = a.tab.S, L = a.tab.L;
    SmemTable T = carve(smem_raw + kTileOffTable, S, L);
    ImgCache cache = img_cache_carve(smem_raw);
    unsigned long long bit = 1ull;
    for (int r = 0; r < S; ++r, bit += bit) seq_row<KDyn>(T, S, L, r, lam, i, a, cache, w.p, w.d, w.I, w.mask, bit);
    return w;
}

template <class K, int RPT, int BLK = kThreads, class M = unsigned long long>
__device__ __forceinline__ void tile_row(const SmemTable& T, int S, int L, int r, const SeqFwdArgs& a, ImgCache cache,
                                         long long i0, V3 (&p)[RPT], V3 (&d)[RPT], float (&I)[RPT],
                                         M (&mask)[RPT], const int (&lam)[RPT],
                                         const bool (&act)[RPT], M bit) {
    const RowDev& R = T.rows[r];
    float t[RPT];
    bool hit[RPT];
#pragma unroll
    for (int j = 0; j < RPT; ++j) hit[j] = tile_test<K>(T.rows, r, p[j], d[j], t[j]) && act[j];
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
        if (hit[j]) {
            Ior io;
            io.ni = io.no = 1.0f; io.mu_enter = io.mu_exit = 0.0f;
            if (uses_ior<K>(R)) io = row_ior(T, S, L, r, lam[j]);
            V3 np, nd, hl; float mod;
            tile_interact<K>(R, p[j], d[j], t[j], io.mu_enter, io.mu_exit, np, nd, mod, hl,
                             make_aux<K>(T.rows, R, io.ni, io.no, i0 + (long long)j * BLK, r, 0));
            if (K::sensor(R)) {
                const int slot = R.i[RTT_I_SENSOR];
                if (slot >= 0 && slot < a.n_sens)
                    sensor_deposit(a.sens[slot], cache, slot, i0 + (long long)j * BLK, hl, I[j], lam[j]);
            }
            p[j] = np; d[j] = nd; I[j] = I[j] * mod;
            mask[j] |= bit;
        }
    }
}

// BLK = threads per block; SYNC: the warps of a block start every tile together (one barrier per iteration of the block's
// grid-stride loop, whose trip count is block-uniform).  With 1024-thread blocks the eight warps of an SM sub-partition
// then fetch the same instructions at the same time: the build for rays generated in the kernel waits on instruction
// fetch for 40 % of its stall samples (profiles/r2_forward_kernels.md, c4cam).
// NARROW: tables of at most 32 rows keep the hit mask and the row bit in ONE register each (the walk updates them at
// every row and every hit; the 64-bit words of the general build cost two instructions each time and three more registers
// of the 64 a 1024-thread block has).  The launcher picks it by a.tab.S; results are the same bits.
// LUT: -1 = the wavelength table is tested at run time (L > 0) at every use — every refracting hit, every ray load;
// 0 / 1 = the launcher has looked: no table / a table, and the tests fold at compile time.
template <int RPT, int MINB, bool GEN, int BLK = kThreads, bool SYNC = false, bool NARROW = false, int LUT = -1>
__global__ void __launch_bounds__(BLK, MINB) k_trace_seq_fwd_tile(const __grid_constant__ SeqFwdArgs a) {
    typedef typename std::conditional<NARROW, unsigned, unsigned long long>::type mask_t;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int S = a.tab.S, L = (LUT == 0) ? 0 : a.tab.L;
    if (LUT == 1) __builtin_assume(L > 0);
    SmemTable T = carve(smem_raw + kTileOffTable, S, L);
    Xf* xf = reinterpret_cast<Xf*>(smem_raw + kTileOffXf);
    ImgCache cache = img_cache_carve(smem_raw);
    img_cache_init(cache);
    stage_table(a.tab, T);
    stage_tile(T, S, xf);
    SourceKey skey; skey.key = 0ull; skey.base = 0ull;
    if (GEN) skey = source_key(a.src);
    const long long tile = (long long)BLK * RPT;
    for (long long base = (long long)blockIdx.x * tile; base < a.n; base += (long long)gridDim.x * tile) {
        if (SYNC) __syncthreads();
        const long long i0 = base + threadIdx.x;
        V3 p[RPT], d[RPT];
        float I[RPT];
        mask_t mask[RPT];
        int lam[RPT];
        bool act[RPT], odd[RPT];
        // Rays in memory: the loads of ALL of this thread's rays are issued before the first value is consumed (one DRAM
        // round trip per tile instead of RPT: written as "load ray, look up its wavelength, load the next ray" the compiler
        // keeps that order, and the warps of a persistent block reach the top of a tile together, so nobody hides it).
        // A ray beyond the bundle reads the last ray's data (a valid address, no branch) and is switched off below.
        RayIn rin[RPT];
        if (!GEN) {
#pragma unroll
            for (int j = 0; j < RPT; ++j) {
                const long long i = i0 + (long long)j * BLK;
                rin[j] = fetch_ray_t<false>(a, skey, i < a.n ? i : a.n - 1, L > 0);
            }
        }
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
            const long long i = i0 + (long long)j * BLK;
            p[j] = v3(0.0f, 0.0f, 0.0f); d[j] = v3(0.0f, 0.0f, 0.0f); I[j] = 0.0f; lam[j] = 0; mask[j] = 0;
            act[j] = false; odd[j] = false;
            if (i < a.n) {
                const RayIn ray = GEN ? fetch_ray_t<true>(a, skey, i, L > 0) : rin[j];
                p[j] = ray.p; d[j] = ray.d; I[j] = ray.I;
                lam[j] = (L > 0) ? wavelength_index(T, L, ray.wav) : 0;
                act[j] = finite_ray(ray.p, ray.d) && regular_dir(ray.d);
                odd[j] = !act[j];                                       // re-read below: un-normalised or non-finite ray
            }
        }
        // The walk reads ONE control word per row (warp-uniform: opcode | frame-change kind << 8 | cull run << 16), fetched a
        // row ahead so that its shared-memory latency runs under the previous row's arithmetic; xf[S].ctl = -1 ends it.
        mask_t bit = 1;
        int next = xf[0].ctl;
        for (int r = 0; next >= 0; ++r, bit += bit) {
            const int ctl = next;
            if (ctl & 0xff00) {
#pragma unroll
                for (int j = 0; j < RPT; ++j) apply_xf_kind(xf[r], ctl & 0x0200, p[j], d[j]);
            }
            if (ctl >= 0x10000) {                                       // lens-edge rows: skip them when no lane can hit
                bool away = true;
#pragma unroll
                for (int j = 0; j < RPT; ++j) away = away & (!act[j] | edge_culled(xf[r], p[j], d[j]));
                if (__all_sync(kFull, away)) {
                    const int skip = (ctl >> 16) - 1;
                    r += skip; bit <<= skip;
                    next = xf[r + 1].ctl;
                    continue;
                }
            }
            next = xf[r + 1].ctl;
            switch (ctl & 0xff) {                                       // warp-uniform
#define RTT_X(OP, SURF, BOUND, SHAPE, PHYS, RS_IDENT, SENSOR)                                         \
                case OP: tile_row<KTile<SURF, BOUND, SHAPE, PHYS, RS_IDENT, SENSOR>, RPT, BLK, mask_t>(T, S, L, r, a, cache, i0, p, d, I, mask, lam, act, bit); break;
                RTT_TILE_SPECS(RTT_X)
#undef RTT_X
                default: tile_row<KDyn, RPT, BLK, mask_t>(T, S, L, r, a, cache, i0, p, d, I, mask, lam, act, bit); break;
            }
        }
        if (xf[S].kind) {
#pragma unroll
            for (int j = 0; j < RPT; ++j) apply_xf(xf[S], p[j], d[j]);
        }
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
            const long long i = i0 + (long long)j * BLK;
            if (odd[j]) {
                // un-normalised direction: walk in the reference's order; NaN / inf ray: hits nothing, stays as it was
                const RayIn ray = fetch_ray_t<GEN>(a, skey, i, L > 0);
                WalkState w;
                w.p = ray.p; w.d = ray.d; w.I = ray.I; w.mask = 0ull;
                if (finite_ray(ray.p, ray.d)) w = seq_walk_generic(a, lam[j], i, w);
                p[j] = w.p; d[j] = w.d; I[j] = w.I; mask[j] = (mask_t)w.mask;
            }
            if (i < a.n) {
                if (a.opos) { store3(a.opos, i, p[j]); store3(a.odir, i, d[j]); a.ointen[i] = I[j]; }
                if (a.hitmask) a.hitmask[i] = (unsigned long long)mask[j];
            }
        }
    }
    img_cache_flush(cache, a.sens);
}

// ============================================================================================
// Sequential trace, forward, FAST variant: packed ray pairs + bulk-async ray streaming
// ============================================================================================
// k_trace_seq_fwd_pair is the tile kernel's successor for Blackwell:
//   * ARITHMETIC: the two rays of a thread are the two lanes of f32x2 operands (rtt_pair.cuh: FFMA2 / FMUL2 / FADD2):
//     the FMA-pipe share of the per-row work issues once for both rays;
//   * RAY I/O (STREAM = true): no thread touches global memory for rays.  A block's tile of 512 rays is a handful of
//     CONTIGUOUS spans of the caller's AoS arrays (pos 6 KB, dir 6 KB, intensity 2 KB, wavelength 2 KB), so one elected
//     thread moves them with 1-D bulk-async copies (cp.async.bulk = the TMA engine, SASS UBLKCP) into a two-slot
//     shared-memory ring, completion signalled on an mbarrier (SYNCS); the threads read their rays from shared memory
//     (stride-3 words: bank-conflict free), write the results back into the same slot, and the slot leaves as bulk
//     stores.  The load of tile k+2 is issued when tile k has left its slot, i.e. a whole tile of arithmetic (~10 us)
//     ahead of its use: the ray loads never sit on a dependency chain, and the address arithmetic of 17 strided
//     LDG / STG per ray (the tile kernel's: 119 of its 137 global accesses) is gone from the issue stream.
//     One __syncthreads per tile (results complete -> elected thread stores).
// Rays generated in the kernel (rtt_source_t), unaligned caller buffers and the ragged last tile take the same
// arithmetic with plain loads / stores (STREAM = false build, or the cooperative copy below).
// NP = ray pairs per thread.  Shipped: 1 (512-ray tiles, 80 registers, 3 blocks / SM).  NP = 2 (1024-ray tiles, two
// independent packed chains per thread, 112 registers, 2 blocks / SM) measured slower on every workload — C2 4.28 vs
// 3.92, C1 2.75 vs 2.51, C4 8.19 vs 6.92 ms per 1e8 rays (profiles/r2_fwd_pair_ab.md) — and is not instantiated.
template <int NP, int BLK = kThreads>
struct PairGeom {
    static constexpr int kTile = NP * 2 * BLK;                     // rays per block iteration
    static constexpr int kPos = 0, kDir = kTile * 12, kInt = kTile * 24, kWav = kTile * 28, kMask = kTile * 32,
                         kBytes = kTile * 40;                           // 20 KB per slot and pair
};
constexpr int kPairSlots = 2;

// ---- bulk-async copy / mbarrier primitives (PTX ISA 8.x, sm_90+; SASS: UBLKCP, SYNCS) --------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded wait: a copy that never lands (a bug) must end the kernel with an error, not hang the GPU
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned done = 0;
    for (unsigned spin = 0; !done; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (!done && spin > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void bulk_g2s(unsigned dst_smem, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, unsigned src_smem, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy writes to shared memory -> visible to the async proxy (bulk stores read the slot)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// shared-memory layout of the pair kernel, all at compile-time offsets: [image cache | Xf[MAX_ROWS+1] | ray slots | mbarriers | table]
template <int LOG, bool STREAM, int NP = 1, int BLK = kThreads>
struct PairLayout {
    static constexpr size_t kOffXf = (img_cache_bytes<LOG>() + 15) / 16 * 16;
    static constexpr size_t kOffSlots = (kOffXf + sizeof(Xf) * (RTT_MAX_ROWS + 1) + 127) / 128 * 128;
    static constexpr size_t kOffBar = kOffSlots + (STREAM ? (size_t)kPairSlots * PairGeom<NP, BLK>::kBytes : 0);
    static constexpr size_t kOffTable = kOffBar + 16;
    __host__ __device__ static size_t bytes(int S, int L) { return kOffTable + smem_table_bytes(S, L); }
};

// reference-order walk of an irregular ray (see seq_walk_generic), for the pair kernel's layout
template <int LOG, bool STREAM, int NP, int BLK>
__device__ __noinline__ WalkState pair_walk_generic(const SeqFwdArgs& a, int lam, long long i, WalkState w) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int S = a.tab.S, L = a.tab.L;
    SmemTable T = carve(smem_raw + PairLayout<LOG, STREAM, NP, BLK>::kOffTable, S, L);
    ImgCacheT<LOG> cache = img_cache_carve<LOG>(smem_raw);
    unsigned long long bit = 1ull;
    for (int r = 0; r < S; ++r, bit += bit) seq_row<KDyn, LOG>(T, S, L, r, lam, i, a, cache, w.p, w.d, w.I, w.mask, bit);
    return w;
}

// ray `loc` of the tile staged in a slot
template <int NP, int BLK>
__device__ __forceinline__ RayIn slot_ray(const unsigned char* slot, int loc, bool want_wav) {
    typedef PairGeom<NP, BLK> GE;
    const float* sp = reinterpret_cast<const float*>(slot + GE::kPos);
    const float* sd = reinterpret_cast<const float*>(slot + GE::kDir);
    RayIn r;
    r.p = v3(sp[3 * loc], sp[3 * loc + 1], sp[3 * loc + 2]);
    r.d = v3(sd[3 * loc], sd[3 * loc + 1], sd[3 * loc + 2]);
    r.I = reinterpret_cast<const float*>(slot + GE::kInt)[loc];
    r.wav = want_wav ? reinterpret_cast<const float*>(slot + GE::kWav)[loc] : 0.0f;
    return r;
}

// sensor deposits of a pair: lane 0 / 1 = rays i0, i0 + kThreads
template <int LOG, int BLK = kThreads>
struct PairDeposit {
    const SeqFwdArgs& a;
    ImgCacheT<LOG> cache;
    long long i0;
    int lam_a, lam_b;
    __device__ __forceinline__ void operator()(int lane, int slot, V3 hl, float w) const {
        if (slot >= 0 && slot < a.n_sens)
            sensor_deposit(a.sens[slot], cache, slot, i0 + (long long)lane * BLK, hl, w, lane ? lam_b : lam_a);
    }
};

// per-lane refractive-index ratios of row r (row values, or the wavelength table's): one 8-byte shared-memory load per
// lane; lamS = wavelength index * S, computed once per ray
__device__ __forceinline__ void pair_ior(const SmemTable& T, int L, int r, int lamS_a, int lamS_b, F2& mu_enter, F2& mu_exit) {
    if (L > 0) {
        const float2 ma = T.mu[lamS_a + r], mb = T.mu[lamS_b + r];
        mu_enter = f2(ma.x, mb.x); mu_exit = f2(ma.y, mb.y);
    } else {
        const float2 m = *reinterpret_cast<const float2*>(T.rows[r].f + D_MU_ENTER);   // (D_MU_ENTER, D_MU_EXIT) adjacent, 8-byte aligned
        mu_enter = bc(m.x); mu_exit = bc(m.y);
    }
}

// a row kind without a packed form: the scalar twin per lane, with the specialised policy K
template <class K, int BLK, class DEP>
__device__ __forceinline__ unsigned pair_row_generic(const SmemTable& T, int S, int L, int r, long long i0, int lam_a, int lam_b,
                                                     P3& P, P3& D, F2& I, unsigned act, DEP& dep) {
    const RowDev& R = T.rows[r];
    F2 mu_enter = bc(0.0f), mu_exit = bc(0.0f);
    PhysAux aux_a = no_aux(), aux_b = no_aux();
    if (uses_ior<K>(R)) {
        pair_ior(T, L, r, lam_a * S, lam_b * S, mu_enter, mu_exit);
        if (K::phys(R) == RTT_PHYS_FRESNEL) {
            const Ior qa = row_ior(T, S, L, r, lam_a), qb = row_ior(T, S, L, r, lam_b);
            aux_a = make_aux<K>(T.rows, R, qa.ni, qa.no, i0, r, 0);
            aux_b = make_aux<K>(T.rows, R, qb.ni, qb.no, i0 + BLK, r, 0);
        }
    }
    return pair_row_scalar<K>(T.rows, r, P, D, I, act, mu_enter, mu_exit, aux_a, aux_b, dep);
}

template <int MINB, bool STREAM, int LOG, int NP, bool GEN, int BLK = kThreads>
__global__ void __launch_bounds__(BLK, MINB) k_trace_seq_fwd_pair(const __grid_constant__ SeqFwdArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typedef PairLayout<LOG, STREAM, NP, BLK> LY;
    typedef PairGeom<NP, BLK> GE;
    const int S = a.tab.S, L = a.tab.L;
    SmemTable T = carve(smem_raw + LY::kOffTable, S, L);
    Xf* xf = reinterpret_cast<Xf*>(smem_raw + LY::kOffXf);
    ImgCacheT<LOG> cache = img_cache_carve<LOG>(smem_raw);
    // tiles are counted in int: a.n <= 2^40 rays = 2^31 tiles of 512
    const int n_tiles = (int)((a.n + GE::kTile - 1) / GE::kTile);
    const int n_full = (int)(a.n / GE::kTile);                          // tiles [0, n_full) are full
    const unsigned bar0 = smem_u32(smem_raw + LY::kOffBar);
    const unsigned slot0 = smem_u32(smem_raw + LY::kOffSlots);
    const bool use_wav = L > 0;
    const unsigned tile_bytes = (unsigned)(GE::kTile * (use_wav ? 32 : 28));

    // elected thread: issue the bulk loads of tile `t` (a FULL tile) into slot `s`
    auto issue_load = [&](int t, int s) {
        const unsigned bar = bar0 + 8u * s, dst = slot0 + (unsigned)(s * GE::kBytes);
        const long long base = (long long)t * GE::kTile;
        mbar_expect_tx(bar, tile_bytes);
        bulk_g2s(dst + GE::kPos, a.pos + 3 * base, GE::kTile * 12, bar);
        bulk_g2s(dst + GE::kDir, a.dir + 3 * base, GE::kTile * 12, bar);
        bulk_g2s(dst + GE::kInt, a.inten + base, GE::kTile * 4, bar);
        if (use_wav) bulk_g2s(dst + GE::kWav, a.wav + base, GE::kTile * 4, bar);
    };

    if (STREAM && threadIdx.x == 0) {
        mbar_init(bar0, 1); mbar_init(bar0 + 8, 1);
        mbar_fence_init();
        for (int k = 0; k < kPairSlots; ++k) {                          // prologue: the first two tiles of this block
            const long long t = (long long)blockIdx.x + (long long)k * gridDim.x;
            if (t < n_full) issue_load((int)t, k);
        }
    }
    img_cache_init(cache);
    stage_table(a.tab, T);
    stage_tile(T, S, xf);
    SourceKey skey; skey.key = 0ull; skey.base = 0ull;
    if (GEN) skey = source_key(a.src);

    unsigned k = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++k) {
        const int s = (int)(k & 1u);
        unsigned char* slot = smem_raw + LY::kOffSlots + (size_t)s * GE::kBytes;
        const bool full = tile < n_full;
        const int cnt = full ? GE::kTile : (int)(a.n - (long long)tile * GE::kTile);
        if (STREAM) {
            if (full) {
                mbar_wait(bar0 + 8u * s, (k >> 1) & 1u);
            } else {                                                    // ragged last tile: cooperative plain copy
                const long long base = (long long)tile * GE::kTile;
                float* sp = reinterpret_cast<float*>(slot + GE::kPos);
                float* sd = reinterpret_cast<float*>(slot + GE::kDir);
                float* si = reinterpret_cast<float*>(slot + GE::kInt);
                float* sw = reinterpret_cast<float*>(slot + GE::kWav);
                for (int idx = threadIdx.x; idx < 3 * cnt; idx += BLK) {
                    sp[idx] = a.pos[3 * base + idx]; sd[idx] = a.dir[3 * base + idx];
                }
                for (int idx = threadIdx.x; idx < cnt; idx += BLK) {
                    si[idx] = a.inten[base + idx];
                    if (use_wav) sw[idx] = a.wav[base + idx];
                }
                __syncthreads();
            }
        }
        // ---- this thread's pairs: pair q = rays base + threadIdx.x + (2q) * BLK and + (2q + 1) * BLK ----
        P3 P[NP], D[NP];
        F2 I[NP];
        int lamS[NP][2];                                                // wavelength index * S (offset into the index table)
        unsigned actb[NP], oddb = 0u;
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            V3 pin[2], din[2];
            float Iin[2];
            actb[q] = 0u;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int loc = threadIdx.x + (2 * q + j) * BLK;
                pin[j] = v3(0.0f, 0.0f, 0.0f); din[j] = v3(0.0f, 0.0f, 0.0f); Iin[j] = 0.0f; lamS[q][j] = 0;
                if (loc < cnt) {
                    const RayIn ray = STREAM ? slot_ray<NP, BLK>(slot, loc, use_wav)
                                             : fetch_ray_t<GEN>(a, skey, (long long)tile * GE::kTile + loc, use_wav);
                    pin[j] = ray.p; din[j] = ray.d; Iin[j] = ray.I;
                    lamS[q][j] = use_wav ? wavelength_index(T, L, ray.wav) * S : 0;
                    if (finite_ray(ray.p, ray.d) && regular_dir(ray.d)) actb[q] |= 1u << j;
                    else oddb |= 1u << (2 * q + j);                     // re-read below: un-normalised or non-finite ray
                }
            }
            P[q] = pack3(pin[0], pin[1]); D[q] = pack3(din[0], din[1]); I[q] = f2(Iin[0], Iin[1]);
        }
        // hit masks: 32-bit words (row r -> bit r & 31 of the current word); the low words are set aside when the walk
        // passes row 32.  A 64-bit mask per lane costs four extra integer instructions per row.
        unsigned m_a[NP], m_b[NP], lo_a[NP], lo_b[NP];
#pragma unroll
        for (int q = 0; q < NP; ++q) { m_a[q] = m_b[q] = lo_a[q] = lo_b[q] = 0u; }
        bool upper = false;
        const long long i0 = (long long)tile * GE::kTile + threadIdx.x;
        for (int r = 0; r < S; ++r) {
            if (r >= 32 && !upper) {
#pragma unroll
                for (int q = 0; q < NP; ++q) { lo_a[q] = m_a[q]; lo_b[q] = m_b[q]; m_a[q] = 0u; m_b[q] = 0u; }
                upper = true;
            }
            const int ctl = xf[r].ctl;                                  // warp-uniform: opcode | kind << 8 | run << 16
            if (ctl & 0xff00) {
#pragma unroll
                for (int q = 0; q < NP; ++q) pair_apply_xf(xf[r], P[q], D[q]);
            }
            const int run = ctl >> 16;
            if (run > 0) {                                              // lens-edge rows: skip them when no lane can hit
                bool away = true;
#pragma unroll
                for (int q = 0; q < NP; ++q) away = away && pair_edge_culled(xf[r], P[q], D[q], actb[q] & 1u, actb[q] & 2u);
                if (__all_sync(kFull, away)) {
                    r += run - 1;
                    continue;
                }
            }
            unsigned hit[NP];
            F2 me[NP], mx[NP];
            switch (ctl & 0xff) {                                       // warp-uniform
                case 1:
#pragma unroll
                    for (int q = 0; q < NP; ++q) pair_ior(T, L, r, lamS[q][0], lamS[q][1], me[q], mx[q]);
#pragma unroll
                    for (int q = 0; q < NP; ++q) {
                        PairDeposit<LOG, BLK> dep{a, cache, i0 + (long long)(2 * q) * BLK, 0, 0};
                        hit[q] = pair_conic_face<true, RTT_SHAPE_SPHERIC_FACE>(T.rows, r, P[q], D[q], I[q], actb[q], me[q], mx[q], dep);
                    }
                    break;
                case 4:
#pragma unroll
                    for (int q = 0; q < NP; ++q) pair_ior(T, L, r, lamS[q][0], lamS[q][1], me[q], mx[q]);
#pragma unroll
                    for (int q = 0; q < NP; ++q) {
                        PairDeposit<LOG, BLK> dep{a, cache, i0 + (long long)(2 * q) * BLK, 0, 0};
                        hit[q] = pair_conic_face<false, RTT_SHAPE_CYL_FACE>(T.rows, r, P[q], D[q], I[q], actb[q], me[q], mx[q], dep);
                    }
                    break;
#define RTT_PLANE_CASE(OP, BOUND, PHYS, SENSOR)                                                                         \
                case OP:                                                                                                    \
                    _Pragma("unroll") for (int q = 0; q < NP; ++q) {                                                        \
                        PairDeposit<LOG, BLK> dep{a, cache, i0 + (long long)(2 * q) * BLK,                                  \
                                             use_wav ? lamS[q][0] / S : 0, use_wav ? lamS[q][1] / S : 0};                   \
                        hit[q] = pair_plane<BOUND, PHYS, SENSOR>(T.rows, r, P[q], D[q], I[q], actb[q], dep);                \
                    }                                                                                                       \
                    break;
                RTT_PLANE_CASE(7, RTT_BOUND_DISK, RTT_PHYS_APERTURE, false)
                RTT_PLANE_CASE(8, RTT_BOUND_DISK, RTT_PHYS_TRANSMIT, true)
                RTT_PLANE_CASE(9, RTT_BOUND_RECT, RTT_PHYS_TRANSMIT, true)
#undef RTT_PLANE_CASE
#define RTT_X(OP, SURF, BOUND, SHAPE, PHYS, RS_IDENT, SENSOR)                                                           \
                case OP:                                                                                                    \
                    _Pragma("unroll") for (int q = 0; q < NP; ++q) {                                                        \
                        const int la = use_wav ? lamS[q][0] / S : 0, lb = use_wav ? lamS[q][1] / S : 0;                     \
                        PairDeposit<LOG, BLK> dep{a, cache, i0 + (long long)(2 * q) * BLK, la, lb};                         \
                        hit[q] = pair_row_generic<KTile<SURF, BOUND, SHAPE, PHYS, RS_IDENT, SENSOR>, BLK>(                       \
                            T, S, L, r, i0 + (long long)(2 * q) * BLK, la, lb, P[q], D[q], I[q], actb[q], dep);        \
                    }                                                                                                       \
                    break;
                RTT_PAIR_SCALAR_SPECS(RTT_X)
#undef RTT_X
                default:
#pragma unroll
                    for (int q = 0; q < NP; ++q) {
                        const int la = use_wav ? lamS[q][0] / S : 0, lb = use_wav ? lamS[q][1] / S : 0;
                        PairDeposit<LOG, BLK> dep{a, cache, i0 + (long long)(2 * q) * BLK, la, lb};
                        hit[q] = pair_row_generic<KDyn, BLK>(T, S, L, r, i0 + (long long)(2 * q) * BLK, la, lb, P[q], D[q], I[q],
                                                        actb[q], dep);
                    }
                    break;
            }
            const unsigned bit = 1u << (r & 31);
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                m_a[q] |= bit & (0u - (hit[q] & 1u));
                m_b[q] |= bit & (0u - (hit[q] >> 1));
            }
        }
        if (xf[S].kind) {
#pragma unroll
            for (int q = 0; q < NP; ++q) pair_apply_xf(xf[S], P[q], D[q]);
        }
        if (!upper) {
#pragma unroll
            for (int q = 0; q < NP; ++q) { lo_a[q] = m_a[q]; lo_b[q] = m_b[q]; m_a[q] = 0u; m_b[q] = 0u; }
        }
        // ---- results ----
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            V3 po[2] = {lane_a(P[q]), lane_b(P[q])}, dout[2] = {lane_a(D[q]), lane_b(D[q])};
            float Io[2] = {I[q].x, I[q].y};
            unsigned long long mo[2] = {((unsigned long long)m_a[q] << 32) | lo_a[q], ((unsigned long long)m_b[q] << 32) | lo_b[q]};
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int loc = threadIdx.x + (2 * q + j) * BLK;
                if ((oddb >> (2 * q + j)) & 1u) {
                    // un-normalised direction: walk in the reference's order; NaN / inf ray: hits nothing, stays as it was
                    // (the slot still holds this ray's inputs: a thread only ever writes its own entries)
                    const long long i = (long long)tile * GE::kTile + loc;
                    const RayIn ray = STREAM ? slot_ray<NP, BLK>(slot, loc, use_wav) : fetch_ray_t<GEN>(a, skey, i, use_wav);
                    WalkState w;
                    w.p = ray.p; w.d = ray.d; w.I = ray.I; w.mask = 0ull;
                    if (finite_ray(ray.p, ray.d))
                        w = pair_walk_generic<LOG, STREAM, NP, BLK>(a, use_wav ? lamS[q][j] / S : 0, i, w);
                    po[j] = w.p; dout[j] = w.d; Io[j] = w.I; mo[j] = w.mask;
                }
                if (loc < cnt) {
                    if (STREAM) {
                        float* sp = reinterpret_cast<float*>(slot + GE::kPos);
                        float* sd = reinterpret_cast<float*>(slot + GE::kDir);
                        sp[3 * loc] = po[j].x; sp[3 * loc + 1] = po[j].y; sp[3 * loc + 2] = po[j].z;
                        sd[3 * loc] = dout[j].x; sd[3 * loc + 1] = dout[j].y; sd[3 * loc + 2] = dout[j].z;
                        reinterpret_cast<float*>(slot + GE::kInt)[loc] = Io[j];
                        reinterpret_cast<unsigned long long*>(slot + GE::kMask)[loc] = mo[j];
                    } else {
                        const long long i = (long long)tile * GE::kTile + loc;
                        if (a.opos) { store3(a.opos, i, po[j]); store3(a.odir, i, dout[j]); a.ointen[i] = Io[j]; }
                        if (a.hitmask) a.hitmask[i] = mo[j];
                    }
                }
            }
        }
        if (STREAM) {
            fence_async_smem();
            __syncthreads();                                            // every result of the tile is in the slot
            const long long base = (long long)tile * GE::kTile;
            if (full) {
                if (threadIdx.x == 0) {
                    const unsigned src = slot0 + (unsigned)(s * GE::kBytes);
                    if (a.opos) {
                        bulk_s2g(a.opos + 3 * base, src + GE::kPos, GE::kTile * 12);
                        bulk_s2g(a.odir + 3 * base, src + GE::kDir, GE::kTile * 12);
                        bulk_s2g(a.ointen + base, src + GE::kInt, GE::kTile * 4);
                    }
                    if (a.hitmask) bulk_s2g(a.hitmask + base, src + GE::kMask, GE::kTile * 8);
                    bulk_commit();
                    const long long next = (long long)tile + (long long)kPairSlots * gridDim.x;
                    if (next < n_tiles) {
                        bulk_wait_read0();                              // the stores have read the slot: refill it
                        if (next < n_full) issue_load((int)next, s);
                    }
                }
            } else {
                const float* sp = reinterpret_cast<const float*>(slot + GE::kPos);
                const float* sd = reinterpret_cast<const float*>(slot + GE::kDir);
                if (a.opos) {
                    for (int idx = threadIdx.x; idx < 3 * cnt; idx += BLK) {
                        a.opos[3 * base + idx] = sp[idx]; a.odir[3 * base + idx] = sd[idx];
                    }
                    for (int idx = threadIdx.x; idx < cnt; idx += BLK)
                        a.ointen[base + idx] = reinterpret_cast<const float*>(slot + GE::kInt)[idx];
                }
                if (a.hitmask)
                    for (int idx = threadIdx.x; idx < cnt; idx += BLK)
                        a.hitmask[base + idx] = reinterpret_cast<const unsigned long long*>(slot + GE::kMask)[idx];
            }
        }
    }
    if (STREAM && threadIdx.x == 0) bulk_wait_all0();                   // the slot must outlive the stores that read it
    img_cache_flush(cache, a.sens);
}

#endif  // RTT_APPROX

// ============================================================================================
// Parameter-gradient reduction helpers (adjoint kernels)
// ============================================================================================
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// Reduce entries [lo, hi) of G over the warp and add them to the block accumulator row.
__device__ __forceinline__ void reduce_span(const RowGrad& G, int lo, int hi, float* acc_row) {
    for (int e = lo; e < hi; ++e) {
        const float s = warp_sum(G.g[e]);
        if ((threadIdx.x & 31) == 0 && s != 0.0f) atomicAdd(acc_row + e, s);
    }
}

__device__ __forceinline__ void reduce_row_grad(const RowGrad& G, int flags, float* acc_row) {
    if (flags & RTT_FLAG_GRAD_POSE_E) reduce_span(G, RTT_F_RE, RTT_F_TE + 3, acc_row);
    if (flags & RTT_FLAG_GRAD_POSE_S) reduce_span(G, RTT_F_RS, RTT_F_TS + 3, acc_row);
    if (flags & RTT_FLAG_GRAD_CK) reduce_span(G, RTT_F_C, RTT_F_K + 1, acc_row);
    if (flags & RTT_FLAG_GRAD_RADIUS) reduce_span(G, RTT_F_RADIUS, RTT_F_RADIUS + 1, acc_row);
    if (flags & RTT_FLAG_GRAD_IOR) reduce_span(G, RTT_F_IOR_IN, RTT_F_IOR_OUT + 1, acc_row);
}

// Per-lane variant for divergent rows (non-sequential adjoint): shared-memory atomics.
__device__ __forceinline__ void scatter_row_grad(const RowGrad& G, int flags, float* acc_row) {
    auto span = [&](int lo, int hi) {
        for (int e = lo; e < hi; ++e) if (G.g[e] != 0.0f) atomicAdd(acc_row + e, G.g[e]);
    };
    if (flags & RTT_FLAG_GRAD_POSE_E) span(RTT_F_RE, RTT_F_TE + 3);
    if (flags & RTT_FLAG_GRAD_POSE_S) span(RTT_F_RS, RTT_F_TS + 3);
    if (flags & RTT_FLAG_GRAD_CK) span(RTT_F_C, RTT_F_K + 1);
    if (flags & RTT_FLAG_GRAD_RADIUS) span(RTT_F_RADIUS, RTT_F_RADIUS + 1);
    if (flags & RTT_FLAG_GRAD_IOR) span(RTT_F_IOR_IN, RTT_F_IOR_OUT + 1);
}

__device__ __forceinline__ void flush_block_grads(const float* acc, int S, float* g_table,
                                                  const float* acc_lut, int L, float* g_lut) {
    __syncthreads();
    if (g_table)
        for (int idx = threadIdx.x; idx < S * RTT_ROW_G; idx += blockDim.x)
            if (acc[idx] != 0.0f) atomicAdd(g_table + idx, acc[idx]);
    if (g_lut)
        for (int idx = threadIdx.x; idx < L * S * 2; idx += blockDim.x)
            if (acc_lut[idx] != 0.0f) atomicAdd(g_lut + idx, acc_lut[idx]);
}

// ============================================================================================
// Sequential trace, adjoint
// ============================================================================================

struct Checkpoint { V3 p, d; };

// Replay of one recorded interaction (no validity tests: the hit mask says it happened).
template <class K>
__device__ __forceinline__ void replay_row(const SmemTable& T, int S, int L, int r, int lam, long long i, V3& p, V3& d) {
    const RowDev& R = T.rows[r];
    const Frames F = to_frames<K>(R, p, d);
    const Roots q = solve_roots<K>(R, F.o, F.dd);
    int which;
    const float t = select_root<K>(R, q, F.o, F.dd, &which);
    Ior io;
    io.ni = io.no = 1.0f; io.mu_enter = io.mu_exit = 0.0f;
    if (uses_ior<K>(R)) io = row_ior(T, S, L, r, lam);
    const Step s = interact<K>(R, F, t, p, d, io.mu_enter, io.mu_exit, make_aux<K>(T.rows, R, io.ni, io.no, i, r, 0));
    p = s.hit_global; d = s.new_dir;
}

// Reverse step through one recorded interaction.
template <class K>
__device__ __forceinline__ void reverse_row(const SmemTable& T, int S, int L, int r, int lam, long long i,
                                            const SeqBwdArgs& a, const Checkpoint& ck, V3& gp, V3& gd, float& gI,
                                            RowGrad& G, int flags) {
    const RowDev& R = T.rows[r];
    Ior io;
    io.ni = io.no = 1.0f; io.mu_enter = io.mu_exit = 1.0f;
    if (uses_ior<K>(R)) io = row_ior(T, S, L, r, lam);
    V3 g_hl = v3(0, 0, 0);
    float g_w = 0.0f;
    if (K::sensor(R)) {
        const int slot = R.i[RTT_I_SENSOR];
        if (slot >= 0 && slot < a.n_sens && a.g_record[slot]) {
            const float4 gr = reinterpret_cast<const float4*>(a.g_record[slot])[i];
            g_hl = v3(gr.x, gr.y, gr.z); g_w = gr.w;
        }
    }
    V3 ngp, ngd; float mod;
    interact_adjoint<K>(R, ck.p, ck.d, io.ni, io.no, io.mu_enter, io.mu_exit,
                        gp, gd, g_hl, v3(0, 0, 0), 0.0f, ngp, ngd, mod, G, flags,
                        make_aux<K>(T.rows, R, io.ni, io.no, i, r, 0).u);
    gp = ngp; gd = ngd;
    gI = gI * mod + g_w;
}

// The generic (KDyn) replay / reverse steps handle every row kind the table format can express and are ~9 k instructions
// of code between them, eight times the specialised lens-face + stop + sensor paths that the BASELINE scenes execute.
// Inlined into the row switches they pushed the kernel to 12.8 k instructions (200 KB) and the hot paths apart: 45 % of
// the adjoint's stall samples on C3 were `no_instruction` (instruction fetch).  Out of line, by value, they cost a call
// only when a row really needs them.
struct ReplayOut { V3 p, d; };
__device__ __noinline__ ReplayOut replay_row_generic(SmemTable T, int S, int L, int r, int lam, long long i, V3 p, V3 d) {
    replay_row<KDyn>(T, S, L, r, lam, i, p, d);
    ReplayOut o; o.p = p; o.d = d;
    return o;
}
struct ReverseOut { V3 gp, gd; float gI; RowGrad G; };
__device__ __noinline__ ReverseOut reverse_row_generic(SmemTable T, int S, int L, int r, int lam, long long i,
                                                       const SeqBwdArgs& a, Checkpoint ck, V3 gp, V3 gd, float gI, int flags) {
    ReverseOut o;
    zero(o.G);
    reverse_row<KDyn>(T, S, L, r, lam, i, a, ck, gp, gd, gI, o.G, flags);
    o.gp = gp; o.gd = gd; o.gI = gI;
    return o;
}

// Rays per block iteration of the sequential adjoint.  When no input-ray gradients are requested, a ray whose
// upstream gradients (final position / direction, sensor-record xyz) are all zero contributes nothing to any
// parameter gradient — every term of the reverse sweep is linear in them — so the block first compacts the
// chunk to the rays that matter (dead rays of an intensity-weighted loss, rays that missed everything) and
// runs the replay + reverse sweep on full warps of those.
constexpr int kAccRows = 12, kAccPerRow = 5;      // private gradient slots: rows x (c, k, radius, ior_in, ior_out)
constexpr int kBwdChunkIters = 16;         // chunk = 16 rays per thread: large enough that the compacted chunk still fills whole blocks of warps
__host__ __device__ inline size_t bwd_queue_bytes(int threads) { return sizeof(unsigned short) * kBwdChunkIters * threads + 16; }

// POSE = false: the caller guarantees that no row requests pose gradients (RTT_MODE_SCALAR_GRADS); the pose-gradient
// outer products and their 24 accumulator registers per row are compiled out.
// CK = checkpoints a thread can hold = rows of the table rounded up (24 or RTT_MAX_ROWS): a 64-entry frame reserved 1.9 KB of
// local memory per thread for scenes that never hit more than 17 rows.
// LEAN (FAST build, POSE = false, no input-ray gradients): rays whose every interaction is a lens face, a stop or a sensor
// take the frame-resident replay / reverse steps of rtt_lean.cuh; the rest of the queue runs the general code below.
// GEN: 0 = rays from memory or from the ray source (run-time test), 1 = memory only, 2 = generated only (see fetch_ray_t).
// BLK = threads per block (256, or one block of 1024 per SM: see the block-size A/B of the forward kernels).
template <int MINB, bool POSE, int CK, bool LEAN, int GEN, int BLK = kThreads>
__global__ void __launch_bounds__(BLK, MINB) RTT_NAME(k_trace_seq_bwd)(const __grid_constant__ SeqBwdArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int S = a.tab.S, L = a.tab.L;
    SmemTable T = carve(smem_raw, S, L);
    float* acc = reinterpret_cast<float*>(smem_raw + ((smem_table_bytes(S, L) + 15) / 16) * 16);
    float* acc_lut = acc + S * RTT_ROW_G;
    int* qcount = reinterpret_cast<int*>(acc_lut + L * S * 2 + ((L * S * 2 + S * RTT_ROW_G) & 1));   // 8-byte aligned
    unsigned short* queue = reinterpret_cast<unsigned short*>(qcount + 2);
    for (int idx = threadIdx.x; idx < S * RTT_ROW_G + L * S * 2; idx += blockDim.x) acc[idx] = 0.0f;
    stage_table(a.tab, T);
#if defined(RTT_APPROX)
    // lean path: frame changes between the rows (rtt_tile.cuh) behind the queue, then the mask of lean rows (two words)
    Xf* xf = reinterpret_cast<Xf*>(reinterpret_cast<unsigned char*>(queue) + bwd_queue_bytes(BLK));
    unsigned* lean_words = reinterpret_cast<unsigned*>(xf + (S + 1));
    if (LEAN) stage_tile(T, S, xf);
#endif

    // Private gradient accumulators: rows whose requested gradients are scalars only (c, k, radius, indices —
    // the usual lens-design variables) add into per-thread slots (local memory, L1) ray after ray and are reduced
    // over the warp ONCE at the end of the block, instead of 5 shuffles + a shared atomic per entry per ray row.
    // Rows with pose gradients, wavelength-resolved indices or beyond kAccRows keep the per-row warp reduction.
    if (threadIdx.x == 0) {
        int used = 0;
        for (int r = 0; r < S; ++r) {
            const int fl = T.rows[r].i[RTT_I_FLAGS];
            const bool scalar_only = fl != 0 && !(fl & (RTT_FLAG_GRAD_POSE_E | RTT_FLAG_GRAD_POSE_S)) &&
                                     !(L > 0 && (fl & RTT_FLAG_GRAD_IOR));
            T.rows[r].f[D_ACC_SLOT] = (a.g_table && scalar_only && used < kAccRows) ? (float)(used++) : -1.0f;
        }
#if defined(RTT_APPROX)
        if (LEAN) {
            // a row is lean when it has a lean step (rtt_lean.cuh) and its requested gradients — pose requests are
            // ignored in this build — all go to private slots
            unsigned long long lean = 0ull;
            for (int r = 0; r < S; ++r) {
                const int fl = a.g_table ? (T.rows[r].i[RTT_I_FLAGS] & ~(RTT_FLAG_GRAD_POSE_E | RTT_FLAG_GRAD_POSE_S)) : 0;
                if (lean_tile_op(T.rows[r].i[DI_TILE_OP]) && (fl == 0 || T.rows[r].f[D_ACC_SLOT] >= 0.0f)) lean |= 1ull << r;
            }
            lean_words[0] = (unsigned)lean; lean_words[1] = (unsigned)(lean >> 32);
            // the lean walk visits only the rows that matter to a lean ray: lean rows and frame changes.
            // word = row | tile opcode << 8 | frame-change kind << 16 | lean << 24
            int nw = 0;
            for (int r = 0; r < S; ++r) {
                const unsigned is_lean = (unsigned)((lean >> r) & 1ull);
                if (is_lean || xf[r].kind != 0)
                    lean_words[3 + nw++] = (unsigned)r | ((unsigned)T.rows[r].i[DI_TILE_OP] << 8) | ((unsigned)xf[r].kind << 16) | (is_lean << 24);
            }
            lean_words[2] = (unsigned)nw;
        }
#endif
    }
    __syncthreads();
#if defined(RTT_APPROX)
    const unsigned long long lean_rows = LEAN ? (((unsigned long long)lean_words[1] << 32) | lean_words[0]) : 0ull;
#endif
    float pacc[kAccRows * kAccPerRow];
#pragma unroll
    for (int e = 0; e < kAccRows * kAccPerRow; ++e) pacc[e] = 0.0f;

    const bool compact = !a.g_pos && !a.g_dir && !a.g_inten;
    SourceKey skey; skey.key = 0ull; skey.base = 0ull;
    if (GEN == 2 || (GEN == 0 && a.src.kind >= 0)) skey = source_key(a.src);
    const int chunk = a.chunk;
    for (long long base = (long long)blockIdx.x * chunk; base < a.n; base += (long long)gridDim.x * chunk) {
      int count = (int)((a.n - base) < (long long)chunk ? (a.n - base) : (long long)chunk);
      if (compact) {
        if (threadIdx.x == 0) *qcount = 0;
        __syncthreads();
        // pass 1: which of this thread's rays matter.  Four rays at a time, every load of the group issued before the
        // first is consumed: the hit mask does not gate the gradient loads (a dependent second round trip to DRAM per
        // ray, with the whole block waiting at the barrier below, was a quarter of this kernel's warp time).
        unsigned needs = 0u;
        constexpr int kGroup = 4;
#pragma unroll
        for (int k0 = 0; k0 < kBwdChunkIters; k0 += kGroup) {
            unsigned long long hm[kGroup];
            V3 ga[kGroup], gb[kGroup];
            float4 gr[kGroup];
            bool ok[kGroup];
#pragma unroll
            for (int j = 0; j < kGroup; ++j) {
                const int loc = (k0 + j) * BLK + threadIdx.x;
                ok[j] = loc < count;
                const long long i = base + (ok[j] ? loc : 0);
                hm[j] = a.hitmask[i];
                ga[j] = a.g_opos ? load3(a.g_opos, i) : v3(0.0f, 0.0f, 0.0f);
                gb[j] = a.g_odir ? load3(a.g_odir, i) : v3(0.0f, 0.0f, 0.0f);
                gr[j] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                if (a.n_sens > 0 && a.g_record[0]) gr[j] = reinterpret_cast<const float4*>(a.g_record[0])[i];
            }
#pragma unroll
            for (int j = 0; j < kGroup; ++j) {
                bool need = ga[j].x != 0.0f || ga[j].y != 0.0f || ga[j].z != 0.0f ||
                            gb[j].x != 0.0f || gb[j].y != 0.0f || gb[j].z != 0.0f ||
                            gr[j].x != 0.0f || gr[j].y != 0.0f || gr[j].z != 0.0f;
                if (ok[j] && hm[j] != 0ull && !need) {                   // further sensors (rare): dependent loads
                    const long long i = base + (k0 + j) * BLK + threadIdx.x;
                    for (int sl = 1; sl < a.n_sens; ++sl)
                        if (a.g_record[sl]) {
                            const float4 g2 = reinterpret_cast<const float4*>(a.g_record[sl])[i];
                            need = need || g2.x != 0.0f || g2.y != 0.0f || g2.z != 0.0f;
                        }
                }
                need = need && ok[j] && hm[j] != 0ull;
                needs |= (need ? 1u : 0u) << (k0 + j);
            }
        }
        // pass 2: append them to the block's queue (warp-aggregated)
        for (int k = 0; k < chunk / BLK; ++k) {
            const bool need = (needs >> k) & 1u;
            const unsigned votes = __ballot_sync(kFull, need);
            if (votes) {
                const int lane = threadIdx.x & 31, lead = __ffs((int)votes) - 1;
                int pos = 0;
                if (lane == lead) pos = atomicAdd(qcount, __popc(votes));
                pos = __shfl_sync(kFull, pos, lead);
                if (need) queue[pos + __popc(votes & ((1u << lane) - 1u))] = (unsigned short)(k * BLK + threadIdx.x);
            }
        }
        __syncthreads();
        count = *qcount;
      }
      for (int q0 = 0; q0 < count; q0 += BLK) {
        const int q = q0 + threadIdx.x;
        const bool live = q < count;
        const long long i = base + (live ? (compact ? (int)queue[q] : q) : 0);
        unsigned long long mask = 0ull;
        V3 p = v3(0, 0, 0), d = v3(0, 0, 1);
        int lam = 0;
        if (live) {
            mask = a.hitmask[i];
            const RayIn ray = GEN == 0 ? fetch_ray(a, skey, i, L > 0) : fetch_ray_t<GEN == 2>(a, skey, i, L > 0);
            p = ray.p; d = ray.d;
            if (L > 0) lam = wavelength_index(T, L, ray.wav);
        }
        Checkpoint ck[CK];
        int nh = 0;
#if defined(RTT_APPROX)
        if (LEAN) {
            // ---- lean rays: frame-resident replay + reverse (rtt_lean.cuh); 7 words per interaction in the same frame ----
            constexpr int kLeanHits = CK * 6 / kLeanCkWords;
            const bool lean_ray = live && mask != 0ull && (mask & ~lean_rows) == 0ull && __popcll(mask) <= kLeanHits &&
                                  finite_ray(p, d) && regular_dir(d);
            if (__any_sync(kFull, lean_ray)) {
                float* ckw = reinterpret_cast<float*>(ck);
                const unsigned m_lo = lean_ray ? (unsigned)mask : 0u, m_hi = lean_ray ? (unsigned)(mask >> 32) : 0u;
                const int nwalk = (int)lean_words[2];
                const int lamS = lam * S;
                V3 lp = p, ld = d;
                for (int k = 0; k < nwalk; ++k) {
                    const unsigned w = lean_words[3 + k];                // warp-uniform
                    const int r = (int)(w & 0xffu);
                    if (w & 0xff0000u) apply_xf(xf[r], lp, ld);
                    if (((((r & 32) ? m_hi : m_lo) >> (r & 31)) & 1u) && (w >> 24)) {
                        const RowDev& R = T.rows[r];
                        const int op = (int)((w >> 8) & 0xffu);
                        LeanCk c;
                        if (op == 1 || op == 4) {
                            const float2 m = (L > 0) ? T.mu[lamS + r] : *reinterpret_cast<const float2*>(R.f + D_MU_ENTER);
                            if (op == 1) lean_face_replay<true>(R, m.x, m.y, lp, ld, c);
                            else lean_face_replay<false>(R, m.x, m.y, lp, ld, c);
                        } else {
                            lean_plane_replay(R, op, lp, ld, c);
                        }
                        lean_ck_store(ckw + kLeanCkWords * nh, c);
                        ++nh;
                    }
                }
                V3 gp = v3(0, 0, 0), gd = v3(0, 0, 0);
                if (lean_ray) {
                    if (a.g_opos) gp = load3(a.g_opos, i);
                    if (a.g_odir) gd = load3(a.g_odir, i);
                }
                lean_xf_transpose(xf[S], gp, gd);                        // global frame -> frame of the last row
                for (int k = nwalk - 1; k >= 0; --k) {
                    const unsigned w = lean_words[3 + k];
                    const int r = (int)(w & 0xffu);
                    if (((((r & 32) ? m_hi : m_lo) >> (r & 31)) & 1u) && (w >> 24)) {
                        --nh;
                        const LeanCk c = lean_ck_load(ckw + kLeanCkWords * nh);
                        const RowDev& R = T.rows[r];
                        const int op = (int)((w >> 8) & 0xffu);
                        if (op == 1 || op == 4) {
                            const int flags = a.g_table ? R.i[RTT_I_FLAGS] : 0;
                            const int slot = (int)R.f[D_ACC_SLOT];
                            // the private slots are read before the arithmetic and written after it: the local-memory
                            // round trip hides behind the reverse step
                            float* pa = pacc + (slot >= 0 ? slot : 0) * kAccPerRow;
                            float g5[kAccPerRow] = {pa[0], pa[1], 0.0f, pa[3], pa[4]};
                            float mu_enter, mu_exit, ni, no;
                            if (L > 0) {
                                const float2 m = T.mu[lamS + r];
                                mu_enter = m.x; mu_exit = m.y; ni = T.lut_ni[lamS + r]; no = T.lut_no[lamS + r];
                            } else {
                                mu_enter = R.f[D_MU_ENTER]; mu_exit = R.f[D_MU_EXIT]; ni = R.f[RTT_F_IOR_IN]; no = R.f[RTT_F_IOR_OUT];
                            }
                            if (op == 1) lean_face_reverse<true>(R, mu_enter, mu_exit, ni, no, c, gp, gd, flags, g5);
                            else lean_face_reverse<false>(R, mu_enter, mu_exit, ni, no, c, gp, gd, flags, g5);
                            if (slot >= 0) {
                                if (flags & RTT_FLAG_GRAD_CK) { pa[0] = g5[0]; pa[1] = g5[1]; }
                                if (flags & RTT_FLAG_GRAD_IOR) { pa[3] = g5[3]; pa[4] = g5[4]; }
                            }
                        } else {
                            V3 g_hl = v3(0, 0, 0);
                            if (op != 7) {
                                const int slot = R.i[RTT_I_SENSOR];
                                if (slot >= 0 && slot < a.n_sens && a.g_record[slot]) {
                                    const float4 gr = reinterpret_cast<const float4*>(a.g_record[slot])[i];
                                    g_hl = v3(gr.x, gr.y, gr.z);
                                }
                            }
                            lean_plane_reverse(R, op, c, g_hl, gp, gd);
                        }
                    }
                    if (((w >> 16) & 0xffu) == 2u) lean_xf_transpose(xf[r], gp, gd);   // frame of row r -> frame of row r - 1
                }
                nh = 0;
            }
            if (lean_ray) mask = 0ull;                                   // done: the general loops below skip this lane
            if (!__any_sync(kFull, mask != 0ull)) continue;
        }
#endif
        // ---- forward replay over the recorded interactions (row loop is warp-uniform) ----
        for (int r = 0; r < S; ++r) {
            const bool hit = (mask >> r) & 1ull;
            if (__ballot_sync(kFull, hit) == 0u) continue;
            const int op = T.rows[r].i[DI_OPCODE];
            if (hit) {
                ck[nh].p = p; ck[nh].d = d; ++nh;
                switch (op) {
#define RTT_X(OP, SURF, BOUND, SHAPE, PHYS, IDENT, SENSOR)                                            \
                    case OP: replay_row<KStatic<SURF, BOUND, SHAPE, PHYS, IDENT, SENSOR>>(T, S, L, r, lam, i, p, d); break;
                    RTT_ROW_SPECS_ADJ(RTT_X)
#undef RTT_X
                    default: { const ReplayOut o = replay_row_generic(T, S, L, r, lam, i, p, d); p = o.p; d = o.d; break; }
                }
            }
        }
        // ---- reverse sweep ----
        V3 gp = v3(0, 0, 0), gd = v3(0, 0, 0);
        float gI = 0.0f;
        if (live) {
            if (a.g_opos) gp = load3(a.g_opos, i);
            if (a.g_odir) gd = load3(a.g_odir, i);
            if (a.g_ointen) gI = a.g_ointen[i];
        }
        for (int r = S - 1; r >= 0; --r) {
            const bool hit = (mask >> r) & 1ull;
            if (__ballot_sync(kFull, hit) == 0u) continue;
            const RowDev& R = T.rows[r];
            const int flags = POSE ? R.i[RTT_I_FLAGS]
                                   : (R.i[RTT_I_FLAGS] & ~(RTT_FLAG_GRAD_POSE_E | RTT_FLAG_GRAD_POSE_S));
            const int op = R.i[DI_OPCODE];
            RowGrad G;
            zero(G);
            if (hit) {
                --nh;
                switch (op) {
#define RTT_X(OP, SURF, BOUND, SHAPE, PHYS, IDENT, SENSOR)                                            \
                    case OP: reverse_row<KStatic<SURF, BOUND, SHAPE, PHYS, IDENT, SENSOR>>(T, S, L, r, lam, i, a, ck[nh], gp, gd, gI, G, flags); break;
                    RTT_ROW_SPECS_ADJ(RTT_X)
#undef RTT_X
                    default: {
                        const ReverseOut o = reverse_row_generic(T, S, L, r, lam, i, a, ck[nh], gp, gd, gI, flags);
                        gp = o.gp; gd = o.gd; gI = o.gI; G = o.G;
                        break;
                    }
                }
            }
            const int slot = (int)R.f[D_ACC_SLOT];
            if (a.g_table && flags && slot >= 0) {
                if (hit) {                                               // G is zero where nothing was requested
                    float* pa = pacc + slot * kAccPerRow;                // only the requested entries: each is a
                    if (flags & RTT_FLAG_GRAD_CK) { pa[0] += G.g[RTT_F_C]; pa[1] += G.g[RTT_F_K]; }   // local-memory
                    if (flags & RTT_FLAG_GRAD_RADIUS) pa[2] += G.g[RTT_F_RADIUS];                      // read-modify-write
                    if (flags & RTT_FLAG_GRAD_IOR) { pa[3] += G.g[RTT_F_IOR_IN]; pa[4] += G.g[RTT_F_IOR_OUT]; }
                }
            } else if (a.g_table && flags) {
                if (L > 0 && (flags & RTT_FLAG_GRAD_IOR)) {
                    // wavelength-resolved index gradients go to the LUT accumulator
                    for (int l = 0; l < L; ++l) {
                        const float s_in = warp_sum((hit && lam == l) ? G.g[RTT_F_IOR_IN] : 0.0f);
                        const float s_out = warp_sum((hit && lam == l) ? G.g[RTT_F_IOR_OUT] : 0.0f);
                        if ((threadIdx.x & 31) == 0) {
                            if (s_in != 0.0f) atomicAdd(acc_lut + ((size_t)l * S + r) * 2, s_in);
                            if (s_out != 0.0f) atomicAdd(acc_lut + ((size_t)l * S + r) * 2 + 1, s_out);
                        }
                    }
                    reduce_row_grad(G, flags & ~RTT_FLAG_GRAD_IOR, acc + r * RTT_ROW_G);
                } else {
                    reduce_row_grad(G, flags, acc + r * RTT_ROW_G);
                }
            }
        }
        if (live) {
            if (a.g_pos) store3(a.g_pos, i, gp);
            if (a.g_dir) store3(a.g_dir, i, gd);
            if (a.g_inten) a.g_inten[i] = gI;
        }
      }
      if (compact) __syncthreads();                                       // the queue is rewritten by the next chunk
    }
    // private slots -> block accumulator: one warp reduction per entry for the whole block's work
    for (int r = 0; r < S; ++r) {
        const int slot = (int)T.rows[r].f[D_ACC_SLOT];
        if (slot < 0) continue;
        const int entry[kAccPerRow] = {RTT_F_C, RTT_F_K, RTT_F_RADIUS, RTT_F_IOR_IN, RTT_F_IOR_OUT};
#pragma unroll
        for (int e = 0; e < kAccPerRow; ++e) {
            const float sum = warp_sum(pacc[slot * kAccPerRow + e]);
            if ((threadIdx.x & 31) == 0 && sum != 0.0f) atomicAdd(acc + r * RTT_ROW_G + entry[e], sum);
        }
    }
    flush_block_grads(acc, S, a.g_table, acc_lut, L, a.g_lut);
}

// ============================================================================================
// Non-sequential trace, forward (scene/base.py:129-235)
// ============================================================================================

// One row of the nearest-hit search (scene/base.py:164-176).
template <class K>
__device__ __forceinline__ void nonseq_probe(const RowDev* rows, int r, V3 p, V3 d, float& best, int& win,
                                             bool& poisoned, Frames& F) {
    Roots q; float t; int which;
    const bool finite_t = intersect_t<K>(rows, r, p, d, F, q, t, which, rows[r].f[D_SAME_ELEM] != 0.0f);
    // rows of a Shape report inf when invalid; bare surfaces may report NaN, which poisons torch.min
    if (K::shape(rows[r]) == RTT_SHAPE_NONE && is_nan(t)) poisoned = true;
    if (finite_t && t < best && shape_ok<K>(rows, r, F, t)) { best = t; win = r; }
}

__global__ void __launch_bounds__(kThreads) RTT_NAME(k_trace_nonseq_fwd)(const __grid_constant__ NonseqFwdArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmemTable T = carve(smem_raw, a.tab.S, a.tab.L);
    unsigned char* after = smem_raw + ((smem_table_bytes(a.tab.S, a.tab.L) + 15) / 16) * 16;
    ImgCache cache = img_cache_carve(after);
    NsCull* cull = reinterpret_cast<NsCull*>(after + ((img_cache_bytes() + 15) / 16) * 16);
    img_cache_init(cache);
    stage_table(a.tab, T);
    const int S = a.tab.S, L = a.tab.L, NB = a.nbounces;
    for (int r = threadIdx.x; r < S; r += blockDim.x) {
        cull[r] = box_cull_info(T.rows, S, r);
        bool same = r > 0 && T.rows[r].i[RTT_I_SHAPE] != RTT_SHAPE_NONE && T.rows[r - 1].i[RTT_I_SHAPE] != RTT_SHAPE_NONE;
        for (int e = RTT_F_RE; same && e < RTT_F_TE + 3; ++e) same = T.rows[r].f[e] == T.rows[r - 1].f[e];
        T.rows[r].f[D_SAME_ELEM] = same ? 1.0f : 0.0f;
    }
    __syncthreads();
    // verified boxes: the element-frame bounding sphere goes into scalar slots a box face (plane, no surface bound)
    // does not use, where shape_in_bounds finds it (rows are this block's private copy)
    for (int r = threadIdx.x; r < S; r += blockDim.x) {
        if (cull[r].run != 6) continue;
        for (int f = 0; f < 6; ++f) {
            RowDev& B = T.rows[r + f];
            B.f[RTT_F_C] = cull[r].ex; B.f[RTT_F_K] = cull[r].ey; B.f[RTT_F_RADIUS] = cull[r].ez;
            B.f[D_SB0SQ] = cull[r].r2;
        }
    }
    __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x;
    const SourceKey skey = fetch_key(a);
    // Lane refill: rays of a warp need different numbers of bounces (absorbed, escaped, still bouncing), so a lane
    // whose ray has ended fetches its next ray at once instead of idling until the slowest ray of the warp is done.
    // One trip of the loop = one bounce for every lane that holds a ray.
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool have = false;
    V3 p = v3(0, 0, 0), d = v3(0, 0, 0);
    float I = 0.0f;
    int lam = 0, nb = 0;
    unsigned cnts = 0u;                                                 // 8 bits per sensor slot
    while (true) {
        if (!have && i < a.n) {
            const RayIn ray = fetch_ray(a, skey, i, L > 0);
            p = ray.p; d = ray.d; I = ray.I;
            lam = (L > 0) ? wavelength_index(T, L, ray.wav) : 0;
            cnts = 0u; nb = 0;
            have = true;
        }
        if (!__any_sync(kFull, have)) break;
        if (!have) continue;
        bool done = (nb >= NB) || !(I > 0.0f) || !finite_ray(p, d);     // base.py:140,201; NaN / inf rays hit nothing
        if (!done) {
            // ray_cast (base.py:164-176): min over all rows, NaN anywhere => no hit
            float best = rtt_inf();
            int win = -1;
            bool poisoned = false;
            Frames Fs;                                                  // element-frame part shared by the rows of an element
            Fs.pe = Fs.de = Fs.den = Fs.o = Fs.dd = v3(0.0f, 0.0f, 0.0f); Fs.len = 0.0f;
            for (int r = 0; r < S; ++r) {
                if (cull[r].run > 0) {                                  // a box: skip its six faces when no lane can hit it
                    const bool missed = sphere_missed(cull[r], p, d);
                    if (__all_sync(__activemask(), missed)) { r += cull[r].run - 1; continue; }
                }
                // min over rows (base.py:169): a row can only become the winner with t < best, so the
                // shape-level rule (the costly part for box faces and lens edges) is evaluated only then
                switch (T.rows[r].i[DI_OPCODE]) {                       // warp-uniform
#define RTT_X(OP, SURF, BOUND, SHAPE, PHYS, IDENT, SENSOR)                                            \
                    case OP: nonseq_probe<KStatic<SURF, BOUND, SHAPE, PHYS, IDENT, SENSOR>>(T.rows, r, p, d, best, win, poisoned, Fs); break;
                    RTT_ROW_SPECS(RTT_X)
#undef RTT_X
                    default: nonseq_probe<KDyn>(T.rows, r, p, d, best, win, poisoned, Fs); break;
                }
            }
            if (poisoned || win < 0) {
                done = true;
            } else {
                Frames F; Roots q; float t; int which;
                intersect<false>(T.rows, win, p, d, F, q, t, which);
                const RowDev& R = T.rows[win];
                const Ior io = row_ior(T, S, L, win, lam);
                const Step s = interact(R, F, t, p, d, io.mu_enter, io.mu_exit,
                                        make_aux(T.rows, R, io.ni, io.no, i, win, nb));
                const int slot = R.i[RTT_I_SENSOR];
                if (slot >= 0 && slot < a.n_sens) {
                    const unsigned c = (cnts >> (8 * slot)) & 255u;
                    sensor_deposit(a.sens[slot], cache, slot, i, s.hit_local, I, lam, (int)c, a.n);
                    if (c < 255u) cnts += 1u << (8 * slot);
                }
                p = s.hit_global; d = s.new_dir; I = I * s.mod;
                if (a.hit_seq) a.hit_seq[i * NB + nb] = (unsigned char)win;
                ++nb;
                // retire in THIS trip when the ray is over (bounce limit, absorbed, blown up): an absorbed ray used to
                // hold its lane for one more, idle trip of the warp loop before it was written back and replaced
                done = (nb >= NB) || !(I > 0.0f) || !finite_ray(p, d);
            }
        }
        if (done) {
            for (int s = 0; s < a.n_sens; ++s)
                if (a.sens[s].count) a.sens[s].count[i] = (unsigned char)((cnts >> (8 * s)) & 255u);
            // (the tail of hit_seq is pre-filled with 255 by the launcher: one memset instead of byte stores per ray)
            if (a.n_hits) a.n_hits[i] = (unsigned char)nb;
            if (a.opos) { store3(a.opos, i, p); store3(a.odir, i, d); a.ointen[i] = I; }
            have = false;
            i += stride;
        }
    }
    img_cache_flush(cache, a.sens);
}

// Lock-step build of the non-sequential forward (A/B: RTT_MODE_TUNE 1..5 of rtt_trace_nonseq_fwd).  The kernel above is
// bound by instruction FETCH (ncu: `no_instruction` is half of its stall samples): the row loop of one search runs through
// ~6 specialised bodies of several hundred instructions each, far more than an SM sub-partition's L0 instruction cache
// holds, and the resident warps sit in different bodies, so every warp streams its code from the L1.5 cache by itself.
// Here the warps of a block start every trip (SYNC >= 1) — or every row of the search (SYNC == 2) — together, so the
// warps of a sub-partition fetch the same lines at the same time.  Same per-ray arithmetic and results.
template <int BLOCK, int SYNC>
__global__ void __launch_bounds__(BLOCK, 1024 / BLOCK) RTT_NAME(k_trace_nonseq_fwd_ls)(const __grid_constant__ NonseqFwdArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmemTable T = carve(smem_raw, a.tab.S, a.tab.L);
    unsigned char* after = smem_raw + ((smem_table_bytes(a.tab.S, a.tab.L) + 15) / 16) * 16;
    ImgCache cache = img_cache_carve(after);
    NsCull* cull = reinterpret_cast<NsCull*>(after + ((img_cache_bytes() + 15) / 16) * 16);
    img_cache_init(cache);
    stage_table(a.tab, T);
    const int S = a.tab.S, L = a.tab.L, NB = a.nbounces;
    for (int r = threadIdx.x; r < S; r += blockDim.x) {
        cull[r] = box_cull_info(T.rows, S, r);
        bool same = r > 0 && T.rows[r].i[RTT_I_SHAPE] != RTT_SHAPE_NONE && T.rows[r - 1].i[RTT_I_SHAPE] != RTT_SHAPE_NONE;
        for (int e = RTT_F_RE; same && e < RTT_F_TE + 3; ++e) same = T.rows[r].f[e] == T.rows[r - 1].f[e];
        T.rows[r].f[D_SAME_ELEM] = same ? 1.0f : 0.0f;
    }
    __syncthreads();
    for (int r = threadIdx.x; r < S; r += blockDim.x) {
        if (cull[r].run != 6) continue;
        for (int f = 0; f < 6; ++f) {
            RowDev& B = T.rows[r + f];
            B.f[RTT_F_C] = cull[r].ex; B.f[RTT_F_K] = cull[r].ey; B.f[RTT_F_RADIUS] = cull[r].ez;
            B.f[D_SB0SQ] = cull[r].r2;
        }
    }
    __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x;
    const SourceKey skey = fetch_key(a);
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool have = false;
    V3 p = v3(0, 0, 0), d = v3(0, 0, 0);
    float I = 0.0f;
    int lam = 0, nb = 0;
    unsigned cnts = 0u;
    while (true) {
        if (!have && i < a.n) {
            const RayIn ray = fetch_ray(a, skey, i, L > 0);
            p = ray.p; d = ray.d; I = ray.I;
            lam = (L > 0) ? wavelength_index(T, L, ray.wav) : 0;
            cnts = 0u; nb = 0;
            have = true;
        }
        if (SYNC == 0) { if (!__any_sync(kFull, have)) break; }         // (A/B: the block size alone, warps run free)
        else if (!__syncthreads_or(have ? 1 : 0)) break;                // every warp of the block starts the trip together
        bool done = !have || (nb >= NB) || !(I > 0.0f) || !finite_ray(p, d);
        float best = rtt_inf();
        int win = -1;
        bool poisoned = false;
        Frames Fs;
        Fs.pe = Fs.de = Fs.den = Fs.o = Fs.dd = v3(0.0f, 0.0f, 0.0f); Fs.len = 0.0f;
        int skip_to = 0;
        for (int r = 0; r < S; ++r) {
            if (SYNC == 2 || (SYNC == 3 && (r & 3) == 0 && r > 0) || (SYNC == 5 && r * 2 == (S & ~1))) __syncthreads();
            if (r < skip_to) continue;
            if (cull[r].run > 0) {
                const bool missed = done || sphere_missed(cull[r], p, d);
                if (__all_sync(kFull, missed)) { skip_to = r + cull[r].run; continue; }
            }
            if (!done) {
                switch (T.rows[r].i[DI_OPCODE]) {                       // warp-uniform
#define RTT_X(OP, SURF, BOUND, SHAPE, PHYS, IDENT, SENSOR)                                            \
                    case OP: nonseq_probe<KStatic<SURF, BOUND, SHAPE, PHYS, IDENT, SENSOR>>(T.rows, r, p, d, best, win, poisoned, Fs); break;
                    RTT_ROW_SPECS(RTT_X)
#undef RTT_X
                    default: nonseq_probe<KDyn>(T.rows, r, p, d, best, win, poisoned, Fs); break;
                }
            }
        }
        if (SYNC == 4 || SYNC == 5) __syncthreads();
        if (!done) {
            if (poisoned || win < 0) {
                done = true;
            } else {
                Frames F; Roots q; float t; int which;
                intersect<false>(T.rows, win, p, d, F, q, t, which);
                const RowDev& R = T.rows[win];
                const Ior io = row_ior(T, S, L, win, lam);
                const Step s = interact(R, F, t, p, d, io.mu_enter, io.mu_exit,
                                        make_aux(T.rows, R, io.ni, io.no, i, win, nb));
                const int slot = R.i[RTT_I_SENSOR];
                if (slot >= 0 && slot < a.n_sens) {
                    const unsigned c = (cnts >> (8 * slot)) & 255u;
                    sensor_deposit(a.sens[slot], cache, slot, i, s.hit_local, I, lam, (int)c, a.n);
                    if (c < 255u) cnts += 1u << (8 * slot);
                }
                p = s.hit_global; d = s.new_dir; I = I * s.mod;
                if (a.hit_seq) a.hit_seq[i * NB + nb] = (unsigned char)win;
                ++nb;
                done = (nb >= NB) || !(I > 0.0f) || !finite_ray(p, d);
            }
        }
        if (have && done) {
            for (int s = 0; s < a.n_sens; ++s)
                if (a.sens[s].count) a.sens[s].count[i] = (unsigned char)((cnts >> (8 * s)) & 255u);
            if (a.n_hits) a.n_hits[i] = (unsigned char)nb;
            if (a.opos) { store3(a.opos, i, p); store3(a.odir, i, d); a.ointen[i] = I; }
            have = false;
            i += stride;
        }
    }
    img_cache_flush(cache, a.sens);
}

// ============================================================================================
// Non-sequential trace, adjoint: replay the recorded hit sequence, then reverse
// ============================================================================================

constexpr int kMaxReplay = 32;   // checkpoints held per ray at a time; deeper hit sequences are differentiated in windows

__global__ void __launch_bounds__(kThreads) RTT_NAME(k_trace_nonseq_bwd)(const __grid_constant__ NonseqBwdArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int S = a.tab.S, L = a.tab.L, NB = a.nbounces;
    SmemTable T = carve(smem_raw, S, L);
    float* acc = reinterpret_cast<float*>(smem_raw + ((smem_table_bytes(S, L) + 15) / 16) * 16);
    float* acc_lut = acc + S * RTT_ROW_G;
    for (int idx = threadIdx.x; idx < S * RTT_ROW_G + L * S * 2; idx += blockDim.x) acc[idx] = 0.0f;
    stage_table(a.tab, T);
    const long long stride = (long long)gridDim.x * blockDim.x;
    const SourceKey skey = fetch_key(a);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
        const RayIn ray = fetch_ray(a, skey, i, L > 0);
        const int lam = (L > 0) ? wavelength_index(T, L, ray.wav) : 0;
        int H = 0;                                                      // interactions of this ray (base.py:201: <= Nbounces)
        while (H < NB && a.hit_seq[i * NB + H] != 255) ++H;
        V3 gp = a.g_opos ? load3(a.g_opos, i) : v3(0, 0, 0);
        V3 gd = a.g_odir ? load3(a.g_odir, i) : v3(0, 0, 0);
        float gI = a.g_ointen ? a.g_ointen[i] : 0.0f;
        Checkpoint ck[kMaxReplay];
        unsigned char rows_hit[kMaxReplay];
        // Windows of kMaxReplay interactions, last window first.  Each window replays the ray from its ORIGINAL state
        // up to the window's end (recompute, not store) and keeps the incoming states of the window's interactions only;
        // a ray with H <= kMaxReplay hits (every BASELINE config) is one window = one replay.
        for (int w0 = (H > 0 ? ((H - 1) / kMaxReplay) * kMaxReplay : 0); w0 >= 0 && H > 0; w0 -= kMaxReplay) {
            const int wend = (w0 + kMaxReplay < H) ? w0 + kMaxReplay : H;
            V3 p = ray.p, d = ray.d;
            int nh = 0;
            unsigned cnts = 0u;                                         // sensor interactions per slot, 8 bits each
            for (int b = 0; b < wend; ++b) {
                const int r = a.hit_seq[i * NB + b];
                if (b >= w0) { ck[nh].p = p; ck[nh].d = d; rows_hit[nh] = (unsigned char)r; ++nh; }
                const RowDev& R = T.rows[r];
                {
                    const int sl = R.i[RTT_I_SENSOR];
                    if (sl >= 0 && sl < a.n_sens && ((cnts >> (8 * sl)) & 255u) < 255u) cnts += 1u << (8 * sl);
                }
                const Frames F = to_frames(R, p, d);
                const Roots q = solve_roots(R, F.o, F.dd);
                int which;
                const float t = select_root(R, q, F.o, F.dd, &which);
                const Ior io = row_ior(T, S, L, r, lam);
                const Step s = interact(R, F, t, p, d, io.mu_enter, io.mu_exit, make_aux(T.rows, R, io.ni, io.no, i, r, b));
                p = s.hit_global; d = s.new_dir;
            }
            while (nh > 0) {
                --nh;
                const int r = rows_hit[nh];
                const RowDev& R = T.rows[r];
                const int flags = R.i[RTT_I_FLAGS];
                const Ior io = row_ior(T, S, L, r, lam);
                RowGrad G;
                zero(G);
                V3 g_hl = v3(0, 0, 0);
                float g_w = 0.0f;
                const int slot = R.i[RTT_I_SENSOR];
                if (slot >= 0 && slot < a.n_sens) {
                    cnts -= 1u << (8 * slot);                           // ordinal of this interaction
                    const int ord = (int)((cnts >> (8 * slot)) & 255u);
                    if (a.g_record[slot] && ord < a.rec_hits[slot]) {
                        const float4 gr = reinterpret_cast<const float4*>(a.g_record[slot])[(long long)ord * a.n + i];
                        g_hl = v3(gr.x, gr.y, gr.z); g_w = gr.w;
                    }
                }
                V3 ngp, ngd; float mod;
                interact_adjoint(R, ck[nh].p, ck[nh].d, io.ni, io.no, io.mu_enter, io.mu_exit,
                                 gp, gd, g_hl, v3(0, 0, 0), 0.0f, ngp, ngd, mod, G, flags,
                                 make_aux(T.rows, R, io.ni, io.no, i, r, w0 + nh).u);   // w0 + nh == bounce index
                gp = ngp; gd = ngd; gI = gI * mod + g_w;
                if (a.g_table && flags) {
                    if (L > 0 && (flags & RTT_FLAG_GRAD_IOR)) {
                        atomicAdd(acc_lut + ((size_t)lam * S + r) * 2, G.g[RTT_F_IOR_IN]);
                        atomicAdd(acc_lut + ((size_t)lam * S + r) * 2 + 1, G.g[RTT_F_IOR_OUT]);
                        scatter_row_grad(G, flags & ~RTT_FLAG_GRAD_IOR, acc + r * RTT_ROW_G);
                    } else {
                        scatter_row_grad(G, flags, acc + r * RTT_ROW_G);
                    }
                }
            }
        }
        if (a.g_pos) store3(a.g_pos, i, gp);
        if (a.g_dir) store3(a.g_dir, i, gd);
        if (a.g_inten) a.g_inten[i] = gI;
    }
    flush_block_grads(acc, S, a.g_table, acc_lut, L, a.g_lut);
}

// ============================================================================================
// Element.intersectTest (elements/parent.py:30-42): [n, k] distances
// ============================================================================================

__global__ void __launch_bounds__(kThreads) RTT_NAME(k_intersect_test)(const __grid_constant__ IsectArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmemTable T = carve(smem_raw, a.tab.S, 0);
    TableDev tb = a.tab; tb.L = 0;
    stage_table(tb, T);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
        const V3 p = load3(a.pos, i), d = load3(a.dir, i);
        for (int j = 0; j < a.k; ++j) {
            Frames F; Roots q; float t; int which;
            const bool valid = intersect<true>(T.rows, a.row0 + j, p, d, F, q, t, which);
            // Shape rows mask invalid hits to inf (shape.py:55); bare surfaces return t as is
            const bool bare = T.rows[a.row0 + j].i[RTT_I_SHAPE] == RTT_SHAPE_NONE;
            a.t_out[i * a.k + j] = bare ? t : (valid ? t : rtt_inf());
        }
    }
}

// ============================================================================================
// Element.forward on one row (elements/parent.py:44-58), forward and adjoint
// ============================================================================================

__global__ void __launch_bounds__(kThreads) RTT_NAME(k_surface_step_fwd)(const __grid_constant__ StepFwdArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmemTable T = carve(smem_raw, a.tab.S, a.tab.L);
    stage_table(a.tab, T);
    const int S = a.tab.S, L = a.tab.L, r = a.row;
    const RowDev& R = T.rows[r];
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
        const V3 p = load3(a.pos, i), d = load3(a.dir, i);
        const int lam = (L > 0) ? wavelength_index(T, L, a.wav[i]) : 0;
        const Frames F = to_frames(R, p, d);
        const Roots q = solve_roots(R, F.o, F.dd);
        int which;
        const float t = select_root(R, q, F.o, F.dd, &which);
        const Ior io = row_ior(T, S, L, r, lam);
        const Step s = interact(R, F, t, p, d, io.mu_enter, io.mu_exit, make_aux(T.rows, R, io.ni, io.no, i, r, 0));
        store3(a.npos, i, s.hit_global); store3(a.ndir, i, s.new_dir);
        a.mod[i] = s.mod;
        if (a.hit_local) store3(a.hit_local, i, s.hit_local);
        if (a.t_out) a.t_out[i] = t;
        if (a.normal) store3(a.normal, i, s.normal);
    }
}


__global__ void __launch_bounds__(kThreads) RTT_NAME(k_surface_step_bwd)(const __grid_constant__ StepBwdArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int S = a.tab.S, L = a.tab.L, r = a.row;
    SmemTable T = carve(smem_raw, S, L);
    float* acc = reinterpret_cast<float*>(smem_raw + ((smem_table_bytes(S, L) + 15) / 16) * 16);
    float* acc_lut = acc + S * RTT_ROW_G;
    for (int idx = threadIdx.x; idx < S * RTT_ROW_G + L * S * 2; idx += blockDim.x) acc[idx] = 0.0f;
    stage_table(a.tab, T);
    const RowDev& R = T.rows[r];
    const int flags = R.i[RTT_I_FLAGS];
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long base = (long long)blockIdx.x * blockDim.x; base < a.n; base += stride) {
        const long long i = base + threadIdx.x;
        const bool live = i < a.n;
        RowGrad G;
        zero(G);
        int lam = 0;
        if (live) {
            const V3 p = load3(a.pos, i), d = load3(a.dir, i);
            if (L > 0) lam = wavelength_index(T, L, a.wav[i]);
            const Ior io = row_ior(T, S, L, r, lam);
            const V3 gnp = a.g_npos ? load3(a.g_npos, i) : v3(0, 0, 0);
            const V3 gnd = a.g_ndir ? load3(a.g_ndir, i) : v3(0, 0, 0);
            const V3 ghl = a.g_hit_local ? load3(a.g_hit_local, i) : v3(0, 0, 0);
            const V3 gnn = a.g_normal ? load3(a.g_normal, i) : v3(0, 0, 0);
            const float gtt = a.g_t ? a.g_t[i] : 0.0f;
            V3 gp, gd; float mod;
            interact_adjoint(R, p, d, io.ni, io.no, io.mu_enter, io.mu_exit, gnp, gnd, ghl, gnn, gtt,
                             gp, gd, mod, G, flags, make_aux(T.rows, R, io.ni, io.no, i, r, 0).u);
            if (a.g_pos) store3(a.g_pos, i, gp);
            if (a.g_dir) store3(a.g_dir, i, gd);
        }
        if (a.g_table && flags) {
            if (L > 0 && (flags & RTT_FLAG_GRAD_IOR)) {
                for (int l = 0; l < L; ++l) {
                    const float s_in = warp_sum((live && lam == l) ? G.g[RTT_F_IOR_IN] : 0.0f);
                    const float s_out = warp_sum((live && lam == l) ? G.g[RTT_F_IOR_OUT] : 0.0f);
                    if ((threadIdx.x & 31) == 0) {
                        if (s_in != 0.0f) atomicAdd(acc_lut + ((size_t)l * S + r) * 2, s_in);
                        if (s_out != 0.0f) atomicAdd(acc_lut + ((size_t)l * S + r) * 2 + 1, s_out);
                    }
                }
                reduce_row_grad(G, flags & ~RTT_FLAG_GRAD_IOR, acc + r * RTT_ROW_G);
            } else {
                reduce_row_grad(G, flags, acc + r * RTT_ROW_G);
            }
        }
    }
    flush_block_grads(acc, S, a.g_table, acc_lut, L, a.g_lut);
}


// ============================================================================================
// Renderer.render_3d (render/camera.py:191-257): nearest hit + normal + shading, one launch
// ============================================================================================
// The reference builds the [N, S] distance matrix of the renderable elements, takes the row-wise minimum, then — per
// winning (element, surface) — gathers the pixel rays, re-runs the surface's forward for the normal and shades
// (0.3 ambient + 0.7 |n . light|) x base colour (render/camera.py:259-301).  Here every pixel ray does its search,
// recomputes the winner's geometry and writes its colour in the same thread; the pinhole camera's rays can be
// generated in the kernel (rtt_source_t, kind CAMERA), so a render reads no ray input at all.
__global__ void __launch_bounds__(kThreads) RTT_NAME(k_render_shade)(const __grid_constant__ RenderArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int S = a.tab.S;
    SmemTable T = carve(smem_raw, S, 0);
    TableDev tb = a.tab; tb.L = 0;
    stage_table(tb, T);
    for (int r = threadIdx.x; r < S; r += blockDim.x) {                 // rows of one element share the element-frame ray
        bool same = r > 0 && T.rows[r].i[RTT_I_SHAPE] != RTT_SHAPE_NONE && T.rows[r - 1].i[RTT_I_SHAPE] != RTT_SHAPE_NONE;
        for (int e = RTT_F_RE; same && e < RTT_F_TE + 3; ++e) same = T.rows[r].f[e] == T.rows[r - 1].f[e];
        T.rows[r].f[D_SAME_ELEM] = same ? 1.0f : 0.0f;
    }
    __syncthreads();
    const V3 light = v3(a.light[0], a.light[1], a.light[2]);
    const long long stride = (long long)gridDim.x * blockDim.x;
    const SourceKey skey = fetch_key(a);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
        V3 p, d;
        if (a.src.kind >= 0) source_ray(a.src, skey, i, p, d);
        else { p = load3(a.pos, i); d = load3(a.dir, i); }
        float best = rtt_inf();
        int win = -1;
        bool poisoned = !finite_ray(p, d);
        Frames Fs;
        Fs.pe = Fs.de = Fs.den = Fs.o = Fs.dd = v3(0.0f, 0.0f, 0.0f); Fs.len = 0.0f;
        for (int r = 0; r < S && !poisoned; ++r) {                      // min over rows, NaN anywhere => no hit (:236-237)
            switch (T.rows[r].i[DI_OPCODE]) {                           // warp-uniform
#define RTT_X(OP, SURF, BOUND, SHAPE, PHYS, IDENT, SENSOR)                                            \
                case OP: nonseq_probe<KStatic<SURF, BOUND, SHAPE, PHYS, IDENT, SENSOR>>(T.rows, r, p, d, best, win, poisoned, Fs); break;
                RTT_ROW_SPECS(RTT_X)
#undef RTT_X
                default: nonseq_probe<KDyn>(T.rows, r, p, d, best, win, poisoned, Fs); break;
            }
        }
        V3 rgb = v3(a.bg[0], a.bg[1], a.bg[2]);
        if (!poisoned && win >= 0) {
            Frames F; Roots q; float t; int which;
            intersect<false>(T.rows, win, p, d, F, q, t, which);        // Shape.forward of the winner (:249)
            const RowDev& R = T.rows[win];
            float nlen;
            const V3 n = normal_global(R, normal_local(R, along(F.o, t, F.dd), &nlen));
            const float shade = 0.3f + 0.7f * fabsf(dot(n, light));     // :296-299
            const float* b = a.base_rgb + 3 * win;
            rgb = v3(fminf(fmaxf(b[0] * shade, 0.0f), 1.0f), fminf(fmaxf(b[1] * shade, 0.0f), 1.0f),
                     fminf(fmaxf(b[2] * shade, 0.0f), 1.0f));
        } else {
            win = 255;
            rgb = v3(fminf(fmaxf(rgb.x, 0.0f), 1.0f), fminf(fmaxf(rgb.y, 0.0f), 1.0f), fminf(fmaxf(rgb.z, 0.0f), 1.0f));
        }
        store3(a.rgb, i, rgb);
        if (a.win) a.win[i] = (unsigned char)win;
    }
}

// ============================================================================================
// Bundle.sample (rays/bundle.py:30-37): materialise the rays of a source
// ============================================================================================
__global__ void __launch_bounds__(kThreads) RTT_NAME(k_sample_bundle)(const __grid_constant__ SampleArgs a) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const SourceKey skey = source_key(a.src);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
        V3 p, d;
        source_ray(a.src, skey, i, p, d);
        store3(a.pos, i, p); store3(a.dir, i, d);
        a.inten[i] = a.src.intensity;
        if (a.wav) a.wav[i] = a.src.wavelength;
    }
}

// ============================================================================================
// Host launchers
// ============================================================================================
inline int sm_count() {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    return sms;
}

// persistent blocks: a multiple of the SM count, never more blocks than there are ray tiles
inline int grid_for(long long n, int blocks_per_sm) {
    const long long tiles = (n + kThreads - 1) / kThreads;
    long long g = (long long)sm_count() * blocks_per_sm;
    if (g <= 0) g = 1;
    if (tiles < g) g = tiles;
    if (g < 1) g = 1;
    return (int)g;
}

inline size_t fwd_smem(int S, int L) { return ((smem_table_bytes(S, L) + 15) / 16) * 16 + img_cache_bytes(); }
inline size_t nonseq_fwd_smem(int S, int L) {
    return ((smem_table_bytes(S, L) + 15) / 16) * 16 + ((img_cache_bytes() + 15) / 16) * 16 + sizeof(NsCull) * (size_t)S;
}

inline size_t bwd_smem(int S, int L) {
    return ((smem_table_bytes(S, L) + 15) / 16) * 16 + sizeof(float) * ((size_t)S * RTT_ROW_G + (size_t)L * S * 2);
}

// the forward kernels carry the 32 KB image cache next to the table: opt in above the 48 KB default
template <class Kern>
inline cudaError_t allow_smem(Kern kern, size_t bytes) {
    if (bytes <= 48 * 1024) return cudaSuccess;
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

#if defined(RTT_APPROX)
template <int RPT, int MINB, bool GEN, int BLK = kThreads, bool SYNC = false, int kWaves = 4, bool NARROW = false, int LUT = -1>
inline cudaError_t launch_tile_g(const SeqFwdArgs& a, cudaStream_t st) {
    const size_t smem = tile_smem_bytes(a.tab.S, a.tab.L);
    if (cudaError_t e = allow_smem(k_trace_seq_fwd_tile<RPT, MINB, GEN, BLK, SYNC, NARROW, LUT>, smem)) return e;
    const long long tiles = (a.n + (long long)BLK * RPT - 1) / ((long long)BLK * RPT);
    // several waves of grid-striding blocks: a block that lands on a busier SM costs 1/kWaves of a launch
    long long g = (long long)sm_count() * MINB * kWaves;
    if (tiles < g) g = tiles;
    if (g < 1) g = 1;
    k_trace_seq_fwd_tile<RPT, MINB, GEN, BLK, SYNC, NARROW, LUT><<<(int)g, BLK, smem, st>>>(a);
    return cudaGetLastError();
}
template <int RPT, int MINB>
inline cudaError_t launch_tile(const SeqFwdArgs& a, cudaStream_t st) {
    return a.src.kind >= 0 ? launch_tile_g<RPT, MINB, true>(a, st) : launch_tile_g<RPT, MINB, false>(a, st);
}
// large blocks: RPT rays per thread, MINB blocks of BLK threads per SM, lock-step (SYNC) or free-running
template <int RPT, int MINB, int BLK, bool SYNC, int WAVES = 4, bool NARROW = false, int LUT = -1>
inline cudaError_t launch_tile_big(const SeqFwdArgs& a, cudaStream_t st) {
    return a.src.kind >= 0 ? launch_tile_g<RPT, MINB, true, BLK, SYNC, WAVES, NARROW, LUT>(a, st) : launch_tile_g<RPT, MINB, false, BLK, SYNC, WAVES, NARROW, LUT>(a, st);
}
// Builds of the frame-resident forward kernel (a.tune, include/rtt_b200.h RTT_MODE_TUNE_*), 256-thread blocks: 1 = 1 ray
// per thread, 2 = 2 rays (80 regs), 3 = 2 rays (64 regs, 4 blocks / SM), 5 = 1 ray at 48 registers / five blocks per SM;
// large blocks: 12 = 2 rays, one persistent block of 1024 threads per SM (the default), 7 / 13 = the same in four / two waves
// of blocks, 6 = 7 with a barrier per tile, 8 = two blocks of 512; 9 = the per-ray kernel of the EXACT variant's structure.  Measured and dropped: 4 rays per thread (128 registers,
// -20 %), 2 rays at 48 registers (spills, -17 %), 3 rays in a 768-thread block (80 registers: C2 4.21, c4cam 9.63), 1 ray in
// a 1024-thread block (C2 4.15), the packed-pair kernel in one block of 768 threads (C2 3.82, C1 2.83, C4 6.91), 2 rays in
// one block of 896 / 768 threads with 72 / 80 registers and no spills (C2 3.88 / 4.12 against 3.63 at 1024 threads and 64
// registers with 84 bytes of spills: resident warps beat registers), and the default build with the pair kernel's bulk-async
// ray streaming (2048-ray tiles in a two-slot ring, 2 x 80 KB: correct, but the block-wide barrier per tile that hands the
// slot to the elected thread stalls 32 warps at once — C2 3.84, C1 2.90, C4 5.74 against 3.63 / 2.30 / 5.43).

template <int MINB, bool STREAM, int LOG, int NP, bool GEN, int BLK = kThreads>
inline cudaError_t launch_pair_g(const SeqFwdArgs& a, cudaStream_t st) {
    const size_t smem = PairLayout<LOG, STREAM, NP, BLK>::bytes(a.tab.S, a.tab.L);
    if (cudaError_t e = allow_smem(k_trace_seq_fwd_pair<MINB, STREAM, LOG, NP, GEN, BLK>, smem)) return e;
    const long long tiles = (a.n + PairGeom<NP, BLK>::kTile - 1) / PairGeom<NP, BLK>::kTile;
    // persistent blocks, one resident set (tiles are handed out round-robin: tile = block + k * grid); the plain build
    // keeps the tile kernel's four waves (a block that lands on a busier SM costs 1/4 of a launch)
    long long g = (long long)sm_count() * MINB * (STREAM ? 1 : 4);
    if (tiles < g) g = tiles;
    if (g < 1) g = 1;
    k_trace_seq_fwd_pair<MINB, STREAM, LOG, NP, GEN, BLK><<<(int)g, BLK, smem, st>>>(a);
    return cudaGetLastError();
}
template <int MINB, bool STREAM, int LOG, int NP = 1, int BLK = kThreads>
inline cudaError_t launch_pair(const SeqFwdArgs& a, cudaStream_t st) {
    if (STREAM) return launch_pair_g<MINB, STREAM, LOG, NP, false, BLK>(a, st);       // streams rays from memory by construction
    return a.src.kind >= 0 ? launch_pair_g<MINB, false, LOG, NP, true, BLK>(a, st) : launch_pair_g<MINB, false, LOG, NP, false, BLK>(a, st);
}
// bulk-async copies need 16-byte aligned global addresses; every full tile starts a multiple of 512 rays into the
// arrays, so the base pointers decide.  Generated rays have no input to stream.
inline bool pair_can_stream(const SeqFwdArgs& a) {
    if (a.src.kind >= 0) return false;
    auto ok = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    if (!ok(a.pos) || !ok(a.dir) || !ok(a.inten)) return false;
    if (a.tab.L > 0 && !ok(a.wav)) return false;
    if (a.opos && (!ok(a.opos) || !ok(a.odir) || !ok(a.ointen))) return false;
    if (a.hitmask && !ok(a.hitmask)) return false;
    return true;
}
#endif

cudaError_t RTT_NAME(launch_seq_fwd)(const SeqFwdArgs& a, cudaStream_t st) {
#if defined(RTT_APPROX)
    // Default build (a.tune == 0): the frame-resident tile kernel with 2 rays per thread in ONE block of 1024 threads per
    // SM (64 registers).  Same-box A/B on B200 (profiles/r2_block_size_ab.md, ms per 1e8 rays, C2 / C1 / C4 / c4cam):
    // 3.65 / 2.35 / 5.47 / 7.27 against 4.00 / 2.46 / 5.81 / 7.77 for the same kernel in four blocks of 256 threads and
    // 3.92 / 2.51 / 6.92 / 9.44 for the packed-pair streaming kernel (16).  What the large block buys: the table staging,
    // image-cache set-up and flush of a block are shared by four times as many threads, and the 32 warps of an SM start
    // together and run the same code at nearly the same time, which the instruction caches like (the same effect, much
    // larger, in the non-sequential kernel).  One default for every table: a bundle generated in the kernel and its
    // materialised twin run the same build and stay bit-identical (tests/test_goals.py).
    // The blocks are persistent (one wave: a 1024-thread block sets up its table, frame changes and image cache once per
    // launch; with four waves C3's forward at 1e7 rays paid 4 % of a step for it: 1.054 -> 1.011 ms, C2 3.656 -> 3.624).
    int build = a.tune;
    if (build == 0) build = 12;
    switch (build) {
        case 1: return launch_tile<1, 4>(a, st);
        case 2: return launch_tile<2, 3>(a, st);
        case 3: return launch_tile<2, 4>(a, st);
        case 5: return launch_tile<1, 5>(a, st);
        case 6: return launch_tile_big<2, 1, 1024, true>(a, st);
        case 7: return launch_tile_big<2, 1, 1024, false>(a, st);
        case 8: return launch_tile_big<2, 2, 512, false>(a, st);
        case 12:
            if (a.tab.S > 32) return launch_tile_big<2, 1, 1024, false, 1>(a, st);
            return a.tab.L > 0 ? launch_tile_big<2, 1, 1024, false, 1, true, 1>(a, st) : launch_tile_big<2, 1, 1024, false, 1, true, 0>(a, st);
        case 13: return launch_tile_big<2, 1, 1024, false, 2>(a, st);
        case 16: return pair_can_stream(a) ? launch_pair<3, true, 11>(a, st) : launch_pair<3, false, 12>(a, st);
        case 17: return launch_pair<4, false, 12>(a, st);
        case 18: return launch_pair<3, false, 12>(a, st);
        case 19: return pair_can_stream(a) ? launch_pair<2, true, 12>(a, st) : launch_pair<3, false, 12>(a, st);
        default: break;
    }
#endif
    if (cudaError_t e = allow_smem(RTT_NAME(k_trace_seq_fwd), fwd_smem(a.tab.S, a.tab.L))) return e;
    RTT_NAME(k_trace_seq_fwd)<<<grid_for(a.n, 8), kThreads, fwd_smem(a.tab.S, a.tab.L), st>>>(a);
    return cudaGetLastError();
}
template <int MINB, bool POSE, int CK, bool LEAN, int GEN, int BLK = kThreads>
inline cudaError_t launch_seq_bwd_ck(const SeqBwdArgs& a, cudaStream_t st) {
    size_t smem = bwd_smem(a.tab.S, a.tab.L) + 8 + bwd_queue_bytes(BLK);
    if (LEAN) smem += sizeof(Xf) * (size_t)(a.tab.S + 1) + sizeof(unsigned) * (size_t)(a.tab.S + 4);   // frame changes, lean mask, walk list
    // chunk: as large as the queue allows, but small launches still spread over every resident block slot
    // (large blocks: one persistent block per SM; ~16 chunks per block keep the tail of the launch short)
    const int sms = sm_count() > 0 ? sm_count() : 1;
    const long long slots = (long long)sms * (BLK == kThreads ? 8 : MINB);
    const long long parts = BLK == kThreads ? slots : slots * 16;
    long long chunk = ((a.n + parts - 1) / parts + BLK - 1) / BLK * BLK;
    if (chunk < 4 * BLK) chunk = 4 * BLK;
    if (chunk > kBwdChunkIters * BLK) chunk = kBwdChunkIters * BLK;
    SeqBwdArgs b = a;
    b.chunk = (int)chunk;
    const long long chunks = (a.n + chunk - 1) / chunk;
    long long g = slots;
    if (chunks < g) g = chunks;
    if (g < 1) g = 1;
    if (cudaError_t e = allow_smem(RTT_NAME(k_trace_seq_bwd)<MINB, POSE, CK, LEAN, GEN, BLK>, smem)) return e;
    RTT_NAME(k_trace_seq_bwd)<MINB, POSE, CK, LEAN, GEN, BLK><<<(int)g, BLK, smem, st>>>(b);
    return cudaGetLastError();
}
template <int MINB, bool POSE>
inline cudaError_t launch_seq_bwd_as(const SeqBwdArgs& a, cudaStream_t st) {
    return a.tab.S <= 24 ? launch_seq_bwd_ck<MINB, POSE, 24, false, 0>(a, st)
                         : launch_seq_bwd_ck<MINB, POSE, RTT_MAX_ROWS, false, 0>(a, st);
}
#if defined(RTT_APPROX)
template <int GEN, int MINB, int BLK>
inline cudaError_t launch_seq_bwd_lean(const SeqBwdArgs& a, cudaStream_t st) {
    return a.tab.S <= 24 ? launch_seq_bwd_ck<MINB, false, 24, true, GEN, BLK>(a, st)
                         : launch_seq_bwd_ck<MINB, false, RTT_MAX_ROWS, true, GEN, BLK>(a, st);
}
#endif
cudaError_t RTT_NAME(launch_seq_bwd)(const SeqBwdArgs& a, cudaStream_t st) {
    // a.tune & 7 = resident blocks per SM the build is compiled for.  Without pose gradients the kernel needs 64-72
    // registers: four resident blocks (measured: C2 -10 %, C4 -5 % against three); with them three blocks (80 registers,
    // a few spills) beat two (~110 registers) by 4-8 %.  a.tune & 8: no lean path (every ray through the general code).
    // a.tune & 16: the lean build in four blocks of 256 threads instead of one block of 1024 per SM (A/B).
    const int minb = a.tune & 7;
    if (a.scalar_grads) {
#if defined(RTT_APPROX)
        if (!(a.tune & 8) && minb != 3 && !a.g_pos && !a.g_dir && !a.g_inten) {
            // one 1024-thread block per SM from ~3e7 rays on (C2, 1e8 rays: 3.31 -> 3.19 ms, C4 12.55 -> 12.12); below, the
            // chunks of a large block are too few to balance the SMs (C3, 1e7 rays: 0.526 ms in 256-thread blocks, 0.548 in one)
            if ((a.tune & 16) || a.n < (1ll << 25))
                return a.src.kind >= 0 ? launch_seq_bwd_lean<2, 4, kThreads>(a, st) : launch_seq_bwd_lean<1, 4, kThreads>(a, st);
            return a.src.kind >= 0 ? launch_seq_bwd_lean<2, 1, 1024>(a, st) : launch_seq_bwd_lean<1, 1, 1024>(a, st);
        }
#endif
        if (minb == 3) return launch_seq_bwd_as<3, false>(a, st);
        return launch_seq_bwd_as<4, false>(a, st);
    }
    if (minb == 2) return launch_seq_bwd_as<2, true>(a, st);
    return launch_seq_bwd_as<3, true>(a, st);
}
cudaError_t RTT_NAME(launch_nonseq_fwd)(const NonseqFwdArgs& a, cudaStream_t st) {
    if (cudaError_t e = allow_smem(RTT_NAME(k_trace_nonseq_fwd), nonseq_fwd_smem(a.tab.S, a.tab.L))) return e;
    if (a.hit_seq && a.nbounces > 0)
        if (cudaError_t e = cudaMemsetAsync(a.hit_seq, 0xFF, (size_t)a.n * a.nbounces, st)) return e;
    // Default: the lock-step build, one block of 1024 threads per SM with one barrier per trip (C5: 67.6 -> 56.0 ms EXACT,
    // 48.8 -> 44.0 FAST; barriers inside the search cost more than they save: profiles/r2_nonseq_lockstep.md).
    // a.tune (A/B): 7 = the free-running kernel of 256-thread blocks; 1 = 256 threads with the barrier; 3 / 4 / 5 = extra
    // barriers (every 4th row / before the interaction / both halves of the search); 6 = 1024 threads, no barrier.
    if (a.tune != 7) {
        const size_t smem = nonseq_fwd_smem(a.tab.S, a.tab.L);
        const int sms = sm_count() > 0 ? sm_count() : 1;
#define RTT_LS(BLOCK, SYNC, WAVES)                                                                                  \
        {                                                                                                           \
            if (cudaError_t e = allow_smem(RTT_NAME(k_trace_nonseq_fwd_ls)<BLOCK, SYNC>, smem)) return e;           \
            long long g = (long long)sms * (1024 / BLOCK) * WAVES;                                                  \
            const long long tiles = (a.n + BLOCK - 1) / BLOCK;                                                      \
            if (tiles < g) g = tiles;                                                                               \
            RTT_NAME(k_trace_nonseq_fwd_ls)<BLOCK, SYNC><<<(int)(g < 1 ? 1 : g), BLOCK, smem, st>>>(a);             \
            return cudaGetLastError();                                                                              \
        }
        if (a.tune == 1) RTT_LS(256, 1, 2)
        if (a.tune == 3) RTT_LS(1024, 3, 1)
        if (a.tune == 4) RTT_LS(1024, 4, 1)
        if (a.tune == 5) RTT_LS(1024, 5, 1)
#if defined(RTT_APPROX)
        if (a.tune == 2) RTT_LS(1024, 1, 1)
        RTT_LS(1024, 0, 1)                 // FAST arithmetic: the large block alone (43.0 ms; with the barrier 44.1)
#else
        if (a.tune == 6) RTT_LS(1024, 0, 1)
        RTT_LS(1024, 1, 1)                 // EXACT: 56.0 ms with the barrier, 64.9 without
#endif
#undef RTT_LS
    }
    RTT_NAME(k_trace_nonseq_fwd)<<<grid_for(a.n, 8), kThreads, nonseq_fwd_smem(a.tab.S, a.tab.L), st>>>(a);
    return cudaGetLastError();
}
cudaError_t RTT_NAME(launch_nonseq_bwd)(const NonseqBwdArgs& a, cudaStream_t st) {
    if (cudaError_t e = allow_smem(RTT_NAME(k_trace_nonseq_bwd), bwd_smem(a.tab.S, a.tab.L))) return e;
    RTT_NAME(k_trace_nonseq_bwd)<<<grid_for(a.n, 4), kThreads, bwd_smem(a.tab.S, a.tab.L), st>>>(a);
    return cudaGetLastError();
}
cudaError_t RTT_NAME(launch_intersect_test)(const IsectArgs& a, cudaStream_t st) {
    if (cudaError_t e = allow_smem(RTT_NAME(k_intersect_test), smem_table_bytes(a.tab.S, 0))) return e;
    RTT_NAME(k_intersect_test)<<<grid_for(a.n, 8), kThreads, smem_table_bytes(a.tab.S, 0), st>>>(a);
    return cudaGetLastError();
}
cudaError_t RTT_NAME(launch_step_fwd)(const StepFwdArgs& a, cudaStream_t st) {
    if (cudaError_t e = allow_smem(RTT_NAME(k_surface_step_fwd), smem_table_bytes(a.tab.S, a.tab.L))) return e;
    RTT_NAME(k_surface_step_fwd)<<<grid_for(a.n, 8), kThreads, smem_table_bytes(a.tab.S, a.tab.L), st>>>(a);
    return cudaGetLastError();
}
cudaError_t RTT_NAME(launch_step_bwd)(const StepBwdArgs& a, cudaStream_t st) {
    if (cudaError_t e = allow_smem(RTT_NAME(k_surface_step_bwd), bwd_smem(a.tab.S, a.tab.L))) return e;
    RTT_NAME(k_surface_step_bwd)<<<grid_for(a.n, 4), kThreads, bwd_smem(a.tab.S, a.tab.L), st>>>(a);
    return cudaGetLastError();
}

cudaError_t RTT_NAME(launch_render)(const RenderArgs& a, cudaStream_t st) {
    if (cudaError_t e = allow_smem(RTT_NAME(k_render_shade), smem_table_bytes(a.tab.S, 0))) return e;
    RTT_NAME(k_render_shade)<<<grid_for(a.n, 8), kThreads, smem_table_bytes(a.tab.S, 0), st>>>(a);
    return cudaGetLastError();
}

cudaError_t RTT_NAME(launch_sample)(const SampleArgs& a, cudaStream_t st) {
    RTT_NAME(k_sample_bundle)<<<grid_for(a.n, 8), kThreads, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace RTT_VARIANT
}  // namespace rtt
