// rtt_cabi.cu — extern "C" entry points declared in include/rtt_b200.h.
// Argument validation, variant dispatch (fast / exact), launch accounting.  No torch types.
#include <cuda_runtime.h>
#include <atomic>
#include <cstring>

#include "../../include/rtt_b200.h"
#include "rtt_internal.h"

// Each variant TU defines the same structs in its own namespace; re-declare what we launch.
#define RTT_VARIANT fast
#include "rtt_kernels_decl.h"
#undef RTT_VARIANT
#define RTT_VARIANT exact
#include "rtt_kernels_decl.h"
#undef RTT_VARIANT

namespace {

std::atomic<long long> g_launches{0};

int check_table(const rtt_table_t* t) {
    if (!t || !t->f || !t->i) return RTT_E_ARG;
    if (t->n_rows < 1 || t->n_rows > RTT_MAX_ROWS) return RTT_E_ROWS;
    if (t->n_lut < 0 || t->n_lut > RTT_MAX_WAVELENGTHS) return RTT_E_ARG;
    if (t->n_lut > 0 && (!t->lut || !t->lut_w)) return RTT_E_ARG;
    return RTT_OK;
}

int have_device() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n < 1) { cudaGetLastError(); return 0; }
    return 1;
}

template <class TD>
TD make_table(const rtt_table_t* t) {
    TD d;
    d.f = t->f; d.i = t->i; d.S = t->n_rows; d.L = t->n_lut; d.lut = t->lut; d.lut_w = t->lut_w;
    return d;
}

template <class SD>
int fill_sensors(SD* dst, const rtt_sensor_t* src, int n) {
    if (n < 0 || n > RTT_MAX_SENSORS) return RTT_E_SENSOR;
    if (n > 0 && !src) return RTT_E_SENSOR;
    for (int s = 0; s < RTT_MAX_SENSORS; ++s) {
        SD q;
        std::memset(&q, 0, sizeof(q));
        if (s < n) {
            if (src[s].image && (src[s].height < 1 || src[s].width < 1 || src[s].channels < 1)) return RTT_E_SENSOR;
            if (src[s].image && (long long)src[s].height * src[s].width * src[s].channels >= (1ll << 28)) return RTT_E_SENSOR;
            if (src[s].record && (reinterpret_cast<uintptr_t>(src[s].record) & 15)) return RTT_E_ALIGN;
            q.image = src[s].image; q.record = src[s].record;
            q.K = src[s].record_hits < 1 ? 1 : src[s].record_hits; q.count = src[s].count;
            q.H = src[s].height; q.W = src[s].width; q.C = src[s].channels;
            q.x0 = src[s].x0; q.y0 = src[s].y0; q.sx = src[s].sx; q.sy = src[s].sy;
        }
        dst[s] = q;
    }
    return RTT_OK;
}

int finish(cudaError_t e) {
    if (e == cudaSuccess) { g_launches.fetch_add(1); return RTT_OK; }
    return (int)e;
}

// rtt_source_t -> the kernels' by-value copy; kind = -1 when the rays come from memory
template <class SD>
int fill_source(SD& d, const rtt_source_t* s) {
    std::memset(&d, 0, sizeof(d));
    d.kind = -1;
    if (!s) return RTT_OK;
    if (s->kind < RTT_SRC_DISK || s->kind > RTT_SRC_CAMERA || !s->pose) return RTT_E_ARG;
    if (s->kind == RTT_SRC_CAMERA && (s->width < 1 || s->height < 1)) return RTT_E_ARG;
    d.kind = s->kind;
    for (int k = 0; k < 4; ++k) d.a[k] = s->a[k];
    d.width = s->width; d.height = s->height; d.pose = s->pose;
    d.seed = s->seed; d.first = s->first;
    d.state = reinterpret_cast<const unsigned long long*>(s->state);
    d.intensity = s->intensity; d.wavelength = s->wavelength;
    return RTT_OK;
}

}  // namespace

int rtt_internal_finish(cudaError_t e) { return finish(e); }
int rtt_internal_have_device() { return have_device(); }

// FP32 issue-rate probe: 8 independent FMA chains per thread, nothing else in the loop.
__global__ void __launch_bounds__(256) k_probe_fp32(int iters, float* out) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.0f, a2 = a0 + 2.0f, a3 = a0 + 3.0f;
    float a4 = a0 + 4.0f, a5 = a0 + 5.0f, a6 = a0 + 6.0f, a7 = a0 + 7.0f;
    const float m = 0.999f + blockIdx.x * 1e-9f, c = 1e-3f;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
            a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
        }
    }
    const float s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 123.456f) out[0] = s;                                      // keeps the chains alive
}

#define RTT_DISPATCH(mode, call_fast, call_exact) (((mode) & RTT_MODE_ARITH_MASK) == RTT_MODE_EXACT ? (call_exact) : (call_fast))

extern "C" {

int rtt_version(void) { return 100; }

int rtt_layout_query(int which) {
    switch (which) {
        case 0: return RTT_ROW_F;
        case 1: return RTT_ROW_I;
        case 2: return RTT_ROW_G;
        case 3: return RTT_MAX_ROWS;
        case 4: return RTT_N_DIFF;
        case 5: return RTT_MAX_SENSORS;
        case 6: return RTT_MAX_WAVELENGTHS;
        case 7: return RTT_MAX_BOUNCES;
        default: return -1;
    }
}

const char* rtt_error_string(int code) {
    switch (code) {
        case RTT_OK: return "ok";
        case RTT_E_ARG: return "rtt: invalid argument (null pointer or size out of range)";
        case RTT_E_ROWS: return "rtt: surface table must have 1..RTT_MAX_ROWS rows";
        case RTT_E_NO_DEVICE: return "rtt: no CUDA device; this library has no CPU path";
        case RTT_E_ALIGN: return "rtt: sensor record buffer must be 16-byte aligned";
        case RTT_E_SENSOR: return "rtt: invalid sensor request";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "rtt: unknown error";
    }
}

int64_t rtt_launch_count(void) { return (int64_t)g_launches.load(); }

int64_t rtt_probe_fp32(int32_t iters, float* scratch, void* stream) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return RTT_E_NO_DEVICE;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) return RTT_E_NO_DEVICE;
    if (iters < 1 || !scratch) return RTT_E_ARG;
    const int blocks = sms * 8;
    k_probe_fp32<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, scratch);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return -(int64_t)e - 1000;
    return (int64_t)blocks * 256 * (int64_t)iters * 64 * 2;            // FLOPs of the launch (FMA = 2)
}

int rtt_trace_seq_fwd(const float* in_pos, const float* in_dir, const float* in_intensity,
                      const float* in_wavelength, const rtt_source_t* source,
                      float* out_pos, float* out_dir, float* out_intensity, uint64_t* hitmask,
                      const rtt_table_t* table, const rtt_sensor_t* sensors, int32_t n_sensors,
                      int64_t n, int32_t mode, void* stream) {
    if (int e = check_table(table)) return e;
    if (n == 0) return RTT_OK;   /* empty bundle: nothing to do, pointers may be NULL */
    if (n < 0) return RTT_E_ARG;
    if (!source && (!in_pos || !in_dir || !in_intensity)) return RTT_E_ARG;
    if (!source && table->n_lut > 0 && !in_wavelength) return RTT_E_ARG;
    if ((out_pos != nullptr) != (out_dir != nullptr) || (out_pos != nullptr) != (out_intensity != nullptr)) return RTT_E_ARG;
    if (!have_device()) return RTT_E_NO_DEVICE;
    auto st = (cudaStream_t)stream;
#define RTT_BODY(NS)                                                                                   \
    {                                                                                                  \
        rtt::NS::SeqFwdArgs a;                                                                         \
        if (int e = fill_source(a.src, source)) return e;                                              \
        a.pos = in_pos; a.dir = in_dir; a.inten = in_intensity; a.wav = in_wavelength;                 \
        a.opos = out_pos; a.odir = out_dir; a.ointen = out_intensity;                                  \
        a.hitmask = reinterpret_cast<unsigned long long*>(hitmask);                                    \
        a.tab = make_table<rtt::NS::TableDev>(table);                                                  \
        if (int e = fill_sensors(a.sens, sensors, n_sensors)) return e;                                \
        a.n_sens = n_sensors; a.n = n;                                                                 \
        a.tune = (mode & RTT_MODE_TUNE_MASK) >> RTT_MODE_TUNE_SHIFT;                                   \
        return finish(rtt::NS::launch_seq_fwd_##NS(a, st));                                            \
    }
    if ((mode & RTT_MODE_ARITH_MASK) == RTT_MODE_EXACT) RTT_BODY(exact) else RTT_BODY(fast)
#undef RTT_BODY
}

int rtt_trace_seq_bwd(const float* in_pos, const float* in_dir, const float* in_intensity,
                      const float* in_wavelength, const rtt_source_t* source, const uint64_t* hitmask,
                      const float* g_out_pos, const float* g_out_dir, const float* g_out_intensity,
                      const float* const* g_record,
                      float* g_in_pos, float* g_in_dir, float* g_in_intensity,
                      float* g_table, float* g_lut,
                      const rtt_table_t* table, int32_t n_sensors,
                      int64_t n, int32_t mode, void* stream) {
    if (int e = check_table(table)) return e;
    if (n == 0) return RTT_OK;   /* empty bundle: nothing to do, pointers may be NULL */
    if (n < 0 || !hitmask) return RTT_E_ARG;
    if (!source && (!in_pos || !in_dir || !in_intensity)) return RTT_E_ARG;
    if (!source && table->n_lut > 0 && !in_wavelength) return RTT_E_ARG;
    if (n_sensors < 0 || n_sensors > RTT_MAX_SENSORS) return RTT_E_SENSOR;
    if (!have_device()) return RTT_E_NO_DEVICE;
    auto st = (cudaStream_t)stream;
#define RTT_BODY(NS)                                                                                   \
    {                                                                                                  \
        rtt::NS::SeqBwdArgs a;                                                                         \
        if (int e = fill_source(a.src, source)) return e;                                              \
        a.pos = in_pos; a.dir = in_dir; a.inten = in_intensity; a.wav = in_wavelength;                 \
        a.hitmask = reinterpret_cast<const unsigned long long*>(hitmask);                              \
        a.g_opos = g_out_pos; a.g_odir = g_out_dir; a.g_ointen = g_out_intensity;                      \
        for (int s = 0; s < RTT_MAX_SENSORS; ++s) {                                                    \
            a.g_record[s] = (g_record && s < n_sensors) ? g_record[s] : nullptr;                       \
            if (a.g_record[s] && (reinterpret_cast<uintptr_t>(a.g_record[s]) & 15)) return RTT_E_ALIGN; \
        }                                                                                              \
        a.g_pos = g_in_pos; a.g_dir = g_in_dir; a.g_inten = g_in_intensity;                            \
        a.g_table = g_table; a.g_lut = (table->n_lut > 0) ? g_lut : nullptr;                           \
        a.tab = make_table<rtt::NS::TableDev>(table);                                                  \
        a.n_sens = n_sensors; a.n = n;                                                                 \
        a.scalar_grads = (mode & RTT_MODE_SCALAR_GRADS) ? 1 : 0;                                       \
        a.tune = (mode & RTT_MODE_TUNE_MASK) >> RTT_MODE_TUNE_SHIFT;                                   \
        return finish(rtt::NS::launch_seq_bwd_##NS(a, st));                                            \
    }
    if ((mode & RTT_MODE_ARITH_MASK) == RTT_MODE_EXACT) RTT_BODY(exact) else RTT_BODY(fast)
#undef RTT_BODY
}

int rtt_trace_nonseq_fwd(const float* in_pos, const float* in_dir, const float* in_intensity,
                         const float* in_wavelength, const rtt_source_t* source,
                         float* out_pos, float* out_dir, float* out_intensity,
                         uint8_t* hit_seq, uint8_t* n_hits,
                         const rtt_table_t* table, const rtt_sensor_t* sensors, int32_t n_sensors,
                         int32_t nbounces, int64_t n, int32_t mode, void* stream) {
    if (int e = check_table(table)) return e;
    if (n == 0) return RTT_OK;   /* empty bundle: nothing to do, pointers may be NULL */
    if (n < 0) return RTT_E_ARG;
    if (!source && (!in_pos || !in_dir || !in_intensity || !out_pos || !out_dir || !out_intensity)) return RTT_E_ARG;
    if ((out_pos != nullptr) != (out_dir != nullptr) || (out_pos != nullptr) != (out_intensity != nullptr)) return RTT_E_ARG;
    if (nbounces < 0 || (hit_seq && nbounces > RTT_MAX_BOUNCES)) return RTT_E_ARG;
    if (table->n_rows > 255) return RTT_E_ROWS;
    if (!source && table->n_lut > 0 && !in_wavelength) return RTT_E_ARG;
    if (!have_device()) return RTT_E_NO_DEVICE;
    auto st = (cudaStream_t)stream;
#define RTT_BODY(NS)                                                                                   \
    {                                                                                                  \
        rtt::NS::NonseqFwdArgs a;                                                                      \
        if (int e = fill_source(a.src, source)) return e;                                              \
        a.pos = in_pos; a.dir = in_dir; a.inten = in_intensity; a.wav = in_wavelength;                 \
        a.opos = out_pos; a.odir = out_dir; a.ointen = out_intensity;                                  \
        a.hit_seq = hit_seq; a.n_hits = n_hits;                                                        \
        a.tab = make_table<rtt::NS::TableDev>(table);                                                  \
        if (int e = fill_sensors(a.sens, sensors, n_sensors)) return e;                                \
        a.n_sens = n_sensors; a.nbounces = nbounces; a.n = n;                                          \
        a.tune = (mode & RTT_MODE_TUNE_MASK) >> RTT_MODE_TUNE_SHIFT;                                   \
        return finish(rtt::NS::launch_nonseq_fwd_##NS(a, st));                                         \
    }
    /* EXACT unless the caller opts in to the FAST arithmetic explicitly: see include/rtt_b200.h */
    if (mode & RTT_MODE_NONSEQ_FAST) RTT_BODY(fast) else RTT_BODY(exact)
#undef RTT_BODY
}

int rtt_trace_nonseq_bwd(const float* in_pos, const float* in_dir, const float* in_intensity,
                         const float* in_wavelength, const rtt_source_t* source,
                         const uint8_t* hit_seq, int32_t nbounces,
                         const float* g_out_pos, const float* g_out_dir, const float* g_out_intensity,
                         const float* const* g_record, const int32_t* record_hits,
                         float* g_in_pos, float* g_in_dir, float* g_in_intensity,
                         float* g_table, float* g_lut,
                         const rtt_table_t* table, int32_t n_sensors,
                         int64_t n, int32_t mode, void* stream) {
    if (int e = check_table(table)) return e;
    if (n == 0) return RTT_OK;   /* empty bundle: nothing to do, pointers may be NULL */
    if (n < 0 || !hit_seq) return RTT_E_ARG;
    if (!source && (!in_pos || !in_dir || !in_intensity)) return RTT_E_ARG;
    if (nbounces < 0 || nbounces > RTT_MAX_BOUNCES) return RTT_E_ARG;
    if (n_sensors < 0 || n_sensors > RTT_MAX_SENSORS) return RTT_E_SENSOR;
    if (!source && table->n_lut > 0 && !in_wavelength) return RTT_E_ARG;
    if (!have_device()) return RTT_E_NO_DEVICE;
    auto st = (cudaStream_t)stream;
#define RTT_BODY(NS)                                                                                   \
    {                                                                                                  \
        rtt::NS::NonseqBwdArgs a;                                                                      \
        if (int e = fill_source(a.src, source)) return e;                                              \
        a.pos = in_pos; a.dir = in_dir; a.inten = in_intensity; a.wav = in_wavelength;                 \
        a.hit_seq = hit_seq;                                                                           \
        a.g_opos = g_out_pos; a.g_odir = g_out_dir; a.g_ointen = g_out_intensity;                      \
        for (int s = 0; s < RTT_MAX_SENSORS; ++s) {                                                    \
            a.g_record[s] = (g_record && s < n_sensors) ? g_record[s] : nullptr;                       \
            a.rec_hits[s] = (record_hits && s < n_sensors && record_hits[s] > 1) ? record_hits[s] : 1; \
            if (a.g_record[s] && (reinterpret_cast<uintptr_t>(a.g_record[s]) & 15)) return RTT_E_ALIGN; \
        }                                                                                              \
        a.g_pos = g_in_pos; a.g_dir = g_in_dir; a.g_inten = g_in_intensity;                            \
        a.g_table = g_table; a.g_lut = (table->n_lut > 0) ? g_lut : nullptr;                           \
        a.tab = make_table<rtt::NS::TableDev>(table);                                                  \
        a.n_sens = n_sensors; a.nbounces = nbounces; a.n = n;                                          \
        return finish(rtt::NS::launch_nonseq_bwd_##NS(a, st));                                         \
    }
    /* EXACT unless the caller opts in to the FAST arithmetic explicitly: see include/rtt_b200.h */
    if (mode & RTT_MODE_NONSEQ_FAST) RTT_BODY(fast) else RTT_BODY(exact)
#undef RTT_BODY
}

int rtt_sample_bundle(const rtt_source_t* source, float* pos, float* dir, float* intensity, float* wavelength,
                      int64_t n, int32_t mode, void* stream) {
    if (n == 0) return RTT_OK;
    if (n < 0 || !source || !pos || !dir || !intensity) return RTT_E_ARG;
    if (!have_device()) return RTT_E_NO_DEVICE;
    auto st = (cudaStream_t)stream;
#define RTT_BODY(NS)                                                                                   \
    {                                                                                                  \
        rtt::NS::SampleArgs a;                                                                         \
        if (int e = fill_source(a.src, source)) return e;                                              \
        a.pos = pos; a.dir = dir; a.inten = intensity; a.wav = wavelength; a.n = n;                    \
        return finish(rtt::NS::launch_sample_##NS(a, st));                                             \
    }
    if ((mode & RTT_MODE_ARITH_MASK) == RTT_MODE_EXACT) RTT_BODY(exact) else RTT_BODY(fast)
#undef RTT_BODY
}

int rtt_intersect_test(const float* in_pos, const float* in_dir, float* t_out,
                       const rtt_table_t* table, int32_t row0, int32_t k,
                       int64_t n, int32_t mode, void* stream) {
    if (int e = check_table(table)) return e;
    if (n == 0) return RTT_OK;   /* empty bundle: nothing to do, pointers may be NULL */
    if (n < 0 || !in_pos || !in_dir || !t_out) return RTT_E_ARG;
    if (row0 < 0 || k < 1 || row0 + k > table->n_rows) return RTT_E_ROWS;
    if (!have_device()) return RTT_E_NO_DEVICE;
    auto st = (cudaStream_t)stream;
#define RTT_BODY(NS)                                                                                   \
    {                                                                                                  \
        rtt::NS::IsectArgs a;                                                                          \
        a.pos = in_pos; a.dir = in_dir; a.t_out = t_out;                                               \
        a.tab = make_table<rtt::NS::TableDev>(table);                                                  \
        a.row0 = row0; a.k = k; a.n = n;                                                               \
        return finish(rtt::NS::launch_intersect_test_##NS(a, st));                                     \
    }
    if ((mode & RTT_MODE_ARITH_MASK) == RTT_MODE_EXACT) RTT_BODY(exact) else RTT_BODY(fast)
#undef RTT_BODY
}

int rtt_surface_step_fwd(const float* in_pos, const float* in_dir, const float* in_wavelength,
                         float* new_pos, float* new_dir, float* mod,
                         float* hit_local, float* t_out, float* normal,
                         const rtt_table_t* table, int32_t row,
                         int64_t n, int32_t mode, void* stream) {
    if (int e = check_table(table)) return e;
    if (n == 0) return RTT_OK;   /* empty bundle: nothing to do, pointers may be NULL */
    if (n < 0 || !in_pos || !in_dir || !new_pos || !new_dir || !mod) return RTT_E_ARG;
    if (row < 0 || row >= table->n_rows) return RTT_E_ROWS;
    if (table->n_lut > 0 && !in_wavelength) return RTT_E_ARG;
    if (!have_device()) return RTT_E_NO_DEVICE;
    auto st = (cudaStream_t)stream;
#define RTT_BODY(NS)                                                                                   \
    {                                                                                                  \
        rtt::NS::StepFwdArgs a;                                                                        \
        a.pos = in_pos; a.dir = in_dir; a.wav = in_wavelength;                                         \
        a.npos = new_pos; a.ndir = new_dir; a.mod = mod;                                               \
        a.hit_local = hit_local; a.t_out = t_out; a.normal = normal;                                   \
        a.tab = make_table<rtt::NS::TableDev>(table);                                                  \
        a.row = row; a.n = n;                                                                          \
        return finish(rtt::NS::launch_step_fwd_##NS(a, st));                                           \
    }
    if ((mode & RTT_MODE_ARITH_MASK) == RTT_MODE_EXACT) RTT_BODY(exact) else RTT_BODY(fast)
#undef RTT_BODY
}

int rtt_surface_step_bwd(const float* in_pos, const float* in_dir, const float* in_wavelength,
                         const float* g_new_pos, const float* g_new_dir, const float* g_hit_local,
                         const float* g_t, const float* g_normal,
                         float* g_in_pos, float* g_in_dir, float* g_table, float* g_lut,
                         const rtt_table_t* table, int32_t row,
                         int64_t n, int32_t mode, void* stream) {
    if (int e = check_table(table)) return e;
    if (n == 0) return RTT_OK;   /* empty bundle: nothing to do, pointers may be NULL */
    if (n < 0 || !in_pos || !in_dir) return RTT_E_ARG;
    if (row < 0 || row >= table->n_rows) return RTT_E_ROWS;
    if (table->n_lut > 0 && !in_wavelength) return RTT_E_ARG;
    if (!have_device()) return RTT_E_NO_DEVICE;
    auto st = (cudaStream_t)stream;
#define RTT_BODY(NS)                                                                                   \
    {                                                                                                  \
        rtt::NS::StepBwdArgs a;                                                                        \
        a.pos = in_pos; a.dir = in_dir; a.wav = in_wavelength;                                         \
        a.g_npos = g_new_pos; a.g_ndir = g_new_dir; a.g_hit_local = g_hit_local;                       \
        a.g_t = g_t; a.g_normal = g_normal;                                                            \
        a.g_pos = g_in_pos; a.g_dir = g_in_dir; a.g_table = g_table;                                   \
        a.g_lut = (table->n_lut > 0) ? g_lut : nullptr;                                                \
        a.tab = make_table<rtt::NS::TableDev>(table);                                                  \
        a.row = row; a.n = n;                                                                          \
        return finish(rtt::NS::launch_step_bwd_##NS(a, st));                                           \
    }
    if ((mode & RTT_MODE_ARITH_MASK) == RTT_MODE_EXACT) RTT_BODY(exact) else RTT_BODY(fast)
#undef RTT_BODY
}

int rtt_render_shade(const float* in_pos, const float* in_dir, const rtt_source_t* source, const rtt_table_t* table,
                     const float* base_rgb, const float* light_dir, const float* background,
                     float* out_rgb, uint8_t* out_row, int64_t n, int32_t, void* stream) {
    if (int e = check_table(table)) return e;
    if (n == 0) return RTT_OK;
    if (n < 0 || !base_rgb || !light_dir || !background || !out_rgb) return RTT_E_ARG;
    if (!source && (!in_pos || !in_dir)) return RTT_E_ARG;
    if (!have_device()) return RTT_E_NO_DEVICE;
    rtt::exact::RenderArgs a;                          /* nearest-hit decisions: the reference's rounding (EXACT) */
    if (int e = fill_source(a.src, source)) return e;
    a.pos = in_pos; a.dir = in_dir;
    a.tab = make_table<rtt::exact::TableDev>(table);
    a.base_rgb = base_rgb;
    for (int k = 0; k < 3; ++k) { a.light[k] = light_dir[k]; a.bg[k] = background[k]; }
    a.rgb = out_rgb; a.win = out_row; a.n = n;
    return finish(rtt::exact::launch_render_exact(a, (cudaStream_t)stream));
}

}  // extern "C"
