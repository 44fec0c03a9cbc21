// rtt_kernels_decl.h — argument records and host launchers of one kernel variant.
// Included once per variant (RTT_VARIANT = fast | exact) by rtt_kernels.inl and rtt_cabi.cu;
// deliberately has no include guard.
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/rtt_b200.h"

#ifndef RTT_CAT
#define RTT_CAT2(a, b) a##b
#define RTT_CAT(a, b) RTT_CAT2(a, b)
#define RTT_NAME(base) RTT_CAT(base, RTT_CAT(_, RTT_VARIANT))
#endif

namespace rtt {
namespace RTT_VARIANT {

struct SourceDev {                  // ray source (rtt_source_t); kind < 0 = rays are read from memory
    int kind;
    float a[4];
    int width, height;
    const float* pose;
    unsigned long long seed;
    long long first;
    const unsigned long long* state;
    float intensity, wavelength;
};

struct SampleArgs {
    SourceDev src;
    float *pos, *dir, *inten, *wav;
    long long n;
};

struct SensorDev {
    float* image;
    float* record;
    int H, W, C;
    float x0, y0, sx, sy;
    int K;                      // record_hits
    unsigned char* count;
};

struct TableDev {
    const float* f;
    const int32_t* i;
    int S, L;
    const float* lut;
    const float* lut_w;
};

struct SeqFwdArgs {
    SourceDev src;
    const float *pos, *dir, *inten, *wav;
    float *opos, *odir, *ointen;
    unsigned long long* hitmask;
    TableDev tab;
    SensorDev sens[RTT_MAX_SENSORS];
    int n_sens;
    long long n;
    int tune;                   // (mode & RTT_MODE_TUNE_MASK) >> RTT_MODE_TUNE_SHIFT: kernel build, 0 = default
};

struct SeqBwdArgs {
    SourceDev src;
    const float *pos, *dir, *inten, *wav;
    const unsigned long long* hitmask;
    const float *g_opos, *g_odir, *g_ointen;
    const float* g_record[RTT_MAX_SENSORS];
    float *g_pos, *g_dir, *g_inten;
    float *g_table, *g_lut;
    TableDev tab;
    int n_sens;
    long long n;
    int chunk;                  // rays per block iteration (set by the launcher, <= kBwdChunk)
    int scalar_grads;           // RTT_MODE_SCALAR_GRADS: no row requests pose gradients
    int tune;                   // (mode & RTT_MODE_TUNE_MASK) >> RTT_MODE_TUNE_SHIFT: resident blocks per SM, 0 = default
};

struct NonseqFwdArgs {
    SourceDev src;
    const float *pos, *dir, *inten, *wav;
    float *opos, *odir, *ointen;
    unsigned char *hit_seq, *n_hits;
    TableDev tab;
    SensorDev sens[RTT_MAX_SENSORS];
    int n_sens, nbounces;
    long long n;
    int tune;                   // (mode & RTT_MODE_TUNE_MASK) >> RTT_MODE_TUNE_SHIFT: kernel build, 0 = default
};

struct NonseqBwdArgs {
    SourceDev src;
    const float *pos, *dir, *inten, *wav;
    const unsigned char* hit_seq;
    const float *g_opos, *g_odir, *g_ointen;
    const float* g_record[RTT_MAX_SENSORS];
    int rec_hits[RTT_MAX_SENSORS];
    float *g_pos, *g_dir, *g_inten;
    float *g_table, *g_lut;
    TableDev tab;
    int n_sens, nbounces;
    long long n;
};

struct IsectArgs {
    const float *pos, *dir;
    float* t_out;
    TableDev tab;
    int row0, k;
    long long n;
};

struct StepFwdArgs {
    const float *pos, *dir, *wav;
    float *npos, *ndir, *mod, *hit_local, *t_out, *normal;
    TableDev tab;
    int row;
    long long n;
};

struct StepBwdArgs {
    const float *pos, *dir, *wav;
    const float *g_npos, *g_ndir, *g_hit_local, *g_t, *g_normal;
    float *g_pos, *g_dir, *g_table, *g_lut;
    TableDev tab;
    int row;
    long long n;
};

struct RenderArgs {                 // Renderer.render_3d: nearest hit of every ray + Lambert shading
    SourceDev src;
    const float *pos, *dir;
    TableDev tab;
    const float* base_rgb;          // [S,3] base colour of every table row
    float light[3], bg[3];
    float* rgb;                     // [n,3]
    unsigned char* win;             // [n] winning row, 255 = background (optional)
    long long n;
};

cudaError_t RTT_NAME(launch_render)(const RenderArgs& a, cudaStream_t st);
cudaError_t RTT_NAME(launch_seq_fwd)(const SeqFwdArgs& a, cudaStream_t st);
cudaError_t RTT_NAME(launch_seq_bwd)(const SeqBwdArgs& a, cudaStream_t st);
cudaError_t RTT_NAME(launch_nonseq_fwd)(const NonseqFwdArgs& a, cudaStream_t st);
cudaError_t RTT_NAME(launch_nonseq_bwd)(const NonseqBwdArgs& a, cudaStream_t st);
cudaError_t RTT_NAME(launch_intersect_test)(const IsectArgs& a, cudaStream_t st);
cudaError_t RTT_NAME(launch_step_fwd)(const StepFwdArgs& a, cudaStream_t st);
cudaError_t RTT_NAME(launch_step_bwd)(const StepBwdArgs& a, cudaStream_t st);
cudaError_t RTT_NAME(launch_sample)(const SampleArgs& a, cudaStream_t st);

}  // namespace RTT_VARIANT
}  // namespace rtt
