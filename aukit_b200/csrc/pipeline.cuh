// pipeline.cuh -- argument block shared by the fused-pipeline kernels.
#pragma once
#include "common.cuh"

struct pipe_args {
    const uint8_t *in;            // interleaved frames [in_first, in_first + in_avail)
    int channels;
    unsigned long long n_total, in_first;
    size_t in_avail;              // frames present in `in`
    double ratio;
    unsigned long long out_first;
    size_t n_out;
    int mono;
    float inv_cn;                 // 1 / channels (exact when channels is a power of two)
    int cn_pow2;
    float *d_max;                 // peak pass: out; apply pass: in
    double peak;
    float *out;
    size_t out_stride;
    // standalone Audio:resample through the polyphase kernels: planar float32 input rows, raw output
    int planar_f32;               // `in` holds float rows, in_stride floats apart (instead of packed interleaved bytes)
    size_t in_stride;
    int raw_out;                  // store the resampled value itself (no normalize scale / clamp)
    // Performance hint only (results do not depend on it): *hint == epoch means an earlier launch of the same call found
    // that the per-channel clamp of A:668 acts on this signal, so run_static_kernel starts on its clamping twin.
    int *hint;
    int epoch;
    // wide-frame kernel (K16): y = RN(1 / ratio) and whether the three-operation quotient fma(fma(-q0, ratio, n), y, q0),
    // q0 = RN(n * y), is proved to equal the IEEE division n / ratio over the call's index range (aukit_quotient_fma_is_exact)
    double y;
    int qfma_ok;
};

// implemented in pipeline_poly.cu; returns 1 when it handled the launch, 0 when the generic
// path must run, -1 on error
int aukit_pipeline_poly_try(aukit_ctx *ctx, const pipe_args &a, const aukit_pipeline_desc *p, bool apply);
// Audio:resample on planar float32 through the same polyphase kernels (resample.cu calls it first); same return convention
int aukit_poly_resample_try(aukit_ctx *ctx, const float *d_in, size_t in_stride, int channels, unsigned long long n_in_total,
                            unsigned long long in_first, size_t in_avail, double srcRate, double dstRate, int interpolation,
                            unsigned long long out_first, size_t n_out, float *d_out, size_t out_stride);
// the same call with TMA-staged, double-buffered tiles (resample_planar.cu: L <= 256, positions < 2^28); same return convention
int aukit_planar_resample_try(aukit_ctx *ctx, const float *d_in, size_t in_stride, int channels, unsigned long long n_in_total,
                              unsigned long long in_first, size_t in_avail, double srcRate, double dstRate, int interpolation,
                              unsigned long long out_first, size_t n_out, float *d_out, size_t out_stride);
// interpolate.sinc (A:267-281) on planar float32, whole range (resample_planar.cu); same return convention
int aukit_planar_sinc_try(aukit_ctx *ctx, const float *d_in, size_t in_stride, int channels, unsigned long long n_in_total,
                          unsigned long long in_first, size_t in_avail, double srcRate, double dstRate,
                          unsigned long long out_first, size_t n_out, float *d_out, size_t out_stride);
// ratio 2^-k (L == 1): every output is a copied sample -- pipeline_decim.cu; same return convention
int aukit_pipeline_decim_try(aukit_ctx *ctx, const pipe_args &a, const aukit_pipeline_desc *p, bool apply, long long M);
