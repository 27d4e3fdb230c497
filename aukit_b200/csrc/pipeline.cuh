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
};

// implemented in pipeline_poly.cu; returns 1 when it handled the launch, 0 when the generic
// path must run, -1 on error
int aukit_pipeline_poly_try(aukit_ctx *ctx, const pipe_args &a, const aukit_pipeline_desc *p, bool apply);
