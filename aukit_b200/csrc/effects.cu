// effects.cu -- K6 mono, K7 amplify, K8 absmax, K9 scale_clamp.
//
// Replaces Audio:mono (A:677-689), effects.amplify (A:3356-3369) and the two passes of
// effects.normalize (A:3431-3459).  All four are pure streaming kernels (128-bit loads and
// stores, grid-stride, no shared memory) and HBM-bound.  The arithmetic of each element is
// done in fp64 exactly as the reference does it (left-to-right channel sum then one division;
// x * multiplier; x * (peak / max)) and narrowed to float32 once at the end, so for identical
// inputs the stored result is the correctly rounded reference value.  HBM bandwidth leaves
// ample fp64 issue slots for that (<= 3 DP ops per 8-12 bytes moved).
//
// absmax: fmaxf drops NaN exactly like Lua's math.max(max, abs(x)) does when x is NaN
// (A:3442); per-thread maxima are combined with warp shuffles and ONE atomicMax per CTA on
// the float's bit pattern (non-negative floats order like unsigned integers).
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256)
mono_kernel(const float *__restrict__ in, size_t in_stride, int C, size_t n, float *__restrict__ out, int vec_ok) {
    const double cn = (double)C;
    // s / cn (A:687): for power-of-two channel counts the reciprocal multiply is the same
    // correctly rounded value and avoids an fp64 division per sample
    const bool pow2 = (C & (C - 1)) == 0;
    const double inv = 1.0 / cn;
    auto div_cn = [&](double s) { return pow2 ? s * inv : s / cn; };
    const size_t nvec = vec_ok ? n / 4 : 0;
    if (C == 2) {
        // two channels: (0 + a) + b is exact in fp64 and the division by 2 is exact, so the reference value is
        // the correctly rounded (a + b) / 2 -- which is exactly what one fp32 add and an exact halving produce
        for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < nvec; t += (size_t)gridDim.x * blockDim.x) {
            const uint4 a4 = ldg_stream(reinterpret_cast<const uint4 *>(in) + t);
            const uint4 b4 = ldg_stream(reinterpret_cast<const uint4 *>(in + in_stride) + t);
            stg_stream(reinterpret_cast<float4 *>(out) + t,
                       make_float4(__fadd_rn(__fadd_rn(0.0f, __uint_as_float(a4.x)), __uint_as_float(b4.x)) * 0.5f,
                                   __fadd_rn(__fadd_rn(0.0f, __uint_as_float(a4.y)), __uint_as_float(b4.y)) * 0.5f,
                                   __fadd_rn(__fadd_rn(0.0f, __uint_as_float(a4.z)), __uint_as_float(b4.z)) * 0.5f,
                                   __fadd_rn(__fadd_rn(0.0f, __uint_as_float(a4.w)), __uint_as_float(b4.w)) * 0.5f));
        }
        for (size_t i = nvec * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
            out[i] = __fadd_rn(__fadd_rn(0.0f, in[i]), in[in_stride + i]) * 0.5f;   // (0 + a) + b keeps the reference's zero sign
        return;
    }
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < nvec; t += (size_t)gridDim.x * blockDim.x) {
        double s0 = 0, s1 = 0, s2 = 0, s3 = 0;                          // local s = 0, A:685
        for (int c = 0; c < C; c++) {
            const uint4 r = ldg_stream(reinterpret_cast<const uint4 *>(in + (size_t)c * in_stride) + t);
            s0 += (double)__uint_as_float(r.x); s1 += (double)__uint_as_float(r.y);
            s2 += (double)__uint_as_float(r.z); s3 += (double)__uint_as_float(r.w);
        }
        stg_stream(reinterpret_cast<float4 *>(out) + t,
                   make_float4((float)div_cn(s0), (float)div_cn(s1), (float)div_cn(s2), (float)div_cn(s3)));
    }
    for (size_t i = nvec * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        double s = 0;
        for (int c = 0; c < C; c++) s += (double)in[(size_t)c * in_stride + i];
        out[i] = (float)div_cn(s);
    }
}

// in place: x <- clamp(x * mult, -1, 1); mult is a kernel argument (amplify) or peak / max
// read from device memory (normalize pass 2).
template <bool FROM_MAX>
__global__ void __launch_bounds__(256)
scale_clamp_kernel(float *__restrict__ d, size_t stride, size_t n, double mult_or_peak,
                   const float *__restrict__ d_max, int independent, int vec_ok) {
    float *row = d + (size_t)blockIdx.y * stride;
    double mult = mult_or_peak;
    if (FROM_MAX) mult = mult_or_peak / (double)d_max[independent ? blockIdx.y : 0];   // A:3444 / A:3451
    const size_t nvec = vec_ok ? n / 4 : 0;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < nvec; t += (size_t)gridDim.x * blockDim.x) {
        float4 v = reinterpret_cast<float4 *>(row)[t];
        v.x = (float)clamp_ref((double)v.x * mult); v.y = (float)clamp_ref((double)v.y * mult);
        v.z = (float)clamp_ref((double)v.z * mult); v.w = (float)clamp_ref((double)v.w * mult);
        stg_stream(reinterpret_cast<float4 *>(row) + t, v);
    }
    for (size_t i = nvec * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        row[i] = (float)clamp_ref((double)row[i] * mult);
}

__global__ void __launch_bounds__(256)
absmax_kernel(const float *__restrict__ d, size_t stride, size_t n, int independent, float *d_max, int vec_ok) {
    const float *row = d + (size_t)blockIdx.y * stride;
    float m = 0.0f;                                                     // local max = 0, A:3438
    const size_t nvec = vec_ok ? n / 4 : 0;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < nvec; t += (size_t)gridDim.x * blockDim.x) {
        const uint4 r = ldg_stream(reinterpret_cast<const uint4 *>(row) + t);
        m = fmaxf(m, fmaxf(fmaxf(fabsf(__uint_as_float(r.x)), fabsf(__uint_as_float(r.y))),
                           fmaxf(fabsf(__uint_as_float(r.z)), fabsf(__uint_as_float(r.w)))));
    }
    for (size_t i = nvec * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        m = fmaxf(m, fabsf(row[i]));
    m = warp_max(m);
    __shared__ float wm[8];
    if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = threadIdx.x < (blockDim.x >> 5) ? wm[threadIdx.x] : 0.0f;
        m = warp_max(m);
        if (threadIdx.x == 0) atomic_max_nonneg(d_max + (independent ? blockIdx.y : 0), m);
    }
}

bool vec4_ok(const void *p, size_t stride) { return ((uintptr_t)p % 16 == 0) && (stride % 4 == 0); }

}  // namespace

extern "C" int aukit_cuda_dev_mono(aukit_ctx *ctx, const float *d_in, size_t in_stride, int channels, size_t n,
                                   float *d_out) {
    if (!ctx) return aukit_fail("aukit_cuda: null context");
    if (channels < 1) return aukit_fail("aukit_cuda: channels < 1");
    if (n == 0) return 0;
    const int threads = 256;
    const int vok = vec4_ok(d_in, in_stride) && vec4_ok(d_out, 4);
    const unsigned grid = aukit_grid((n + 3) / 4, threads, (size_t)ctx->num_sms * 8 * 4);
    mono_kernel<<<grid, threads, 0, ctx->stream>>>(d_in, in_stride, channels, n, d_out, vok);
    ctx->launches++;
    return aukit_cuda_check(cudaGetLastError(), "mono_kernel launch");
}

extern "C" int aukit_cuda_dev_amplify(aukit_ctx *ctx, float *d, size_t stride, int channels, size_t n,
                                      double multiplier) {
    if (!ctx) return aukit_fail("aukit_cuda: null context");
    if (multiplier == 1.0 || n == 0 || channels < 1) return 0;          // A:3359
    const int threads = 256;
    dim3 grid(aukit_grid((n + 3) / 4, threads, (size_t)ctx->num_sms * 8 * 4), channels);
    scale_clamp_kernel<false><<<grid, threads, 0, ctx->stream>>>(d, stride, n, multiplier, nullptr, 0, vec4_ok(d, stride));
    ctx->launches++;
    return aukit_cuda_check(cudaGetLastError(), "amplify launch");
}

extern "C" int aukit_cuda_dev_absmax(aukit_ctx *ctx, const float *d, size_t stride, int channels, size_t n,
                                     int independent, float *d_max) {
    if (!ctx) return aukit_fail("aukit_cuda: null context");
    if (n == 0 || channels < 1) return 0;
    const int threads = 256;
    dim3 grid(aukit_grid((n + 3) / 4, threads * 4, (size_t)ctx->num_sms * 8), channels);
    absmax_kernel<<<grid, threads, 0, ctx->stream>>>(d, stride, n, independent, d_max, vec4_ok(d, stride));
    ctx->launches++;
    return aukit_cuda_check(cudaGetLastError(), "absmax_kernel launch");
}

extern "C" int aukit_cuda_dev_scale_clamp(aukit_ctx *ctx, float *d, size_t stride, int channels, size_t n,
                                          double peakAmplitude, int independent, const float *d_max) {
    if (!ctx) return aukit_fail("aukit_cuda: null context");
    if (n == 0 || channels < 1) return 0;
    const int threads = 256;
    dim3 grid(aukit_grid((n + 3) / 4, threads, (size_t)ctx->num_sms * 8 * 4), channels);
    scale_clamp_kernel<true><<<grid, threads, 0, ctx->stream>>>(d, stride, n, peakAmplitude, d_max, independent,
                                                                vec4_ok(d, stride));
    ctx->launches++;
    return aukit_cuda_check(cudaGetLastError(), "scale_clamp launch");
}
