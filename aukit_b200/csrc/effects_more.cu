// effects_more.cu -- the remaining elementwise in-place effects of SURVEY 8(f) rank 4:
//   effects.invert (A:3412-3419), effects.fade (A:3392-3410), effects.delay (A:3500-3513),
//   effects.center (A:3465-3478).
// All HBM-streaming (4 B read + 4 B written per touched sample), arithmetic in fp64 like the reference's Lua
// numbers, NaN-transparent clamp (A:228).  Where the reference's loops would index a table with a non-integer
// or out-of-range key (and raise "attempt to perform arithmetic on a nil value" half way through), the entry
// points raise the same message BEFORE touching the audio.
#include "common.cuh"

#include <math.h>

namespace {

// Rows are processed as [head | float4 body | tail] around the first 16-byte aligned element, so any row offset
// (fade starts mid-row) keeps 128-bit accesses.
template <class F>
__device__ __forceinline__ void for_each_vec(float *row, size_t n, F f) {
    const size_t mis = ((uintptr_t)row >> 2) & 3, head = mis ? (4 - mis < n ? 4 - mis : n) : 0;
    const size_t nvec = (n - head) / 4, tail0 = head + nvec * 4;
    const size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x, step = (size_t)gridDim.x * blockDim.x;
    float4 *body = reinterpret_cast<float4 *>(row + head);
    size_t t = t0;
    for (; t + 3 * step < nvec; t += 4 * step) {                         // four loads in flight per thread
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) v[u] = body[t + u * step];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const size_t i = head + 4 * (t + u * step);
            v[u].x = f(v[u].x, i); v[u].y = f(v[u].y, i + 1); v[u].z = f(v[u].z, i + 2); v[u].w = f(v[u].w, i + 3);
            stg_stream(body + t + u * step, v[u]);
        }
    }
    for (; t < nvec; t += step) {
        float4 v = body[t];
        const size_t i = head + 4 * t;
        v.x = f(v.x, i); v.y = f(v.y, i + 1); v.z = f(v.z, i + 2); v.w = f(v.w, i + 3);
        stg_stream(body + t, v);
    }
    if (t0 < head) row[t0] = f(row[t0], t0);
    if (tail0 + t0 < n) row[tail0 + t0] = f(row[tail0 + t0], tail0 + t0);
}

__global__ void __launch_bounds__(256) invert_kernel(float *__restrict__ d, size_t stride, size_t n) {
    for_each_vec(d + (size_t)blockIdx.y * stride, n, [](float v, size_t) { return -v; });
}

// ch[i] = clamp(ch[i] * (m * (i - start) + startAmplitude)) for 1-based i = start .. start + count - 1
__global__ void __launch_bounds__(256)
fade_kernel(float *__restrict__ d, size_t stride, size_t first0, size_t count, double m, double startAmp) {
    for_each_vec(d + (size_t)blockIdx.y * stride + first0, count, [=](float v, size_t k) {
        const double g = __dadd_rn(__dmul_rn(m, (double)k), startAmp);                // i - start == k exactly
        return (float)clamp_ref(__dmul_rn((double)v, g));
    });
}

// o[i] = clamp(o[i] + original[i - samples] * multiplier) for i > samples (0-based: i >= samples); `orig` is a copy
__global__ void __launch_bounds__(256)
delay_kernel(float *__restrict__ d, const float *__restrict__ orig, size_t stride, size_t n, size_t samples, double mult) {
    float *row = d + (size_t)blockIdx.y * stride;
    const float *src = orig + (size_t)blockIdx.y * stride;
    for (size_t i = samples + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        row[i] = (float)clamp_ref(__dadd_rn((double)src[i], __dmul_rn((double)src[i - samples], mult)));
}

// one CTA per (block of `rate` samples, channel): fp64 sum -> average -> subtract + clamp
__global__ void __launch_bounds__(1024) center_kernel(float *__restrict__ d, size_t stride, size_t n, size_t rate) {
    float *row = d + (size_t)blockIdx.y * stride;
    __shared__ double part[32];
    __shared__ double s_avg;
    for (size_t b = blockIdx.x; b * rate < n; b += gridDim.x) {
        const size_t i0 = b * rate, l = n - i0 < rate ? n - i0 : rate;
        double s = 0.0;
        {
            const float *blk = row + i0;
            const size_t mis = ((uintptr_t)blk >> 2) & 3, head = mis ? (4 - mis < l ? 4 - mis : l) : 0;
            const size_t nvec = (l - head) / 4, tail0 = head + nvec * 4;
            const float4 *body = reinterpret_cast<const float4 *>(blk + head);
            for (size_t t = threadIdx.x; t < nvec; t += blockDim.x) {
                const float4 v = body[t];
                s += ((double)v.x + (double)v.y) + ((double)v.z + (double)v.w);
            }
            if (threadIdx.x < head) s += (double)blk[threadIdx.x];
            if (tail0 + threadIdx.x < l) s += (double)blk[tail0 + threadIdx.x];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += part[w];
            s_avg = t / (double)l;                                      // A:3473
        }
        __syncthreads();
        const double avg = s_avg;
        {
            float *blk = row + i0;
            const size_t mis = ((uintptr_t)blk >> 2) & 3, head = mis ? (4 - mis < l ? 4 - mis : l) : 0;
            const size_t nvec = (l - head) / 4, tail0 = head + nvec * 4;
            float4 *body = reinterpret_cast<float4 *>(blk + head);
            for (size_t t = threadIdx.x; t < nvec; t += blockDim.x) {
                float4 v = body[t];
                v.x = (float)clamp_ref((double)v.x - avg); v.y = (float)clamp_ref((double)v.y - avg);
                v.z = (float)clamp_ref((double)v.z - avg); v.w = (float)clamp_ref((double)v.w - avg);
                stg_stream(body + t, v);
            }
            if (threadIdx.x < head) blk[threadIdx.x] = (float)clamp_ref((double)blk[threadIdx.x] - avg);
            if (tail0 + threadIdx.x < l) blk[tail0 + threadIdx.x] = (float)clamp_ref((double)blk[tail0 + threadIdx.x] - avg);
        }
        __syncthreads();
    }
}

int nil_arith() { return aukit_fail("attempt to perform arithmetic on a nil value (field '?')"); }

}  // namespace

extern "C" int aukit_cuda_dev_invert(aukit_ctx *ctx, float *d, size_t stride, int channels, size_t n) {
    if (!ctx) return aukit_fail("aukit_cuda: null context");
    if (n == 0 || channels < 1) return 0;
    dim3 grid(aukit_grid(n, 256 * 4, (size_t)ctx->num_sms * 8), channels);
    invert_kernel<<<grid, 256, 0, ctx->stream>>>(d, stride, n);
    ctx->launches++;
    return aukit_cuda_check(cudaGetLastError(), "invert_kernel launch");
}

extern "C" int aukit_cuda_dev_fade(aukit_ctx *ctx, float *d, size_t stride, int channels, size_t n, double sampleRate,
                                   double startTime, double startAmplitude, double endTime, double endAmplitude) {
    if (!ctx) return aukit_fail("aukit_cuda: null context");
    if (startAmplitude == 1.0 && endAmplitude == 1.0) return 0;                       // A:3398
    if (channels < 1) return 0;
    const double start = startTime * sampleRate, last = endTime * sampleRate;          // for i = start, endTime * rate
    if (!(last >= start)) return 0;                                                    // empty loop
    const double cnt = floor(last - start) + 1.0;
    // ch[i] must exist for every i the loop visits: an integer start inside [1, #ch] and the last index too
    if (start != floor(start) || start < 1.0 || start + cnt - 1.0 > (double)n) return nil_arith();
    const double m = (endAmplitude - startAmplitude) / ((endTime - startTime) * sampleRate);   // A:3402
    const size_t count = (size_t)cnt;
    dim3 grid(aukit_grid(count, 256 * 4, (size_t)ctx->num_sms * 8), channels);
    fade_kernel<<<grid, 256, 0, ctx->stream>>>(d, stride, (size_t)start - 1, count, m, startAmplitude);
    ctx->launches++;
    return aukit_cuda_check(cudaGetLastError(), "fade_kernel launch");
}

extern "C" int aukit_cuda_dev_delay(aukit_ctx *ctx, float *d, size_t stride, int channels, size_t n, double sampleRate,
                                    double delay, double multiplier) {
    if (!ctx) return aukit_fail("aukit_cuda: null context");
    if (channels < 1 || n == 0) return 0;
    const double sd = floor(delay * sampleRate);                                       // A:3505
    if (sd < 0.0) return nil_arith();                                                  // original[i - samples] past the end
    if (sd >= (double)n) return 0;                                                     // for i = samples + 1, #o: empty
    const size_t samples = (size_t)sd;
    void *copy = nullptr;
    const size_t bytes = ((size_t)(channels - 1) * stride + n) * sizeof(float);
    if (aukit_dev_alloc(ctx, bytes, &copy)) return -1;
    int rc = aukit_cuda_check(cudaMemcpyAsync(copy, d, bytes, cudaMemcpyDeviceToDevice, ctx->stream), "D2D");
    if (!rc) {
        dim3 grid(aukit_grid(n - samples, 256 * 4, (size_t)ctx->num_sms * 8), channels);
        delay_kernel<<<grid, 256, 0, ctx->stream>>>(d, static_cast<const float *>(copy), stride, n, samples, multiplier);
        ctx->launches++;
        rc = aukit_cuda_check(cudaGetLastError(), "delay_kernel launch");
    }
    aukit_dev_free(ctx, copy);
    return rc;
}

extern "C" int aukit_cuda_dev_center(aukit_ctx *ctx, float *d, size_t stride, int channels, size_t n, double sampleRate) {
    if (!ctx) return aukit_fail("aukit_cuda: null context");
    if (channels < 1 || n == 0) return 0;
    if (!(sampleRate > 0.0)) return aukit_fail("'for' step must be positive");         // for i = 0, #ch - 1, sampleRate
    // a fractional rate makes the second block start at a non-integer index: ch[i + j] is nil there
    if (sampleRate != floor(sampleRate) && sampleRate < (double)n) return nil_arith();
    const size_t rate = sampleRate >= (double)n ? n : (size_t)sampleRate;
    const size_t blocks = (n + rate - 1) / rate;
    dim3 grid(aukit_grid(blocks, 1, (size_t)ctx->num_sms * 8), channels);
    center_kernel<<<grid, 1024, 0, ctx->stream>>>(d, stride, n, rate);
    ctx->launches++;
    return aukit_cuda_check(cudaGetLastError(), "center_kernel launch");
}

// handle-level entry points (Audio objects)
extern "C" int aukit_cuda_invert(aukit_ctx *ctx, aukit_audio *a) {
    if (!ctx || !a) return aukit_fail("aukit_cuda: null argument");
    return aukit_cuda_dev_invert(ctx, a->data, a->stride, a->channels, a->frames);
}
extern "C" int aukit_cuda_fade(aukit_ctx *ctx, aukit_audio *a, double startTime, double startAmplitude, double endTime,
                               double endAmplitude) {
    if (!ctx || !a) return aukit_fail("aukit_cuda: null argument");
    return aukit_cuda_dev_fade(ctx, a->data, a->stride, a->channels, a->frames, a->sampleRate, startTime, startAmplitude, endTime,
                               endAmplitude);
}
extern "C" int aukit_cuda_delay(aukit_ctx *ctx, aukit_audio *a, double delay, double multiplier) {
    if (!ctx || !a) return aukit_fail("aukit_cuda: null argument");
    return aukit_cuda_dev_delay(ctx, a->data, a->stride, a->channels, a->frames, a->sampleRate, delay, multiplier);
}
extern "C" int aukit_cuda_center(aukit_ctx *ctx, aukit_audio *a) {
    if (!ctx || !a) return aukit_fail("aukit_cuda: null argument");
    return aukit_cuda_dev_center(ctx, a->data, a->stride, a->channels, a->frames, a->sampleRate);
}
