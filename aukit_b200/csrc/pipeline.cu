// pipeline.cu -- K10: fused unpack -> resample -> [mono] -> normalize (auplay.lua:12-27).
//
// Two launches around the abs-max barrier (the max over data >> L2 is a true barrier):
//   pass 1 (peak)  reads the packed PCM once, computes every output value in registers and
//                  reduces |value| to one float (warp shuffles + one atomicMax per CTA);
//   pass 2 (apply) re-reads the packed PCM, recomputes the same values, multiplies by
//                  peak / max, clamps and writes float32.
// Algorithmic HBM traffic = 2 * B_in + B_out; no f32 intermediate is ever materialised.
//
// This file holds the GENERIC path: any sample format / channel count / ratio, one thread per
// output frame, fp64 position exactly as A:666 (see resample.cu), taps loaded straight from
// the packed interleaved input (L1 serves the 4x tap overlap between neighbouring outputs).
// The rational-ratio fast path for 16-bit input lives in pipeline_poly.cu.
#include "common.cuh"
#include "sample_formats.cuh"
#include "pipeline.cuh"

#include <math.h>
#include <stdlib.h>

using namespace aukit_fmt;

namespace {

template <int B, int KIND, bool BE, int MODE, bool APPLY>
__global__ void __launch_bounds__(256) pipeline_kernel(pipe_args a) {
    __shared__ float lut[(KIND == K_ALAW || KIND == K_ULAW) ? 256 : 1];
    __shared__ float wm[8];
    float mult = 0.f;
    if (APPLY) mult = (float)(a.peak / (double)a.d_max[0]);            // A:3444
    float m = 0.f;
    const int C = a.channels;
    const size_t fstride = (size_t)C * B;
    for (size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o < a.n_out;
         o += (size_t)gridDim.x * blockDim.x) {
        const unsigned long long i0 = a.out_first + o;
        const double x = __dadd_rn(__ddiv_rn((double)i0, a.ratio), 1.0);   // A:666
        const double fl = floor(x);
        const bool hit = (x == fl);
        const long long f = (long long)fl;
        const double t = x - fl;
        float w0 = 0.f, w1 = 0.f, w2 = 0.f, w3 = 0.f;
        const float fx = (float)t;
        if (MODE == AUKIT_INTERP_CUBIC) {
            const double t2 = t * t, t3 = t2 * t;
            w0 = (float)(-0.5 * t3 + t2 - 0.5 * t);
            w1 = (float)(1.5 * t3 - 2.5 * t2 + 1.0);
            w2 = (float)(-1.5 * t3 + 2.0 * t2 + 0.5 * t);
            w3 = (float)(0.5 * t3 - 0.5 * t2);
        }
        const long long n = (long long)a.n_total;
        const uint8_t *p1p = a.in + (size_t)(f - 1 - (long long)a.in_first) * fstride;
        float s = 0.f;
        for (int c = 0; c < C; c++) {
            const uint8_t *q = p1p + (size_t)c * B;
            const float p1 = convert<B, KIND>(load_raw_aligned<B, BE>(q), lut);
            float v;
            if (hit) v = p1;
            else if (MODE == AUKIT_INTERP_NONE) v = clamp_ref(p1);
            else if (MODE == AUKIT_INTERP_LINEAR) {
                const float p2 = (f + 1 <= n) ? convert<B, KIND>(load_raw_aligned<B, BE>(q + fstride), lut) : p1;
                v = clamp_ref(__fmaf_rn(p2 - p1, fx, p1));
            } else {
                const float p0 = (f - 1 >= 1) ? convert<B, KIND>(load_raw_aligned<B, BE>(q - fstride), lut) : p1;
                const float p2 = (f + 1 <= n) ? convert<B, KIND>(load_raw_aligned<B, BE>(q + fstride), lut) : p1;
                const float p3 = (f + 2 <= n) ? convert<B, KIND>(load_raw_aligned<B, BE>(q + 2 * fstride), lut) : p2;
                v = clamp_ref(__fmaf_rn(w3, p3, __fmaf_rn(w2, p2, __fmaf_rn(w1, p1, w0 * p0))));
            }
            if (a.mono) s += v;                                         // s = s + data[c][i], A:686
            else if (APPLY) a.out[(size_t)c * a.out_stride + o] = clamp_ref(v * mult);
            else m = fmaxf(m, fabsf(v));
        }
        if (a.mono) {
            const float mv = a.cn_pow2 ? s * a.inv_cn : __fdiv_rn(s, (float)C);   // s / cn, A:687
            if (APPLY) a.out[o] = clamp_ref(mv * mult);                 // A:3455
            else m = fmaxf(m, fabsf(mv));
        }
    }
    if (!APPLY) {
        m = warp_max(m);
        if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = m;
        __syncthreads();
        if (threadIdx.x < 32) {
            m = threadIdx.x < (blockDim.x >> 5) ? wm[threadIdx.x] : 0.0f;
            m = warp_max(m);
            if (threadIdx.x == 0) atomic_max_nonneg(a.d_max, m);
        }
    }
}

// K16 -- wide float frames: 32-bit float little-endian input with 4 or 8 interleaved channels (BASELINE config 5':
// 96 -> 44.1 kHz, 8 channels).  A frame is 16 / 32 bytes, so the four taps of an output frame are 64 / 128 CONTIGUOUS,
// 16-byte aligned bytes: one thread per output frame fetches them with 128-bit loads straight from global memory
// (neighbouring outputs overlap by one or two frames: L1 serves that), evaluates the reference's fp64 position once
// for all channels (exact at ANY position -- no polyphase position modes, no shared memory, no barriers), blends,
// and writes one coalesced 128-byte line per warp and channel.
template <int C, int MODE, bool MONO, bool APPLY>
__global__ void __launch_bounds__(256) wide_f32_kernel(pipe_args a) {
    constexpr int V = C / 4;
    __shared__ float wm[8];
    // apply pass, all channels kept: a CTA iteration produces 256 CONSECUTIVE frames, i.e. 1 KB of every channel row.  They are
    // staged (lane t writes float t: conflict-free) and leave as one bulk async store per channel (TMA, cp.async.bulk
    // shared -> global, double-buffered) instead of one 4-byte STG per thread and channel (what took K5 from 0.77 to 0.92).
    constexpr bool BULK = APPLY && !MONO;
    __shared__ __align__(128) float ost[BULK ? 2 : 1][BULK ? C : 1][BULK ? 256 : 1];
    const bool bulk_ok = BULK && (((uintptr_t)a.out & 15) == 0) && ((a.out_stride & 3) == 0) && blockDim.x == 256;
    float mult = 0.f;
    if (APPLY) mult = (float)(a.peak / (double)a.d_max[0]);            // A:3444
    float m = 0.f;
    const uint4 *in = reinterpret_cast<const uint4 *>(a.in);
    const long long n = (long long)a.n_total, first = (long long)a.in_first;
    int it = 0;
    for (size_t ob = (size_t)blockIdx.x * blockDim.x; ob < a.n_out; ob += (size_t)gridDim.x * blockDim.x, it++) {
        const size_t o = ob + threadIdx.x;
        const bool staged = bulk_ok && ob + 256 <= a.n_out;            // a whole group of 256 frames (uniform over the CTA)
        if (o < a.n_out) {
        const unsigned long long i0 = a.out_first + o;
        // A:666: (i - 1) / ratio + 1 with the correctly rounded quotient -- by the IEEE division, or by the 3-operation FMA
        // sequence where the host proved it equal for every index of the call (a third of this kernel's instructions)
        const double nd = (double)i0;
        double qd;
        if (a.qfma_ok) {
            const double q0 = __dmul_rn(nd, a.y);
            qd = __fma_rn(__fma_rn(-q0, a.ratio, nd), a.y, q0);
        } else {
            qd = __ddiv_rn(nd, a.ratio);
        }
        const double x = __dadd_rn(qd, 1.0);
        const double fl = floor(x);
        const bool hit = (x == fl);
        const long long f = (long long)fl;                             // 1-based index of p1
        const double t = x - fl;
        float w0 = 0.f, w1 = 0.f, w2 = 0.f, w3 = 0.f;
        const float fx = (float)t;
        if (MODE == AUKIT_INTERP_CUBIC) {
            const double t2 = t * t, t3 = t2 * t;
            w0 = (float)(-0.5 * t3 + t2 - 0.5 * t);
            w1 = (float)(1.5 * t3 - 2.5 * t2 + 1.0);
            w2 = (float)(-1.5 * t3 + 2.0 * t2 + 0.5 * t);
            w3 = (float)(0.5 * t3 - 0.5 * t2);
        }
        // nil neighbours == clamped index (A:259, A:264)
        const long long g1 = f - 1 - first;
        const long long g0 = (f - 1 >= 1 ? f - 2 : f - 1) - first;
        const long long g2 = (f + 1 <= n ? f : f - 1) - first;
        const long long g3 = (f + 2 <= n ? f + 1 : (f + 1 <= n ? f : f - 1)) - first;
        uint4 q1[V], q0[V], q2[V], q3[V];
#pragma unroll
        for (int k = 0; k < V; k++) {
            q1[k] = __ldg(in + (size_t)g1 * V + k);
            if (MODE == AUKIT_INTERP_CUBIC) { q0[k] = __ldg(in + (size_t)g0 * V + k); q3[k] = __ldg(in + (size_t)g3 * V + k); }
            if (MODE != AUKIT_INTERP_NONE) q2[k] = __ldg(in + (size_t)g2 * V + k);
        }
        const float *p1 = reinterpret_cast<const float *>(q1), *p0 = reinterpret_cast<const float *>(q0);
        const float *p2 = reinterpret_cast<const float *>(q2), *p3 = reinterpret_cast<const float *>(q3);
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < C; c++) {
            float v;
            if (hit) v = p1[c];                                        // copied unclamped, A:667
            else if (MODE == AUKIT_INTERP_NONE) v = clamp_ref(p1[c]);
            else if (MODE == AUKIT_INTERP_LINEAR) v = clamp_ref(__fmaf_rn(p2[c] - p1[c], fx, p1[c]));
            else v = clamp_ref(__fmaf_rn(w3, p3[c], __fmaf_rn(w2, p2[c], __fmaf_rn(w1, p1[c], w0 * p0[c]))));
            if (MONO) s += v;                                          // A:686
            else if (APPLY) {
                if (BULK && staged) ost[BULK ? (it & 1) : 0][BULK ? c : 0][BULK ? threadIdx.x : 0] = clamp_ref(v * mult);
                else a.out[(size_t)c * a.out_stride + o] = clamp_ref(v * mult);
            }
            else m = fmaxf(m, fabsf(v));
        }
        if (MONO) {
            const float mv = s * a.inv_cn;                             // C is a power of two: s / cn exactly (A:687)
            if (APPLY) a.out[o] = clamp_ref(mv * mult);
            else m = fmaxf(m, fabsf(mv));
        }
        }
        if (BULK && bulk_ok) {
            if (staged) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // my staging writes before the copy engine's reads
            // the rows staged one iteration ago must have been read before the next iteration overwrites them
            if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncthreads();
            if (threadIdx.x == 0 && staged) {
                for (int c = 0; c < C; c++)
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(a.out + (size_t)c * a.out_stride + ob),
                                 "r"((uint32_t)__cvta_generic_to_shared(&ost[BULK ? (it & 1) : 0][BULK ? c : 0][0])), "r"(1024u) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
    }
    if (BULK && bulk_ok && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (!APPLY) {
        m = warp_max(m);
        if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = m;
        __syncthreads();
        if (threadIdx.x < 32) {
            m = threadIdx.x < (blockDim.x >> 5) ? wm[threadIdx.x] : 0.0f;
            m = warp_max(m);
            if (threadIdx.x == 0) atomic_max_nonneg(a.d_max, m);
        }
    }
}

template <int C, bool APPLY>
int launch_wide_c(aukit_ctx *ctx, const pipe_args &a, int interp) {
    const unsigned grid = aukit_grid(a.n_out, 256, (size_t)ctx->num_sms * 8 * 4);
#define AUKIT_WIDE(MODE)                                                                      \
    if (a.mono) wide_f32_kernel<C, MODE, true, APPLY><<<grid, 256, 0, ctx->stream>>>(a);      \
    else wide_f32_kernel<C, MODE, false, APPLY><<<grid, 256, 0, ctx->stream>>>(a)
    switch (interp) {
    case AUKIT_INTERP_NONE: AUKIT_WIDE(AUKIT_INTERP_NONE); break;
    case AUKIT_INTERP_LINEAR: AUKIT_WIDE(AUKIT_INTERP_LINEAR); break;
    default: AUKIT_WIDE(AUKIT_INTERP_CUBIC); break;
    }
#undef AUKIT_WIDE
    ctx->launches++;
    return aukit_cuda_check(cudaGetLastError(), "wide_f32_kernel launch");
}

// 1: handled, 0: not this shape, -1: error.  The choice depends on the format only (never on the position), so every
// time shard of a buffer takes the same kernel and shards stay bit-identical to a single pass.
template <bool APPLY>
int wide_try(aukit_ctx *ctx, const pipe_args &a, const aukit_pipeline_desc *p) {
    static const bool disabled = getenv("AUKIT_DISABLE_WIDE") && getenv("AUKIT_DISABLE_WIDE")[0] == '1';
    if (disabled) return 0;
    if (p->dataType != AUKIT_FLOAT || p->bitDepth != 32 || p->bigEndian) return 0;
    if (p->channels != 4 && p->channels != 8) return 0;
    if (((uintptr_t)a.in & 15) != 0) return 0;
    int e2 = 0;
    if (frexp(a.ratio, &e2) == 0.5 && a.ratio <= 1.0) return 0;        // 1 / 2^k: the strided gather (K15) is the better kernel
    pipe_args b = a;
    b.y = 1.0 / a.ratio;
    // quotients up to the largest position of the WHOLE signal (not of this shard): every shard then takes the same arithmetic
    int kmax = 0;
    frexp((double)a.n_total + 4.0, &kmax);
    b.qfma_ok = aukit_quotient_fma_is_exact(a.ratio, kmax + 1) ? 1 : 0;
    const int rc = p->channels == 8 ? launch_wide_c<8, APPLY>(ctx, b, p->interpolation) : launch_wide_c<4, APPLY>(ctx, b, p->interpolation);
    return rc ? -1 : 1;
}

template <int B, int KIND, bool BE, bool APPLY>
int launch_mode(aukit_ctx *ctx, const pipe_args &a, int interp) {
    const int threads = 256;
    const unsigned grid = aukit_grid(a.n_out, threads, (size_t)ctx->num_sms * 8 * 8);
    switch (interp) {
    case AUKIT_INTERP_NONE: pipeline_kernel<B, KIND, BE, AUKIT_INTERP_NONE, APPLY><<<grid, threads, 0, ctx->stream>>>(a); break;
    case AUKIT_INTERP_LINEAR: pipeline_kernel<B, KIND, BE, AUKIT_INTERP_LINEAR, APPLY><<<grid, threads, 0, ctx->stream>>>(a); break;
    default: pipeline_kernel<B, KIND, BE, AUKIT_INTERP_CUBIC, APPLY><<<grid, threads, 0, ctx->stream>>>(a); break;
    }
    ctx->launches++;
    return aukit_cuda_check(cudaGetLastError(), "pipeline_kernel launch");
}

template <bool APPLY>
int launch_fmt(aukit_ctx *ctx, const pipe_args &a, const aukit_pipeline_desc *p) {
    const int B = p->bitDepth / 8;
    const bool be = p->bigEndian != 0 && B > 1;
#define AUKIT_PIPE(BB, KK) \
    return be ? launch_mode<BB, KK, (BB > 1), APPLY>(ctx, a, p->interpolation) : launch_mode<BB, KK, false, APPLY>(ctx, a, p->interpolation)
    if (p->dataType == AUKIT_FLOAT) { AUKIT_PIPE(4, K_FLOAT); }
    if (p->dataType == AUKIT_SIGNED) {
        switch (B) { case 1: AUKIT_PIPE(1, K_SIGNED); case 2: AUKIT_PIPE(2, K_SIGNED);
                     case 3: AUKIT_PIPE(3, K_SIGNED); default: AUKIT_PIPE(4, K_SIGNED); }
    }
    switch (B) { case 1: AUKIT_PIPE(1, K_UNSIGNED); case 2: AUKIT_PIPE(2, K_UNSIGNED);
                 case 3: AUKIT_PIPE(3, K_UNSIGNED); default: AUKIT_PIPE(4, K_UNSIGNED); }
#undef AUKIT_PIPE
}

}  // namespace

static int validate(aukit_ctx *ctx, const aukit_pipeline_desc *p, const void *d_in, pipe_args *a) {
    if (!ctx || !p) return aukit_fail("aukit_cuda: null argument");
    if (p->bitDepth != 8 && p->bitDepth != 16 && p->bitDepth != 24 && p->bitDepth != 32)
        return aukit_fail("bad argument #2 (invalid bit depth)");
    if (p->dataType != AUKIT_SIGNED && p->dataType != AUKIT_UNSIGNED && p->dataType != AUKIT_FLOAT)
        return aukit_fail("bad argument #3 (invalid data type)");
    if (p->dataType == AUKIT_FLOAT && p->bitDepth != 32)
        return aukit_fail("bad argument #2 (float audio must have 32-bit depth)");
    if (p->channels < 1) return aukit_fail("aukit_cuda: channels < 1");
    if (p->interpolation == AUKIT_INTERP_SINC) return aukit_fail("aukit_cuda: sinc interpolation is available through Audio:resample only, not the fused chain");
    if (p->interpolation < 0 || p->interpolation > 2) return aukit_fail("bad argument #2 (invalid interpolation type)");
    const int B = p->bitDepth / 8;
    if ((B == 2 || B == 4) && (uintptr_t)d_in % B) return aukit_fail("aukit_cuda: packed input must be sample-aligned");
    const uint64_t total_out = aukit_resample_out_len(p->n_in_total, p->srcRate, p->dstRate);
    if (p->out_first + p->n_out > total_out) return aukit_fail("aukit_cuda: output range exceeds floor(n_in * ratio)");
    if (p->n_out) {
        uint64_t nf = 0, nc = 0;
        if (aukit_resample_window(p->n_in_total, p->srcRate, p->dstRate, p->interpolation, p->out_first, p->n_out, &nf, &nc))
            return -1;
        if (floor(aukit_resample_position(p->out_first + p->n_out, p->srcRate, p->dstRate)) > (double)p->n_in_total)
            return aukit_fail("aukit_cuda: position past the end of the input");
        if (nf < p->in_first || nf + nc > p->in_first + p->in_avail)
            return aukit_fail("aukit_cuda: input window does not cover the frames this shard needs (halo missing)");
    }
    a->in = static_cast<const uint8_t *>(d_in);
    a->channels = p->channels;
    a->n_total = p->n_in_total;
    a->in_first = p->in_first;
    a->in_avail = p->in_avail;
    a->ratio = p->dstRate / p->srcRate;
    a->out_first = p->out_first;
    a->n_out = p->n_out;
    a->mono = p->mono != 0;
    a->inv_cn = 1.0f / (float)p->channels;
    a->cn_pow2 = (p->channels & (p->channels - 1)) == 0;
    a->hint = ctx->d_hint;
    a->epoch = ctx->epoch;
    return 0;
}

extern "C" int aukit_cuda_dev_pipeline_peak(aukit_ctx *ctx, const aukit_pipeline_desc *p, const void *d_in,
                                            float *d_max) {
    pipe_args a{};
    if (ctx) ctx->epoch = ctx->epoch >= 0x7FFFFFF0 ? 1 : ctx->epoch + 1;   // a new call: hints of earlier calls no longer apply
    if (validate(ctx, p, d_in, &a)) return -1;
    if (a.n_out == 0) return 0;
    a.d_max = d_max;
    const int w = wide_try<false>(ctx, a, p);
    if (w != 0) return w < 0 ? -1 : 0;
    const int r = aukit_pipeline_poly_try(ctx, a, p, false);
    if (r != 0) return r < 0 ? -1 : 0;
    return launch_fmt<false>(ctx, a, p);
}

extern "C" int aukit_cuda_dev_pipeline_apply(aukit_ctx *ctx, const aukit_pipeline_desc *p, const void *d_in,
                                             double peakAmplitude, const float *d_max, float *d_out,
                                             size_t out_stride) {
    pipe_args a{};
    if (validate(ctx, p, d_in, &a)) return -1;
    if (a.n_out == 0) return 0;
    if (!p->mono && p->channels > 1 && out_stride < p->n_out) return aukit_fail("aukit_cuda: out_stride < n_out");
    a.d_max = const_cast<float *>(d_max);
    a.peak = peakAmplitude;
    a.out = d_out;
    a.out_stride = out_stride;
    const int w = wide_try<true>(ctx, a, p);
    if (w != 0) return w < 0 ? -1 : 0;
    const int r = aukit_pipeline_poly_try(ctx, a, p, true);
    if (r != 0) return r < 0 ? -1 : 0;
    return launch_fmt<true>(ctx, a, p);
}
