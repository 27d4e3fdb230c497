// pipeline_decim.cu -- K15: the fused pipeline (and Audio:resample) for ratios 1 / 2^k: 96 -> 48 kHz (BASELINE config 5),
// 48 -> 48 kHz, 192 -> 48 kHz ...
//
// With ratio = newRate / oldRate = 2^-k the reference's position x = (i - 1) / ratio + 1 (A:666) is an exact integer for
// every output, so `x % 1 == 0` always holds and the sample is COPIED, unclamped, whatever the interpolation mode
// (A:667): output frame o (0-based) is input frame o * M, M = 2^k.  No taps, no weights, no halo arithmetic -- the path
// is a strided gather, de-interleave and (for the fused chain) mono mix + normalize:
//   pass 1 (peak)   max |value| over the selected frames -> d_max (warp shuffles + one atomicMax per CTA)
//   pass 2 (apply)  clamp(value * peak / max) -> planar float32 rows (A:3444-3455)
//   raw             the value itself (standalone Audio:resample on planar float32)
// Every kernel here is HBM-bound: the input is touched once per pass (only the sectors of the selected frames), the
// output written once with full 128-byte lines per warp and channel.
#include "common.cuh"
#include "pipeline.cuh"
#include "sample_formats.cuh"

#include <math.h>

using namespace aukit_fmt;

namespace {

enum { DE_PEAK = 0, DE_APPLY = 1, DE_RAW = 2 };

struct decim_args {
    pipe_args a;
    unsigned long long M;         // input frames per output frame
    int fmt;                      // (B << 8) | (KIND << 4) | BE, packed input only
};

template <int B, int KIND, bool BE>
__device__ __forceinline__ float dconv(const uint8_t *p) {
    if (B == 2 && KIND == K_SIGNED) return s16_to_float((int)(int16_t)load_raw_aligned<2, BE>(p));
    return convert<B, KIND>(load_raw_aligned<B, BE>(p), nullptr);
}

__device__ __forceinline__ float dsample(const decim_args &d, const uint8_t *frame, int c) {
    switch (d.fmt) {
#define AUKIT_DS(BB, KK, EE) case ((BB << 8) | (KK << 4) | EE): return dconv<BB, KK, (EE != 0)>(frame + c * BB);
        AUKIT_DS(1, K_SIGNED, 0) AUKIT_DS(1, K_UNSIGNED, 0)
        AUKIT_DS(2, K_SIGNED, 0) AUKIT_DS(2, K_SIGNED, 1) AUKIT_DS(2, K_UNSIGNED, 0) AUKIT_DS(2, K_UNSIGNED, 1)
        AUKIT_DS(3, K_SIGNED, 0) AUKIT_DS(3, K_SIGNED, 1) AUKIT_DS(3, K_UNSIGNED, 0) AUKIT_DS(3, K_UNSIGNED, 1)
        AUKIT_DS(4, K_SIGNED, 0) AUKIT_DS(4, K_SIGNED, 1) AUKIT_DS(4, K_UNSIGNED, 0) AUKIT_DS(4, K_UNSIGNED, 1)
        AUKIT_DS(4, K_FLOAT, 0) AUKIT_DS(4, K_FLOAT, 1)
#undef AUKIT_DS
    default: return 0.f;
    }
}

__device__ __forceinline__ void block_max_to(float m, float *d_max, int channel_slots) {
    __shared__ float wm[32];
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = threadIdx.x < (blockDim.x >> 5) ? wm[threadIdx.x] : 0.0f;
        m = warp_max(m);
        if (threadIdx.x == 0) atomic_max_nonneg(d_max, m);
    }
}

// any format / channel count: one thread per output frame
template <int EPI>
__global__ void __launch_bounds__(256) decim_kernel(decim_args d) {
    const pipe_args &a = d.a;
    const int C = a.channels;
    float mult = 0.f;
    if (EPI == DE_APPLY) mult = (float)(a.peak / (double)a.d_max[0]);    // A:3444; max == 0: 0 * inf = NaN passes the clamp
    float m = 0.f;
    const size_t fb = a.planar_f32 ? 0 : (size_t)C * (size_t)(d.fmt >> 8);
    for (size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o < a.n_out; o += (size_t)gridDim.x * blockDim.x) {
        const unsigned long long g = (a.out_first + o) * d.M - a.in_first;   // frame o*M, relative to the window held
        float s = 0.f;
        for (int c = 0; c < C; c++) {
            const float v = a.planar_f32 ? reinterpret_cast<const float *>(a.in)[(size_t)c * a.in_stride + g]
                                         : dsample(d, a.in + (size_t)g * fb, c);
            if (a.mono) s += v;                                          // s = s + data[c][i], A:686
            else if (EPI == DE_PEAK) m = fmaxf(m, fabsf(v));
            else a.out[(size_t)c * a.out_stride + o] = EPI == DE_RAW ? v : clamp_ref(v * mult);
        }
        if (a.mono) {
            const float mv = a.cn_pow2 ? s * a.inv_cn : __fdiv_rn(s, (float)C);   // s / cn, A:687
            if (EPI == DE_PEAK) m = fmaxf(m, fabsf(mv));
            else a.out[o] = EPI == DE_RAW ? mv : clamp_ref(mv * mult);   // A:3455
        }
    }
    if (EPI == DE_PEAK) block_max_to(m, a.d_max, 1);
}

// float32 little-endian interleaved, 4 or 8 channels (config 5), not mono: a thread owns FOUR output frames 256 apart;
// all of its 128-bit loads are issued before anything is used, every store instruction of a warp is one 128-byte line
template <int C, int EPI>
__global__ void __launch_bounds__(256) decim_f32_kernel(decim_args d) {
    const pipe_args &a = d.a;
    constexpr int V = C / 4;                                             // uint4 per frame
    float mult = 0.f;
    if (EPI == DE_APPLY) mult = (float)(a.peak / (double)a.d_max[0]);
    float m = 0.f;
    const size_t tile = 1024;
    const size_t ntiles = (a.n_out + tile - 1) / tile;
    const uint4 *in = reinterpret_cast<const uint4 *>(a.in);
    const unsigned long long rel0 = a.out_first * d.M - a.in_first;
    for (size_t tl = blockIdx.x; tl < ntiles; tl += gridDim.x) {
        const size_t o0 = tl * tile + threadIdx.x;
        uint4 v[4][V];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const size_t o = o0 + 256 * k;
            if (o < a.n_out) {
                const uint4 *p = in + (size_t)(rel0 + (unsigned long long)o * d.M) * V;
#pragma unroll
                for (int q = 0; q < V; q++) v[k][q] = ldg_stream(p + q);
            } else {
#pragma unroll
                for (int q = 0; q < V; q++) v[k][q] = make_uint4(0, 0, 0, 0);
            }
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const size_t o = o0 + 256 * k;
            const float *f = reinterpret_cast<const float *>(&v[k][0]);
            if (EPI == DE_PEAK) {
#pragma unroll
                for (int c = 0; c < C; c++) m = fmaxf(m, fabsf(f[c]));
            } else if (o < a.n_out) {
#pragma unroll
                for (int c = 0; c < C; c++) {
                    const float r = clamp_ref(f[c] * mult);
                    asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(a.out + (size_t)c * a.out_stride + o), "f"(r) : "memory");
                }
            }
        }
    }
    if (EPI == DE_PEAK) block_max_to(m, a.d_max, 1);
}

}  // namespace

// Returns 1 when it handled the launch, 0 when not applicable, -1 on error.
int aukit_pipeline_decim_try(aukit_ctx *ctx, const pipe_args &a, const aukit_pipeline_desc *p, bool apply, long long M) {
    static const bool disabled = getenv("AUKIT_DISABLE_DECIM") && getenv("AUKIT_DISABLE_DECIM")[0] == '1';
    if (disabled || M < 1) return 0;
    decim_args d{};
    d.a = a;
    d.M = (unsigned long long)M;
    const int B = p->bitDepth / 8;
    const int kind = p->dataType == AUKIT_FLOAT ? K_FLOAT : (p->dataType == AUKIT_UNSIGNED ? K_UNSIGNED : K_SIGNED);
    d.fmt = (B << 8) | (kind << 4) | ((p->bigEndian && B > 1) ? 1 : 0);
    const int epi = a.raw_out ? DE_RAW : (apply ? DE_APPLY : DE_PEAK);
    const int C = a.channels;
    const bool f32le = !a.planar_f32 && d.fmt == ((4 << 8) | (K_FLOAT << 4)) && !a.mono && epi != DE_RAW &&
                       (C == 8 || C == 4) && ((uintptr_t)a.in & 15) == 0;
    if (f32le) {
        const size_t ntiles = (a.n_out + 1023) / 1024;
        const unsigned grid = aukit_grid(ntiles, 1, (size_t)ctx->num_sms * 8);
        if (C == 8) {
            if (epi == DE_APPLY) decim_f32_kernel<8, DE_APPLY><<<grid, 256, 0, ctx->stream>>>(d);
            else decim_f32_kernel<8, DE_PEAK><<<grid, 256, 0, ctx->stream>>>(d);
        } else {
            if (epi == DE_APPLY) decim_f32_kernel<4, DE_APPLY><<<grid, 256, 0, ctx->stream>>>(d);
            else decim_f32_kernel<4, DE_PEAK><<<grid, 256, 0, ctx->stream>>>(d);
        }
    } else {
        const unsigned grid = aukit_grid(a.n_out, 256, (size_t)ctx->num_sms * 8 * 4);
        if (epi == DE_RAW) decim_kernel<DE_RAW><<<grid, 256, 0, ctx->stream>>>(d);
        else if (epi == DE_APPLY) decim_kernel<DE_APPLY><<<grid, 256, 0, ctx->stream>>>(d);
        else decim_kernel<DE_PEAK><<<grid, 256, 0, ctx->stream>>>(d);
    }
    ctx->launches++;
    return aukit_cuda_check(cudaGetLastError(), "decim_kernel launch") ? -1 : 1;
}
