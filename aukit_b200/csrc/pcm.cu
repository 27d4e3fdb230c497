// pcm.cu -- K1 pcm_unpack + K2 g711_decode.
//
// Replaces aukit.pcm (A:1049-1171) and aukit.g711 (A:1361-1384): packed sample bytes ->
// planar float32 [C][N].  HBM-bound streaming kernels: each thread owns one contiguous,
// 16-byte-aligned chunk of the packed input (FPT frames x C channels), loads it with
// 128-bit streaming loads, byte-swaps / sign-extends / de-interleaves in registers (static
// PRMT selectors after unrolling) and writes one float4 per channel per 4 frames.
// Integer scaling reproduces A:1133 / A:1152 bit-exactly (see common.cuh::s16_to_float and
// tests/test_scaling_exact.py); 32-bit integer input divides in fp64 because float(s) is
// inexact there.  G.711 and 8-bit PCM use exact float bit constructions (sample_formats.cuh::convert8).
#include "common.cuh"
#include "sample_formats.cuh"

namespace {
using namespace aukit_fmt;

template <int B, int KIND>
__device__ __forceinline__ float conv_lut(uint32_t raw, const float *lut) {
    if (B == 1) return convert8<KIND>(raw & 0xFFu);
    return convert<B, KIND>(raw, lut);
}

// C > 0: interleaved with compile-time channel count (vector path).
// C == 0: runtime channel count (per-sample loads, L1-served).
// Grid: x over chunks of frames; y over planar channel rows (planar layout runs as C == 1
// with a per-row base offset).
template <int B, int KIND, bool BE, int C>
__global__ void __launch_bounds__(256)
pcm_unpack_kernel(const uint8_t *__restrict__ in, float *__restrict__ out, size_t frames,
                  size_t out_stride, int channels_rt, size_t planar_row_bytes, int vec_ok) {
    // every format converts in registers (8-bit ones through convert8: no table)
    const float *lut = nullptr;
    const uint8_t *src = in + (size_t)blockIdx.y * planar_row_bytes;
    float *dst = out + (size_t)blockIdx.y * out_stride;
    constexpr int CC = C > 0 ? C : 1;
    constexpr int FPT = C > 0 ? frames_per_thread(B, CC) : 4;
    if (C == 0) {
        // runtime channel count: one thread per (4 frames, channel), channel fastest, so the lanes of a warp
        // read runs of C consecutive samples (full sectors) and every thread stores one float4
        const int nc = channels_rt;
        const size_t nquads = (frames + 3) / 4, items = nquads * (size_t)nc;
        for (size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x; id < items; id += (size_t)gridDim.x * blockDim.x) {
            const size_t q = id / (size_t)nc;
            const int c = (int)(id % (size_t)nc);
            const size_t f0 = q * 4;
            const uint8_t *p = src + (f0 * (size_t)nc + c) * B;
            const size_t fs = (size_t)nc * B;
            if (f0 + 4 <= frames) {
                float4 o;
                o.x = conv_lut<B, KIND>(load_raw_aligned_or_bytes<B, BE>(p, vec_ok), lut);
                o.y = conv_lut<B, KIND>(load_raw_aligned_or_bytes<B, BE>(p + fs, vec_ok), lut);
                o.z = conv_lut<B, KIND>(load_raw_aligned_or_bytes<B, BE>(p + 2 * fs, vec_ok), lut);
                o.w = conv_lut<B, KIND>(load_raw_aligned_or_bytes<B, BE>(p + 3 * fs, vec_ok), lut);
                float *d = dst + (size_t)c * out_stride + f0;
                if (vec_ok) stg_stream(reinterpret_cast<float4 *>(d), o);
                else { d[0] = o.x; d[1] = o.y; d[2] = o.z; d[3] = o.w; }
            } else {
                for (size_t f = f0; f < frames; f++)
                    dst[(size_t)c * out_stride + f] = conv_lut<B, KIND>(load_raw<B, BE>(src + (f * (size_t)nc + c) * B), lut);
            }
        }
        return;
    }
    const size_t nchunks = (frames + FPT - 1) / FPT;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < nchunks;
         t += (size_t)gridDim.x * blockDim.x) {
        const size_t f0 = t * FPT;
        if (vec_ok && f0 + FPT <= frames) {
            constexpr int WORDS = FPT * CC * B / 4;
            uint32_t w[WORDS + 1];
            const uint4 *p = reinterpret_cast<const uint4 *>(src + f0 * (size_t)(CC * B));
#pragma unroll
            for (int k = 0; k < WORDS / 4; k++) {
                uint4 v = ldg_stream(p + k);
                w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w;
            }
            w[WORDS] = 0;
#pragma unroll
            for (int c = 0; c < CC; c++) {
#pragma unroll
                for (int q = 0; q < FPT / 4; q++) {
                    float4 o;
                    o.x = conv_lut<B, KIND>(extract<B, BE>(w, ((4 * q + 0) * CC + c) * B), lut);
                    o.y = conv_lut<B, KIND>(extract<B, BE>(w, ((4 * q + 1) * CC + c) * B), lut);
                    o.z = conv_lut<B, KIND>(extract<B, BE>(w, ((4 * q + 2) * CC + c) * B), lut);
                    o.w = conv_lut<B, KIND>(extract<B, BE>(w, ((4 * q + 3) * CC + c) * B), lut);
                    stg_stream(reinterpret_cast<float4 *>(dst + (size_t)c * out_stride + f0 + 4 * q), o);
                }
            }
        } else {
            const size_t fend = f0 + FPT < frames ? f0 + FPT : frames;
            for (int c = 0; c < CC; c++)
                for (size_t f = f0; f < fend; f++)
                    dst[(size_t)c * out_stride + f] =
                        conv_lut<B, KIND>(load_raw<B, BE>(src + (f * (size_t)CC + c) * B), lut);
        }
    }
}

// Narrow frames (4 frames = 4 or 8 bytes: 8-bit mono / stereo, 16-bit mono): a 16-byte load per thread would make
// every thread store 2-4 consecutive float4 per channel, i.e. warp stores of half-filled sectors 32-64 bytes apart
// (the scattered-store pattern that capped these formats at ~70 %).  Instead a warp owns 32 * SUB groups of 4 frames
// and lane l takes groups l, l + 32, ...: every load instruction reads one contiguous 128/256-byte span and every
// store instruction writes one contiguous 512-byte span per channel.
template <int B, int KIND, bool BE, int C>
__global__ void __launch_bounds__(256)
pcm_unpack_narrow_kernel(const uint8_t *__restrict__ in, float *__restrict__ out, size_t frames, size_t out_stride,
                         size_t planar_row_bytes) {
    constexpr int SB = 4 * C * B;                   // bytes per group of 4 frames: 4, 8 (8/16-bit) or 12, 24 (24-bit)
    constexpr int SUB = SB <= 8 ? 16 / SB : 1;      // groups per lane
    static_assert(SB == 4 || SB == 8 || SB == 12 || SB == 24, "narrow frames only");
    const uint8_t *src = in + (size_t)blockIdx.y * planar_row_bytes;
    float *dst = out + (size_t)blockIdx.y * out_stride;
    const size_t ngroups = (frames + 3) / 4;
    const int lane = threadIdx.x & 31;
    const size_t warps = ((size_t)gridDim.x * blockDim.x) >> 5;
    const size_t full_groups = frames / 4;
    for (size_t wg = (((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * (32 * SUB); wg < ngroups; wg += warps * (32 * SUB)) {
        if (SUB > 1 && wg + 32 * SUB <= full_groups) {
            // interior chunk: issue every load of the lane first (SUB independent requests in flight), then convert
            uint32_t ww[SUB][SB / 4 + 1];
#pragma unroll
            for (int k = 0; k < SUB; k++) {
                const uint8_t *p = src + (wg + (size_t)k * 32 + lane) * SB;
                if (SB == 8) { const uint2 v = __ldg(reinterpret_cast<const uint2 *>(p)); ww[k][0] = v.x; ww[k][1] = v.y; }
                else ww[k][0] = __ldg(reinterpret_cast<const uint32_t *>(p));
                ww[k][SB / 4] = 0;
            }
#pragma unroll
            for (int k = 0; k < SUB; k++) {
                const size_t f0 = (wg + (size_t)k * 32 + lane) * 4;
#pragma unroll
                for (int c = 0; c < C; c++) {
                    float4 o;
                    o.x = conv_lut<B, KIND>(extract<B, BE>(ww[k], (0 * C + c) * B), nullptr);
                    o.y = conv_lut<B, KIND>(extract<B, BE>(ww[k], (1 * C + c) * B), nullptr);
                    o.z = conv_lut<B, KIND>(extract<B, BE>(ww[k], (2 * C + c) * B), nullptr);
                    o.w = conv_lut<B, KIND>(extract<B, BE>(ww[k], (3 * C + c) * B), nullptr);
                    stg_stream(reinterpret_cast<float4 *>(dst + (size_t)c * out_stride + f0), o);
                }
            }
            continue;
        }
#pragma unroll
        for (int k = 0; k < SUB; k++) {
            const size_t g = wg + (size_t)k * 32 + lane;
            if (g >= ngroups) continue;
            const size_t f0 = g * 4;
            if (f0 + 4 <= frames) {
                uint32_t w[SB / 4 + 1];
                if (SB % 8 == 0) {
#pragma unroll
                    for (int i = 0; i < SB / 8; i++) {
                        const uint2 v = __ldg(reinterpret_cast<const uint2 *>(src + g * SB) + i);
                        w[2 * i] = v.x; w[2 * i + 1] = v.y;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < SB / 4; i++) w[i] = __ldg(reinterpret_cast<const uint32_t *>(src + g * SB) + i);
                }
                w[SB / 4] = 0;
#pragma unroll
                for (int c = 0; c < C; c++) {
                    float4 o;
                    o.x = conv_lut<B, KIND>(extract<B, BE>(w, (0 * C + c) * B), nullptr);
                    o.y = conv_lut<B, KIND>(extract<B, BE>(w, (1 * C + c) * B), nullptr);
                    o.z = conv_lut<B, KIND>(extract<B, BE>(w, (2 * C + c) * B), nullptr);
                    o.w = conv_lut<B, KIND>(extract<B, BE>(w, (3 * C + c) * B), nullptr);
                    stg_stream(reinterpret_cast<float4 *>(dst + (size_t)c * out_stride + f0), o);
                }
            } else {
                for (int c = 0; c < C; c++)
                    for (size_t f = f0; f < frames; f++)
                        dst[(size_t)c * out_stride + f] = conv_lut<B, KIND>(load_raw<B, BE>(src + (f * (size_t)C + c) * B), nullptr);
            }
        }
    }
}

// Many channels with a frame size that is a multiple of 16 bytes (f32 / s32 with C % 4 == 0, s16 with C % 8 == 0,
// 8-bit with C % 16 == 0): one thread owns 4 frames x G = 16/B channels -- four 128-bit loads, a 4 x G transpose
// in registers, G float4 stores.  The per-sample-load path above spent its time in the LSU queue (lg_throttle):
// 1.25 memory instructions per sample against (4 + G) / (4 G) here.
template <int B, int KIND, bool BE>
__global__ void __launch_bounds__(256)
pcm_unpack_groups_kernel(const uint8_t *__restrict__ in, float *__restrict__ out, size_t frames, size_t out_stride, int nc) {
    constexpr int G = 16 / B;
    const int ngroups = nc / G;
    const size_t nquads = frames / 4, items = nquads * (size_t)ngroups;    // the ragged last frames go sample by sample
    const size_t fbytes = (size_t)nc * B;
    for (size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x; id < items; id += (size_t)gridDim.x * blockDim.x) {
        const size_t q = id / (size_t)ngroups;
        const int g = (int)(id % (size_t)ngroups);
        const uint8_t *p = in + q * 4 * fbytes + (size_t)g * 16;
        uint32_t w[4][5];
#pragma unroll
        for (int f = 0; f < 4; f++) {
            const uint4 v = ldg_stream(reinterpret_cast<const uint4 *>(p + f * fbytes));
            w[f][0] = v.x; w[f][1] = v.y; w[f][2] = v.z; w[f][3] = v.w; w[f][4] = 0;
        }
        float *d = out + (size_t)(g * G) * out_stride + q * 4;
#pragma unroll
        for (int j = 0; j < G; j++) {
            float4 o;
            o.x = conv_lut<B, KIND>(extract<B, BE>(w[0], j * B), nullptr);
            o.y = conv_lut<B, KIND>(extract<B, BE>(w[1], j * B), nullptr);
            o.z = conv_lut<B, KIND>(extract<B, BE>(w[2], j * B), nullptr);
            o.w = conv_lut<B, KIND>(extract<B, BE>(w[3], j * B), nullptr);
            stg_stream(reinterpret_cast<float4 *>(d + (size_t)j * out_stride), o);
        }
    }
    // ragged tail: frames % 4 leftover frames
    const size_t tail0 = nquads * 4, ntail = (frames - tail0) * (size_t)nc;
    for (size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x; id < ntail; id += (size_t)gridDim.x * blockDim.x) {
        const size_t f = tail0 + id / (size_t)nc;
        const int c = (int)(id % (size_t)nc);
        out[(size_t)c * out_stride + f] = conv_lut<B, KIND>(load_raw<B, BE>(in + (f * (size_t)nc + c) * B), nullptr);
    }
}

template <int B, int KIND, bool BE>
int launch_c(aukit_ctx *ctx, const uint8_t *d_in, float *d_out, size_t frames, size_t out_stride,
             int channels, bool interleaved) {
    if (frames == 0) return 0;
    const int vec_ok = ((uintptr_t)d_in % 16 == 0) && ((uintptr_t)d_out % 16 == 0) && (out_stride % 4 == 0);
    const int threads = 256;
    auto grid_for = [&](int fpt) {
        size_t chunks = (frames + fpt - 1) / fpt;
        return aukit_grid(chunks, threads, (size_t)ctx->num_sms * 8 * 16);
    };
    if (!interleaved || channels == 1) {
        // planar rows: row c starts at byte c*frames*B; vector path only if every row is 16B aligned
        const size_t row_bytes = frames * (size_t)B;
        const int vok = vec_ok && (channels == 1 || row_bytes % 16 == 0);
        dim3 grid(grid_for(frames_per_thread(B, 1)), channels);
        if constexpr (B <= 3) {
            if (vok) {
                dim3 ngrid(grid_for(B == 3 ? 4 : 16), channels);
                pcm_unpack_narrow_kernel<B, KIND, BE, 1><<<ngrid, threads, 0, ctx->stream>>>(d_in, d_out, frames, out_stride, row_bytes);
                ctx->launches++;
                return aukit_cuda_check(cudaGetLastError(), "pcm_unpack_narrow launch");
            }
        }
        pcm_unpack_kernel<B, KIND, BE, 1><<<grid, threads, 0, ctx->stream>>>(d_in, d_out, frames, out_stride, 1,
                                                                              row_bytes, vok);
    } else if (channels == 2) {
        if constexpr (B == 1 || B == 3) {
            if (vec_ok) {
                pcm_unpack_narrow_kernel<B, KIND, BE, 2><<<grid_for(B == 3 ? 4 : 8), threads, 0, ctx->stream>>>(d_in, d_out, frames, out_stride, 0);
                ctx->launches++;
                return aukit_cuda_check(cudaGetLastError(), "pcm_unpack_narrow launch");
            }
        }
        pcm_unpack_kernel<B, KIND, BE, 2><<<grid_for(frames_per_thread(B, 2)), threads, 0, ctx->stream>>>(
            d_in, d_out, frames, out_stride, 2, 0, vec_ok);
    } else if (B != 3 && vec_ok && ((size_t)channels * B) % 16 == 0) {
        const size_t items = frames / 4 * (size_t)(channels / (16 / B));
        pcm_unpack_groups_kernel<B, KIND, BE><<<aukit_grid(items ? items : 1, threads, (size_t)ctx->num_sms * 8 * 16), threads, 0, ctx->stream>>>(
            d_in, d_out, frames, out_stride, channels);
    } else {
        // vec_ok here = samples are naturally aligned (natural-width loads) and float4 stores are legal
        const int vok = ((uintptr_t)d_in % B == 0 || B == 3) && ((uintptr_t)d_out % 16 == 0) && (out_stride % 4 == 0);
        const size_t items = (frames + 3) / 4 * (size_t)channels;
        pcm_unpack_kernel<B, KIND, BE, 0><<<aukit_grid(items, threads, (size_t)ctx->num_sms * 8 * 16), threads, 0, ctx->stream>>>(
            d_in, d_out, frames, out_stride, channels, 0, vok);
    }
    ctx->launches++;
    return aukit_cuda_check(cudaGetLastError(), "pcm_unpack launch");
}

template <int B, int KIND>
int launch_e(aukit_ctx *ctx, const uint8_t *d_in, float *d_out, size_t frames, size_t out_stride,
             int channels, bool interleaved, bool be) {
    if (B == 1 || !be) return launch_c<B, KIND, false>(ctx, d_in, d_out, frames, out_stride, channels, interleaved);
    return launch_c<B, KIND, (B > 1)>(ctx, d_in, d_out, frames, out_stride, channels, interleaved);
}

}  // namespace

extern "C" int aukit_cuda_dev_pcm(aukit_ctx *ctx, const void *d_in, size_t nbytes, int bitDepth,
                                  int dataType, int channels, int interleaved, int bigEndian,
                                  float *d_out, size_t out_stride) {
    if (!ctx) return aukit_fail("aukit_cuda: null context");
    if (bitDepth != 8 && bitDepth != 16 && bitDepth != 24 && bitDepth != 32)
        return aukit_fail("bad argument #2 (invalid bit depth)");                            // A:1058
    if (dataType != AUKIT_SIGNED && dataType != AUKIT_UNSIGNED && dataType != AUKIT_FLOAT)
        return aukit_fail("bad argument #3 (invalid data type)");                            // A:1059
    if (dataType == AUKIT_FLOAT && bitDepth != 32)
        return aukit_fail("bad argument #2 (float audio must have 32-bit depth)");           // A:1060
    if (channels < 1) return aukit_fail("number outside of range (expected %d to be at least 1)", channels);
    const size_t B = (size_t)bitDepth / 8;
    if (nbytes % (B * (size_t)channels) != 0)
        return aukit_fail("bad argument #1 (uneven amount of data per channel)");           // A:1064
    const size_t frames = nbytes / B / (size_t)channels;                                     // A:1065
    if (channels > 1 && out_stride < frames) return aukit_fail("aukit_cuda: out_stride < frames");
    const uint8_t *in = static_cast<const uint8_t *>(d_in);
    const bool il = interleaved != 0, be = bigEndian != 0;
#define AUKIT_PCM_CASE(BB, KK) return launch_e<BB, KK>(ctx, in, d_out, frames, out_stride, channels, il, be)
    if (dataType == AUKIT_FLOAT) AUKIT_PCM_CASE(4, K_FLOAT);
    if (dataType == AUKIT_SIGNED) {
        switch (B) { case 1: AUKIT_PCM_CASE(1, K_SIGNED); case 2: AUKIT_PCM_CASE(2, K_SIGNED);
                     case 3: AUKIT_PCM_CASE(3, K_SIGNED); default: AUKIT_PCM_CASE(4, K_SIGNED); }
    }
    switch (B) { case 1: AUKIT_PCM_CASE(1, K_UNSIGNED); case 2: AUKIT_PCM_CASE(2, K_UNSIGNED);
                 case 3: AUKIT_PCM_CASE(3, K_UNSIGNED); default: AUKIT_PCM_CASE(4, K_UNSIGNED); }
#undef AUKIT_PCM_CASE
}

extern "C" int aukit_cuda_dev_g711(aukit_ctx *ctx, const void *d_in, size_t nbytes, int ulaw,
                                   int channels, float *d_out, size_t out_stride) {
    if (!ctx) return aukit_fail("aukit_cuda: null context");
    if (channels < 1) return aukit_fail("aukit_cuda: channels < 1");
    const uint8_t *in = static_cast<const uint8_t *>(d_in);
    const size_t full = nbytes / (size_t)channels, rem = nbytes % (size_t)channels;
    if (channels > 1 && out_stride < full + (rem ? 1 : 0)) return aukit_fail("aukit_cuda: out_stride < frames");
    int rc = ulaw ? launch_c<1, K_ULAW, false>(ctx, in, d_out, full, out_stride, channels, true)
                  : launch_c<1, K_ALAW, false>(ctx, in, d_out, full, out_stride, channels, true);
    if (rc) return rc;
    if (rem) {
        // ragged tail (A:1379): the last partial frame feeds channels 0..rem-1 one more sample.
        // Decode it as `rem` one-frame planar rows written at frame index `full`.
        rc = ulaw ? launch_c<1, K_ULAW, false>(ctx, in + full * (size_t)channels, d_out + full, 1, out_stride,
                                               (int)rem, false)
                  : launch_c<1, K_ALAW, false>(ctx, in + full * (size_t)channels, d_out + full, 1, out_stride,
                                               (int)rem, false);
    }
    return rc;
}
