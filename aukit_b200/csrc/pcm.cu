// pcm.cu -- K1 pcm_unpack + K2 g711_decode.
//
// Replaces aukit.pcm (A:1049-1171) and aukit.g711 (A:1361-1384): packed sample bytes ->
// planar float32 [C][N].  HBM-bound streaming kernels: each thread owns one contiguous,
// 16-byte-aligned chunk of the packed input (FPT frames x C channels), loads it with
// 128-bit streaming loads, byte-swaps / sign-extends / de-interleaves in registers (static
// PRMT selectors after unrolling) and writes one float4 per channel per 4 frames.
// Integer scaling reproduces A:1133 / A:1152 bit-exactly (see common.cuh::s16_to_float and
// tests/test_scaling_exact.py); 32-bit integer input divides in fp64 because float(s) is
// inexact there.  G.711 uses a 256-entry shared-memory table built from A:1374-1379.
#include "common.cuh"
#include "sample_formats.cuh"

namespace {
using namespace aukit_fmt;

template <int B, int KIND>
__device__ __forceinline__ float conv_lut(uint32_t raw, const float *lut) {
    if (B == 1) return lut[raw & 0xFFu];
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
    // 8-bit formats (G.711 and 8-bit PCM) decode through a 256-entry shared table built with the exact
    // per-sample formula (A:1374-1379 / A:1133 / A:1152); wider formats compute in registers
    constexpr bool USE_LUT = (B == 1);
    __shared__ float lut[USE_LUT ? 256 : 1];
    if (USE_LUT) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x)
            lut[i] = (KIND == K_ALAW || KIND == K_ULAW) ? g711_value(i, KIND == K_ULAW) : convert<B, KIND>((uint32_t)i, nullptr);
        __syncthreads();
    }
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
        pcm_unpack_kernel<B, KIND, BE, 1><<<grid, threads, 0, ctx->stream>>>(d_in, d_out, frames, out_stride, 1,
                                                                              row_bytes, vok);
    } else if (channels == 2) {
        pcm_unpack_kernel<B, KIND, BE, 2><<<grid_for(frames_per_thread(B, 2)), threads, 0, ctx->stream>>>(
            d_in, d_out, frames, out_stride, 2, 0, vec_ok);
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
