// encode.cu -- K12 requantisation: encodePCM (A:868-894) behind Audio:pcm (A:901-911) and Audio:wav's
// sample packing (A:981-985).  SURVEY 8(f) rank 2: the step after the preload path in both CLIs
// (auplay.lua:34 plays Audio:stream chunks, auconvert.lua:414 writes Audio:wav).
//
//   value = d * (d < 0 and 2^(b-1) or 2^(b-1) - 1) + (unsigned and 2^(b-1) or 0)      (A:874; float: d, A:873)
//
// The reference returns these UN-ROUNDED Lua numbers; aukit_cuda_dev_encode_pcm writes them as fp64 (the
// product of an f32 sample and a 32-bit scale is exact in fp64, so the values are bit-identical to the
// reference's for every sample an Audio can hold).  aukit_cuda_dev_encode_pcm_bytes narrows the same values
// to packed little-endian integers, which is what the host's string.pack does inside Audio:wav; how that
// rounds is a property of the host Lua, so the mode is a parameter (truncate = Cobalt / a C cast, floor,
// nearest-even); values outside the integer range saturate.
// Layout: interleaved -> out[n*C + c] (A:883), otherwise out[c*len + n] (A:894).
// HBM-bound elementwise work: 4 B read per sample, 8 B (values) or b/8 B (bytes) written.
#include "common.cuh"

namespace {

__device__ __forceinline__ double encode_value(float d, double maxv, double add, bool is_float) {
    if (is_float) return (double)d;                                         // A:873
    const double x = (double)d;
    return __dadd_rn(__dmul_rn(x, d < 0.0f ? maxv : maxv - 1.0), add);      // A:874 (NaN: the comparison is false)
}

__device__ __forceinline__ long long round_mode(double v, int mode) {
    double r = mode == 0 ? trunc(v) : (mode == 1 ? floor(v) : rint(v));
    if (!(r == r)) return 0;                                                // NaN
    if (r > 9.0e18) r = 9.0e18;
    if (r < -9.0e18) r = -9.0e18;
    return (long long)r;
}

// Rounding without the conversion unit, for results that fit 32 bits (8/16/24-bit output): adding 2^52 + 2^51 in
// the wanted rounding direction leaves the integer in the low mantissa word.  ncu showed the packed kernels
// instruction-bound on FRND / F2I (XU pipe) and 64-bit integer clamps; this path is one F2F, one DFMA, one DADD.
//   floor: round-down add;  nearest-even: round-to-nearest add;  truncate: floor of |v| with the sign put back.
// The sample is first clamped to [-4, 4] in fp32 (exact for everything in range; outside it the result saturates
// either way), so the sum stays far below 2^31.  NaN rounds to 0 like round_mode().
__device__ __forceinline__ int quant_magic(float d, double maxv, double add, int mode, int lo, int hi) {
    const float dc = fminf(fmaxf(d, -4.0f), 4.0f);
    const double v = __fma_rn((double)dc, dc < 0.0f ? maxv : maxv - 1.0, add);      // A:874, one rounding like d*s + add
    constexpr double MAGIC = 6755399441055744.0;                                    // 2^52 + 2^51
    int q;
    if (mode == 1) q = __double2loint(__dadd_rd(v, MAGIC));
    else if (mode == 2) q = __double2loint(__dadd_rn(v, MAGIC));
    else {
        const int m = __double2loint(__dadd_rd(fabs(v), MAGIC));
        q = v < 0.0 ? -m : m;
    }
    if (!(d == d)) q = 0;
    return min(max(q, lo), hi);
}

// values: one thread per frame, all channels (coalesced reads per channel row; the C values of a frame are
// adjacent in the interleaved output, so the lanes of a warp write one contiguous span)
__global__ void __launch_bounds__(256)
encode_values_kernel(const float *__restrict__ in, size_t stride, int C, size_t n, double maxv, double add, int is_float,
                     int interleaved, double *__restrict__ out) {
    const size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x, step = (size_t)gridDim.x * blockDim.x;
    size_t i = t0;
    if (interleaved && C == 2) {
        for (; i + 3 * step < n; i += 4 * step) {                        // eight loads in flight per thread
            float a[4], b[4];
#pragma unroll
            for (int u = 0; u < 4; u++) { a[u] = in[i + u * step]; b[u] = in[stride + i + u * step]; }
#pragma unroll
            for (int u = 0; u < 4; u++)
                *reinterpret_cast<double2 *>(out + 2 * (i + u * step)) =
                    make_double2(encode_value(a[u], maxv, add, is_float), encode_value(b[u], maxv, add, is_float));
        }
    }
    for (; i < n; i += step) {
        if (interleaved && C == 2) {
            const double a = encode_value(in[i], maxv, add, is_float), b = encode_value(in[stride + i], maxv, add, is_float);
            *reinterpret_cast<double2 *>(out + 2 * i) = make_double2(a, b);
        } else {
            for (int c = 0; c < C; c++) {
                const double v = encode_value(in[(size_t)c * stride + i], maxv, add, is_float);
                out[interleaved ? i * (size_t)C + c : (size_t)c * n + i] = v;
            }
        }
    }
}

template <int B>
__device__ __forceinline__ void store_le(uint8_t *p, long long q) {
    if (B == 1) p[0] = (uint8_t)q;
    else if (B == 2) *reinterpret_cast<uint16_t *>(p) = (uint16_t)q;
    else if (B == 4) *reinterpret_cast<uint32_t *>(p) = (uint32_t)q;
    else { p[0] = (uint8_t)q; p[1] = (uint8_t)(q >> 8); p[2] = (uint8_t)(q >> 16); }
}

// bytes, fast shapes: 4 consecutive OUTPUT samples per thread, stored as one 4*B-byte word group.
//   ROWS  (mono, or planar with frames % 4 == 0): the 4 samples are 4 consecutive frames of one input row;
//   !ROWS (interleaved stereo): 2 frames x 2 channels.
template <int B, bool ROWS>
__global__ void __launch_bounds__(256)
encode_bytes_vec_kernel(const float *__restrict__ in, size_t stride, size_t n, size_t nrows, double maxv, double add, int is_float,
                        int is_unsigned, int mode, uint8_t *__restrict__ out) {
    const long long lo = is_unsigned ? 0 : -(1ll << (8 * B - 1)), hi = is_unsigned ? (1ll << (8 * B)) - 1 : (1ll << (8 * B - 1)) - 1;
    auto quant = [&](float d) -> uint32_t {
        if (is_float) return __float_as_uint(d);
        if (B <= 3) return (uint32_t)quant_magic(d, maxv, add, mode, (int)lo, (int)hi) & ((1u << (8 * (B & 3))) - 1u);
        long long q = round_mode(encode_value(d, maxv, add, false), mode);
        q = q < lo ? lo : (q > hi ? hi : q);
        return (uint32_t)q;
    };
    const size_t groups_per_row = ROWS ? (n + 3) / 4 : (n + 1) / 2;     // !ROWS: one "row" of frame pairs
    const size_t total = groups_per_row * (ROWS ? nrows : 1);
    // one group: load (returns how many of the 4 samples exist), then quantise + store
    auto load = [&](size_t g, float (&v)[4], size_t &obase) -> int {
        int valid = 4;
        if (ROWS) {
            const size_t r = g / groups_per_row, i = (g % groups_per_row) * 4;
            const float *src = in + r * stride + i;
            if (i + 4 <= n) { const float4 f = *reinterpret_cast<const float4 *>(src); v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w; }
            else { valid = (int)(n - i); for (int k = 0; k < 4; k++) v[k] = k < valid ? src[k] : 0.f; }
            obase = r * n + i;
        } else {
            const size_t i = g * 2;
            if (i + 2 <= n) {
                const float2 a = *reinterpret_cast<const float2 *>(in + i), b = *reinterpret_cast<const float2 *>(in + stride + i);
                v[0] = a.x; v[1] = b.x; v[2] = a.y; v[3] = b.y;
            } else { v[0] = in[i]; v[1] = in[stride + i]; v[2] = v[3] = 0.f; valid = 2; }
            obase = i * 2;
        }
        return valid;
    };
    auto emit = [&](const float (&v)[4], size_t obase, int valid) {
        const uint32_t q0 = quant(v[0]), q1 = quant(v[1]), q2 = quant(v[2]), q3 = quant(v[3]);
        uint8_t *dst = out + obase * B;
        if (valid == 4 && ((uintptr_t)dst % (4 * (B == 3 ? 1 : B))) == 0) {
            if (B == 1) *reinterpret_cast<uint32_t *>(dst) = q0 | (q1 << 8) | (q2 << 16) | (q3 << 24);
            else if (B == 2) *reinterpret_cast<uint2 *>(dst) = make_uint2(q0 | (q1 << 16), q2 | (q3 << 16));
            else if (B == 4) *reinterpret_cast<uint4 *>(dst) = make_uint4(q0, q1, q2, q3);
            else {                                                      // 12 bytes = three words
                uint32_t *w = reinterpret_cast<uint32_t *>(dst);
                w[0] = q0 | (q1 << 24); w[1] = (q1 >> 8) | (q2 << 16); w[2] = (q2 >> 16) | (q3 << 8);
            }
        } else {
            const uint32_t q[4] = {q0, q1, q2, q3};
            for (int k = 0; k < valid; k++) store_le<B>(dst + k * B, (long long)q[k]);
        }
    };
    const size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x, step = (size_t)gridDim.x * blockDim.x;
    size_t g = t0;          // (four groups per iteration was tried: more registers, slower)
    for (; g < total; g += step) {
        float v[4];
        size_t ob;
        const int va = load(g, v, ob);
        emit(v, ob, va);
    }
}

// bytes, 8- and 16-bit, the shapes the players use (mono / planar rows, interleaved stereo): 16 OUTPUT bytes per thread --
// 16 / B consecutive samples loaded as float4s, one STG.128 -- rows on blockIdx.y so that no thread divides.  (The
// 4-byte-per-thread kernel above spent its time on ALU work: ncu r2, s8 mono: ALU pipe 59 %, DRAM 32 %.)
//   PAIRS: interleaved stereo, 16 / (2 B) frames x 2 channels;  !PAIRS: 16 / B frames of row blockIdx.y.
// Requires 16-byte aligned input rows and (n * B) % 16 == 0 when there is more than one row; a row's last, partial
// group is written sample by sample.
template <int B, bool PAIRS>
__global__ void __launch_bounds__(256)
encode_bytes_wide_kernel(const float *__restrict__ in, size_t stride, size_t n, double maxv, double add, int is_unsigned, int mode,
                         uint8_t *__restrict__ out) {
    constexpr int SPT = 16 / B;                                         // samples per thread
    constexpr int FPT = PAIRS ? SPT / 2 : SPT;                          // frames per thread
    const int lo = is_unsigned ? 0 : -(1 << (8 * B - 1)), hi = is_unsigned ? (1 << (8 * B)) - 1 : (1 << (8 * B - 1)) - 1;
    auto quant = [&](float d) -> uint32_t { return (uint32_t)quant_magic(d, maxv, add, mode, lo, hi) & ((1u << (8 * B)) - 1u); };
    const size_t r = PAIRS ? 0 : blockIdx.y;
    const float *row = in + r * stride;
    uint8_t *orow = out + r * n * B;
    const size_t groups = (n + FPT - 1) / FPT;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (size_t)gridDim.x * blockDim.x) {
        const size_t i = g * FPT;
        if (i + FPT <= n) {
            float v[SPT];
            if (PAIRS) {
#pragma unroll
                for (int k = 0; k < FPT / 4; k++) {
                    const float4 a = ldg_stream_f4(reinterpret_cast<const float4 *>(row + i) + k);
                    const float4 b = ldg_stream_f4(reinterpret_cast<const float4 *>(row + stride + i) + k);
                    v[8 * k + 0] = a.x; v[8 * k + 1] = b.x; v[8 * k + 2] = a.y; v[8 * k + 3] = b.y;
                    v[8 * k + 4] = a.z; v[8 * k + 5] = b.z; v[8 * k + 6] = a.w; v[8 * k + 7] = b.w;
                }
            } else {
#pragma unroll
                for (int k = 0; k < SPT / 4; k++) {
                    const float4 a = ldg_stream_f4(reinterpret_cast<const float4 *>(row + i) + k);
                    v[4 * k + 0] = a.x; v[4 * k + 1] = a.y; v[4 * k + 2] = a.z; v[4 * k + 3] = a.w;
                }
            }
            uint32_t w[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (B == 1) w[k] = quant(v[4 * k]) | (quant(v[4 * k + 1]) << 8) | (quant(v[4 * k + 2]) << 16) | (quant(v[4 * k + 3]) << 24);
                else w[k] = quant(v[2 * k]) | (quant(v[2 * k + 1]) << 16);
            }
            asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                         ::"l"(orow + (PAIRS ? 2 * i : i) * B), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
        } else {
            for (size_t f = i; f < n; f++) {
                if (PAIRS) {
                    store_le<B>(orow + (2 * f) * B, (long long)quant(row[f]));
                    store_le<B>(orow + (2 * f + 1) * B, (long long)quant(row[stride + f]));
                } else {
                    store_le<B>(orow + f * B, (long long)quant(row[f]));
                }
            }
        }
    }
}

// bytes, any shape: sample by sample
template <int B>
__global__ void __launch_bounds__(256)
encode_bytes_kernel(const float *__restrict__ in, size_t stride, int C, size_t n, double maxv, double add, int is_float,
                    int is_unsigned, int interleaved, int mode, uint8_t *__restrict__ out) {
    const long long lo = is_unsigned ? 0 : -(1ll << (8 * B - 1)), hi = is_unsigned ? (1ll << (8 * B)) - 1 : (1ll << (8 * B - 1)) - 1;
    auto quant = [&](float d) -> long long {
        if (is_float) return (long long)__float_as_uint(d);                 // 32-bit float samples keep their bits
        long long q = round_mode(encode_value(d, maxv, add, false), mode);
        return q < lo ? lo : (q > hi ? hi : q);
    };
    const size_t total = n * (size_t)C;
    for (size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (size_t)gridDim.x * blockDim.x) {
        size_t c, i;
        if (interleaved) { i = o / (size_t)C; c = o % (size_t)C; }
        else { c = o / n; i = o % n; }
        store_le<B>(out + o * B, quant(in[c * stride + i]));
    }
}

}  // namespace

static int encode_args(int bitDepth, int dataType, double *maxv, double *add) {
    if (bitDepth != 8 && bitDepth != 16 && bitDepth != 24 && bitDepth != 32) return aukit_fail("bad argument #2 (invalid bit depth)");
    if (dataType != AUKIT_SIGNED && dataType != AUKIT_UNSIGNED && dataType != AUKIT_FLOAT) return aukit_fail("bad argument #3 (invalid data type)");
    if (dataType == AUKIT_FLOAT && bitDepth != 32) return aukit_fail("bad argument #2 (float audio must have 32-bit depth)");   // A:909
    *maxv = ldexp(1.0, bitDepth - 1);
    *add = dataType == AUKIT_UNSIGNED ? *maxv : 0.0;
    return 0;
}

extern "C" int aukit_cuda_dev_encode_pcm(aukit_ctx *ctx, const float *d, size_t stride, int channels, size_t n, int bitDepth,
                                         int dataType, int interleaved, double *d_out) {
    if (!ctx) return aukit_fail("aukit_cuda: null context");
    double maxv, add;
    if (encode_args(bitDepth, dataType, &maxv, &add)) return -1;
    if (channels < 1 || n == 0) return 0;
    if (interleaved && channels == 2 && ((uintptr_t)d_out & 15)) return aukit_fail("aukit_cuda: output must be 16-byte aligned");
    const unsigned grid = aukit_grid(n, 256, (size_t)ctx->num_sms * 32);
    encode_values_kernel<<<grid, 256, 0, ctx->stream>>>(d, stride, channels, n, maxv, add, dataType == AUKIT_FLOAT, interleaved, d_out);
    ctx->launches++;
    return aukit_cuda_check(cudaGetLastError(), "encode_values_kernel launch");
}

extern "C" int aukit_cuda_dev_encode_pcm_bytes(aukit_ctx *ctx, const float *d, size_t stride, int channels, size_t n, int bitDepth,
                                               int dataType, int interleaved, int rounding, void *d_out) {
    if (!ctx) return aukit_fail("aukit_cuda: null context");
    double maxv, add;
    if (encode_args(bitDepth, dataType, &maxv, &add)) return -1;
    if (rounding < 0 || rounding > 2) return aukit_fail("aukit_cuda: rounding must be 0 (truncate), 1 (floor) or 2 (nearest)");
    if (channels < 1 || n == 0) return 0;
    uint8_t *o = static_cast<uint8_t *>(d_out);
    const int isf = dataType == AUKIT_FLOAT, isu = dataType == AUKIT_UNSIGNED;
    // fast shapes (see encode_bytes_vec_kernel); rows must keep the 16-byte input alignment
    const bool in_ok = ((uintptr_t)d % 16 == 0) && (channels == 1 || stride % 4 == 0);
    const bool rows = in_ok && (channels == 1 || (!interleaved && n % 4 == 0));
    const bool pairs = in_ok && interleaved && channels == 2;
    // 8 / 16-bit integer output: 16 bytes per thread
    if ((rows || pairs) && !isf && (bitDepth == 8 || bitDepth == 16) && (uintptr_t)o % 16 == 0 &&
        (pairs || channels == 1 || (n * (size_t)(bitDepth / 8)) % 16 == 0)) {
        const size_t fpt = (size_t)(16 / (bitDepth / 8)) / (pairs ? 2 : 1);
        const dim3 wg(aukit_grid((n + fpt - 1) / fpt, 256, (size_t)ctx->num_sms * 32), pairs ? 1u : (unsigned)channels);
        if (bitDepth == 8) {
            if (pairs) encode_bytes_wide_kernel<1, true><<<wg, 256, 0, ctx->stream>>>(d, stride, n, maxv, add, isu, rounding, o);
            else encode_bytes_wide_kernel<1, false><<<wg, 256, 0, ctx->stream>>>(d, stride, n, maxv, add, isu, rounding, o);
        } else {
            if (pairs) encode_bytes_wide_kernel<2, true><<<wg, 256, 0, ctx->stream>>>(d, stride, n, maxv, add, isu, rounding, o);
            else encode_bytes_wide_kernel<2, false><<<wg, 256, 0, ctx->stream>>>(d, stride, n, maxv, add, isu, rounding, o);
        }
        ctx->launches++;
        return aukit_cuda_check(cudaGetLastError(), "encode_bytes_wide_kernel launch");
    }
    if (rows || pairs) {
        const size_t groups = rows ? (n + 3) / 4 * (size_t)channels : (n + 1) / 2;
        const unsigned vg = aukit_grid(groups, 256, (size_t)ctx->num_sms * 32);
#define AUKIT_ENC_VEC(BB)                                                                                                \
    if (rows) encode_bytes_vec_kernel<BB, true><<<vg, 256, 0, ctx->stream>>>(d, stride, n, (size_t)channels, maxv, add, isf, isu, rounding, o); \
    else encode_bytes_vec_kernel<BB, false><<<vg, 256, 0, ctx->stream>>>(d, stride, n, 1, maxv, add, isf, isu, rounding, o)
        switch (bitDepth) {
        case 8: AUKIT_ENC_VEC(1); break;
        case 16: AUKIT_ENC_VEC(2); break;
        case 24: AUKIT_ENC_VEC(3); break;
        default: AUKIT_ENC_VEC(4); break;
        }
#undef AUKIT_ENC_VEC
        ctx->launches++;
        return aukit_cuda_check(cudaGetLastError(), "encode_bytes_vec_kernel launch");
    }
    const unsigned grid = aukit_grid(n * (size_t)channels, 256, (size_t)ctx->num_sms * 32);
    switch (bitDepth) {
    case 8: encode_bytes_kernel<1><<<grid, 256, 0, ctx->stream>>>(d, stride, channels, n, maxv, add, isf, isu, interleaved, rounding, o); break;
    case 16: encode_bytes_kernel<2><<<grid, 256, 0, ctx->stream>>>(d, stride, channels, n, maxv, add, isf, isu, interleaved, rounding, o); break;
    case 24: encode_bytes_kernel<3><<<grid, 256, 0, ctx->stream>>>(d, stride, channels, n, maxv, add, isf, isu, interleaved, rounding, o); break;
    default: encode_bytes_kernel<4><<<grid, 256, 0, ctx->stream>>>(d, stride, channels, n, maxv, add, isf, isu, interleaved, rounding, o); break;
    }
    ctx->launches++;
    return aukit_cuda_check(cudaGetLastError(), "encode_bytes_kernel launch");
}

// Audio:pcm(bitDepth, dataType, interleaved) on a handle, results to HOST memory (n * channels doubles / samples)
extern "C" int aukit_cuda_audio_pcm(aukit_ctx *ctx, const aukit_audio *a, int bitDepth, int dataType, int interleaved, double *h_out) {
    if (!ctx || !a || !h_out) return aukit_fail("aukit_cuda: null argument");
    const size_t total = a->frames * (size_t)a->channels;
    void *d_out = nullptr;
    if (aukit_dev_alloc(ctx, total * sizeof(double) + 16, &d_out)) return -1;
    int rc = aukit_cuda_dev_encode_pcm(ctx, a->data, a->stride, a->channels, a->frames, bitDepth, dataType, interleaved,
                                       static_cast<double *>(d_out));
    if (!rc && total) rc = aukit_cuda_check(cudaMemcpyAsync(h_out, d_out, total * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream), "D2H");
    aukit_dev_free(ctx, d_out);
    if (!rc) rc = aukit_cuda_synchronize(ctx);
    return rc;
}

extern "C" int aukit_cuda_audio_pcm_bytes(aukit_ctx *ctx, const aukit_audio *a, int bitDepth, int dataType, int interleaved,
                                          int rounding, void *h_out) {
    if (!ctx || !a || !h_out) return aukit_fail("aukit_cuda: null argument");
    const size_t total = a->frames * (size_t)a->channels * (size_t)(bitDepth / 8);
    void *d_out = nullptr;
    if (aukit_dev_alloc(ctx, total + 16, &d_out)) return -1;
    int rc = aukit_cuda_dev_encode_pcm_bytes(ctx, a->data, a->stride, a->channels, a->frames, bitDepth, dataType, interleaved, rounding, d_out);
    if (!rc && total) rc = aukit_cuda_check(cudaMemcpyAsync(h_out, d_out, total, cudaMemcpyDeviceToHost, ctx->stream), "D2H");
    aukit_dev_free(ctx, d_out);
    if (!rc) rc = aukit_cuda_synchronize(ctx);
    return rc;
}

// Audio:stream (A:921-937): one chunk of the iterator -- encodePCM with `multiple` set (A:881-891): frames
// [first, first + count) of every channel as un-rounded values, planar h_out[c * count + k].  `count` is clipped to
// the end of the audio; *got receives the frames written (0 past the end: the iterator then returns nil, A:878).
extern "C" int aukit_cuda_audio_stream_chunk(aukit_ctx *ctx, const aukit_audio *a, int bitDepth, int dataType, size_t first,
                                             size_t count, double *h_out, size_t *got) {
    if (!ctx || !a || !got) return aukit_fail("aukit_cuda: null argument");
    double maxv, add;
    if (encode_args(bitDepth, dataType, &maxv, &add)) return -1;
    *got = 0;
    if (first >= a->frames || count == 0) return 0;
    const size_t n = a->frames - first < count ? a->frames - first : count;
    if (!h_out) return aukit_fail("aukit_cuda: null argument");
    const size_t total = n * (size_t)a->channels;
    void *d_out = nullptr;
    if (aukit_dev_alloc(ctx, total * sizeof(double) + 16, &d_out)) return -1;
    int rc = aukit_cuda_dev_encode_pcm(ctx, a->data + first, a->stride, a->channels, n, bitDepth, dataType, 0, static_cast<double *>(d_out));
    // rows of the chunk are n apart on the device; the caller's rows are `count` apart only when n == count
    if (!rc) rc = aukit_cuda_check(cudaMemcpy2DAsync(h_out, count * sizeof(double), d_out, n * sizeof(double), n * sizeof(double),
                                                     (size_t)a->channels, cudaMemcpyDeviceToHost, ctx->stream), "D2H");
    aukit_dev_free(ctx, d_out);
    if (!rc) rc = aukit_cuda_synchronize(ctx);
    if (!rc) *got = n;
    return rc;
}
