// sample_formats.cuh -- packed sample bytes -> float, shared by pcm.cu and pipeline.cu.
// Scaling follows aukit.pcm's read() closures (A:1097-1154) and aukit.g711 (A:1374-1379).
#pragma once
#include "common.cuh"

namespace aukit_fmt {

enum Kind : int { K_SIGNED = 0, K_UNSIGNED = 1, K_FLOAT = 2, K_ALAW = 3, K_ULAW = 4 };

__host__ __device__ constexpr int cgcd(int a, int b) { return b == 0 ? a : cgcd(b, a % b); }
// frames per thread: chunk = FPT*C*B bytes is a multiple of 16, FPT a multiple of 4
__host__ __device__ constexpr int frames_per_thread(int B, int C) {
    int f = 16 / cgcd(16, C * B);
    return f < 4 ? 4 : f;
}

// A:1374-1379 for one byte; exact in f32 (integer / 8192)
__device__ __forceinline__ float g711_value(unsigned byte, bool ulaw) {
    unsigned b = byte ^ (ulaw ? 0xFFu : 0x55u);
    unsigned m = b & 0x0Fu, e = (b >> 4) & 7u;
    if (!ulaw && e == 0) m = m * 4 + 2;
    else m = (m * 2 + 33) << e;
    int mi = (int)m;
    if (ulaw) mi -= 33;
    bool neg = ((b & 0x80u) != 0) == ulaw;
    float v = (float)mi * (1.0f / 8192.0f);
    return neg ? -v : v;   // m / -0x2000: yields -0.0 for m == 0, like the reference
}

// 8-bit formats without a table (a 256-entry shared-memory table costs ~2.5 bank-conflict wavefronts per
// random lookup; ncu showed K1/K2 on 8-bit input stalled on mio_throttle).  All forms are exact:
//   PCM (A:1133 / A:1152): with u = s + 128 (signed) or the byte itself (unsigned), float bits 0x4B000000 | u
//     are 2^23 + u, so one FADD gives u - 128; then the same FMA as s16_to_float with 127 * 128 in place of
//     32767 * 32768 (checked for all 256 values against (float)((double)s / 127)).
//   G.711 (A:1374-1379): after the XOR, ((2m + 33) << e) / 8192 has float bits 0x3B840000 + ((b & 0x7F) << 19)
//     -- the exponent and mantissa fields of the code line up with the float's; mu-law subtracts 33/8192
//     (exact), A-law's e == 0 segment is (4m + 2) / 8192; the sign bit is XORed in, which also yields the
//     reference's -0.0 for m == 0.
template <int KIND>
__device__ __forceinline__ float convert8(uint32_t byte) {
    if (KIND == K_ULAW || KIND == K_ALAW) {
        const uint32_t b = byte ^ (KIND == K_ULAW ? 0xFFu : 0x55u);
        float v = __uint_as_float(0x3B840000u + ((b & 0x7Fu) << 19));
        uint32_t sign;
        if (KIND == K_ULAW) {
            v = __fadd_rn(v, -33.0f / 8192.0f);
            sign = (b & 0x80u) << 24;
        } else {
            if ((b & 0x70u) == 0) v = (float)(4 * (b & 0x0Fu) + 2) * (1.0f / 8192.0f);
            sign = (~b & 0x80u) << 24;
        }
        return __uint_as_float(__float_as_uint(v) ^ sign);
    }
    const uint32_t u = KIND == K_SIGNED ? (byte ^ 0x80u) : byte;
    const float lo = __fmul_rn(__fadd_rn(__uint_as_float(0x4B000000u | u), -8388736.0f), 1.0f / 128.0f);
    return __fmaf_rn(__saturatef(lo), (1.0f / (127.0f * 128.0f)) * 128.0f, lo);
}

template <int B, int KIND>
__device__ __forceinline__ float convert(uint32_t raw, const float *lut) {
    if (KIND == K_FLOAT) return __uint_as_float(raw);
    if (KIND == K_ALAW || KIND == K_ULAW) return lut[raw & 0xFFu];
    if (KIND == K_SIGNED) {
        int s = (int)(raw << (32 - 8 * B)) >> (32 - 8 * B);
        if (B == 4) {
            double d = (double)s;
            return (float)(s < 0 ? d / 2147483648.0 : d / 2147483647.0);   // A:1133 in fp64
        }
        constexpr float maxv = (float)(1u << (8 * B - 1));
        float f = (float)s;
        return s < 0 ? f * (1.0f / maxv) : __fdiv_rn(f, maxv - 1.0f);
    }
    // unsigned, A:1152: (s - 128) / (s < 128 and max or max-1) -- literal 128 at every depth
    if (B == 4) {
        double d = (double)raw - 128.0;
        return (float)(raw < 128u ? d / 2147483648.0 : d / 2147483647.0);
    }
    constexpr float maxv = (float)(1u << (8 * B - 1));
    float f = (float)((int)raw - 128);
    return raw < 128u ? f * (1.0f / maxv) : __fdiv_rn(f, maxv - 1.0f);
}

// one sample from global memory at its natural (byte) alignment -- tail / generic path
template <int B, bool BE>
__device__ __forceinline__ uint32_t load_raw(const uint8_t *p) {
    uint32_t v = 0;
    if (BE) {
#pragma unroll
        for (int k = 0; k < B; k++) v = (v << 8) | p[k];
    } else {
#pragma unroll
        for (int k = B - 1; k >= 0; k--) v = (v << 8) | p[k];
    }
    return v;
}

// one sample whose address is a multiple of its size (B = 3: byte loads)
template <int B, bool BE>
__device__ __forceinline__ uint32_t load_raw_aligned(const uint8_t *p) {
    if (B == 1) return *p;
    if (B == 2) {
        const uint32_t v = *reinterpret_cast<const uint16_t *>(p);
        return BE ? __byte_perm(v, 0, 0x4401) : v;
    }
    if (B == 4) {
        const uint32_t v = *reinterpret_cast<const uint32_t *>(p);
        return BE ? __byte_perm(v, 0, 0x0123) : v;
    }
    return load_raw<B, BE>(p);
}

template <int B, bool BE>
__device__ __forceinline__ uint32_t load_raw_aligned_or_bytes(const uint8_t *p, int aligned) {
    return aligned ? load_raw_aligned<B, BE>(p) : load_raw<B, BE>(p);
}

// sample at static byte offset `off` of a register-resident chunk
template <int B, bool BE>
__device__ __forceinline__ uint32_t extract(const uint32_t *w, int off) {
    const int wi = off >> 2, sh = (off & 3) * 8;
    uint32_t v = ((off & 3) + B > 4) ? __funnelshift_r(w[wi], w[wi + 1], sh) : (w[wi] >> sh);
    if (B == 1) return v & 0xFFu;
    if (B == 2) return BE ? __byte_perm(v, 0, 0x4401) : (v & 0xFFFFu);
    if (B == 3) return BE ? __byte_perm(v, 0, 0x4012) : (v & 0xFFFFFFu);
    return BE ? __byte_perm(v, 0, 0x0123) : v;
}

}  // namespace aukit_fmt
