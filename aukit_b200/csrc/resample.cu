// resample.cu -- K5 resample (none / linear / cubic) on planar float32.
//
// Replaces Audio:resample (A:653-673) + interpolate.none/linear/cubic (A:253-266).
// Positions are reproduced in fp64 exactly as the reference computes them,
//     x = (i - 1) / ratio + 1,   ratio = newRate / oldRate            (A:658, A:666)
// from the GLOBAL 1-based output index i with an IEEE round-to-nearest division, so
// floor(x), the `x % 1 == 0` exact-hit test (A:667: copy, unclamped) and the fractional
// part are bit-identical to Lua's -- on one GPU or on any time shard (SURVEY finding 5).
// The blend itself runs in fp32: cubic as four bounded Catmull-Rom weights (sum |w| <= 1.25)
// derived from the fp64 fraction, combined with FMAs; the result is clamped to [-1, 1]
// exactly where the reference clamps (A:668).  Missing neighbours at the ends of the signal
// follow the nil substitutions of A:259 and A:264.
//
// One thread per output frame: the position is computed once and shared by all channels.
#include "common.cuh"
#include <stdlib.h>
#include "pipeline.cuh"

#include <math.h>

namespace {

struct resample_args {
    const float *in;      // frames [in_first, in_first + in_avail) of each channel
    size_t in_stride;
    int channels;
    unsigned long long n_total;   // frames of the whole signal (nil beyond)
    unsigned long long in_first;
    double ratio;
    unsigned long long out_first;
    size_t n_out;
    float *out;
    size_t out_stride;
    double y;             // RN(1 / ratio)
    int quotient_fma_ok;  // host-proved: fma(fma(-q0, r, n), y, q0) == RN(n / r) over the whole index range
    double base_d;        // (double)(in_first + 1): 1-based index of the first frame held in `in`
    const float2 *sinc_tab;        // [L][21] {weight, d(weight)/d(fx)} at fx = j / L, or null (sinc only)
    unsigned long long L, M;       // dst / gcd, src / gcd when sinc_tab is set
};

// sinc weights of A:273-276 at the L rational phases, in fp64, with their derivative for the first-order
// correction to the reference's own (rounded) fraction
__global__ void sinc_table_kernel(float2 *tab, unsigned long long L) {
    const unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= L * 21) return;
    const double fx = (double)(e / 21) / (double)L;
    const double px = 3.14159265358979323846 * (fx - (double)((int)(e % 21) - 10));
    if (px == 0.0) { tab[e] = make_float2(1.0f, 0.0f); return; }
    const double sn = sin(px), cs = cos(px);
    tab[e] = make_float2((float)(sn / px), (float)(3.14159265358979323846 * (cs * px - sn) / (px * px)));
}

template <int MODE, bool QFMA>
__global__ void __launch_bounds__(256) resample_kernel(resample_args a) {
    for (size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o < a.n_out;
         o += (size_t)gridDim.x * blockDim.x) {
        const unsigned long long i0 = a.out_first + o;                 // = i - 1
        const double nd = (double)i0;
        // x = (i - 1) / ratio + 1 (A:666): the correctly rounded quotient, either by the 3-operation FMA
        // sequence the host proved exact for this ratio and range, or by the IEEE division
        double q;
        if (QFMA) {
            const double q0 = __dmul_rn(nd, a.y);
            q = __fma_rn(__fma_rn(-q0, a.ratio, nd), a.y, q0);
        } else {
            q = __ddiv_rn(nd, a.ratio);
        }
        const double x = __dadd_rn(q, 1.0);
        const double fl = floor(x);
        const bool hit = (x == fl);                                    // x % 1 == 0, A:667
        const long long f = (long long)fl;                             // 1-based index of p1
        const double fxd = x - fl;                                     // exact
        float w0 = 0.f, w1 = 0.f, w2 = 0.f, w3 = 0.f;
        const float fx = (float)fxd;
        if (MODE == AUKIT_INTERP_CUBIC) {
            // Catmull-Rom weights of A:265, evaluated in fp64 then narrowed
            const double t = fxd, t2 = t * t, t3 = t2 * t;
            w0 = (float)(-0.5 * t3 + t2 - 0.5 * t);
            w1 = (float)(1.5 * t3 - 2.5 * t2 + 1.0);
            w2 = (float)(-1.5 * t3 + 2.0 * t2 + 0.5 * t);
            w3 = (float)(0.5 * t3 - 0.5 * t2);
        }
        const long long n = (long long)a.n_total;
        const long long base = f - 1 - (long long)a.in_first;          // offset of p1 in `in`
        if (MODE == AUKIT_INTERP_SINC) {
            // A:267-281: sum over k = -10..10 of data[f + k] * sin(pi (fx - k)) / (pi (fx - k)), taps outside the
            // signal skipped (not clamped).  The 21 weights are evaluated in fp64 like the reference's and shared
            // by the channels of the frame.
            float w[21];
            bool tabled = false;
            if (a.sinc_tab && !hit) {
                // rational position n*M/L = F + j/L: when the reference's floor agrees (it can be one lower at an
                // exact multiple, SURVEY finding 5) the weights come from row j of the table, corrected to first
                // order for the difference between j/L and the reference's rounded fraction
                const unsigned long long nm = i0 * a.M, F = nm / a.L, j = nm - F * a.L;
                // shift = 1: the reference's floor is one below the rational one (only at exact multiples, j == 0):
                // its fraction is then just under 1, i.e. phase 0 seen from one tap earlier
                const int shift = (int)((long long)F + 1 - f);
                if (shift == 0 || (shift == 1 && j == 0)) {
                    tabled = true;
                    const float dfx = (float)(fxd - ((double)j / (double)a.L + (double)shift));
                    const float2 *row = a.sinc_tab + j * 21;
#pragma unroll
                    for (int k = 0; k < 21; k++) {
                        float2 t = make_float2(0.f, 0.f);
                        if (k - shift >= 0) t = __ldg(row + (k - shift));
                        w[k] = (f + k - 10 >= 1 && f + k - 10 <= n) ? __fmaf_rn(dfx, t.y, t.x) : 0.0f;
                    }
                }
            }
            if (!tabled) {
#pragma unroll
                for (int k = -10; k <= 10; k++) {
                    const double px = 3.14159265358979323846 * (fxd - (double)k);
                    w[k + 10] = (f + k >= 1 && f + k <= n) ? (px == 0.0 ? 1.0f : (float)(sin(px) / px)) : 0.0f;
                }
            }
            for (int c = 0; c < a.channels; c++) {
                const float *ch = a.in + (size_t)c * a.in_stride;
                float v;
                if (hit) {
                    v = ch[base];                                      // copied unclamped, A:667
                } else {
                    float sum = 0.f;
#pragma unroll
                    for (int k = -10; k <= 10; k++)
                        if (f + k >= 1 && f + k <= n) sum = __fmaf_rn(ch[base + k], w[k + 10], sum);
                    v = clamp_ref(sum);
                }
                a.out[(size_t)c * a.out_stride + o] = v;
            }
            continue;
        }
        for (int c = 0; c < a.channels; c++) {
            const float *ch = a.in + (size_t)c * a.in_stride;
            const float p1 = ch[base];
            float v;
            if (hit) {
                v = p1;                                                // copied unclamped
            } else if (MODE == AUKIT_INTERP_NONE) {
                v = clamp_ref(p1);
            } else if (MODE == AUKIT_INTERP_LINEAR) {
                const float p2 = (f + 1 <= n) ? ch[base + 1] : p1;     // data[ffx+1] or data[ffx]
                v = clamp_ref(__fmaf_rn(p2 - p1, fx, p1));
            } else {
                const float p0 = (f - 1 >= 1) ? ch[base - 1] : p1;     // A:264
                const float p2 = (f + 1 <= n) ? ch[base + 1] : p1;
                const float p3 = (f + 2 <= n) ? ch[base + 2] : p2;
                v = clamp_ref(__fmaf_rn(w3, p3, __fmaf_rn(w2, p2, __fmaf_rn(w1, p1, w0 * p0))));
            }
            a.out[(size_t)c * a.out_stride + o] = v;
        }
    }
}

}  // namespace

// Is q1 = fma(fma(-q0, r, n), y, q0) with q0 = RN(n * y), y = RN(1 / r) the correctly rounded n / r for
// EVERY integer n with n / r < 2^max_k?  The value v = q0 + rem * y differs from n / r by at most
// 1.5 * 2^(k-105) in binade k, so RN(v) can differ from RN(n / r) only if n / r lies that close to a
// rounding midpoint mu = U * 2^(k-53) (U odd).  With r = R * 2^g (R odd) and s = 53 - g - k that means
// |n * 2^s - R * U| < 3, i.e. R * U = n * 2^s -+ 1 (the left side is odd, the right side's first term
// even).  For s >= 54 there is exactly one residue U mod 2^s that satisfies it, and it is a midpoint
// only if it falls in [2^53, 2^54).  If no binade has such a U the three-operation quotient is exact
// everywhere in range; otherwise (or when s < 54) the kernel uses the IEEE division instead.
bool aukit_quotient_fma_is_exact(double r, int max_k) {
    typedef unsigned __int128 u128;
    if (!(r > 0) || !isfinite(r)) return false;
    int e = 0;
    const double fr = frexp(r, &e);                       // r = fr * 2^e, fr in [0.5, 1)
    unsigned long long R = (unsigned long long)ldexp(fr, 53);
    int g = e - 53;
    while ((R & 1) == 0) { R >>= 1; g++; }
    // inverse of R modulo 2^128 (Newton), R odd
    u128 inv = R;
    for (int i = 0; i < 8; i++) inv *= (u128)2 - (u128)R * inv;
    int kmin = 0;
    frexp(1.0 / r, &kmin);                                // smallest non-zero quotient is 1 / r
    for (int k = kmin - 2; k <= max_k; k++) {
        const int s = 53 - g - k;
        if (s < 54) return false;                         // several candidates per binade: do not claim exactness
        if (s > 127) return false;                        // residue not decidable in 128 bits
        const u128 mask = (((u128)1) << s) - 1;
        for (int sign = 0; sign < 2; sign++) {
            // R * U == -+1 (mod 2^s)  =>  U == -+inv (mod 2^s)
            u128 U = sign ? (inv & mask) : ((~inv + 1) & mask);
            if (U >= (((u128)1) << 53) && U < (((u128)1) << 54)) return false;
        }
    }
    return true;
}


extern "C" uint64_t aukit_resample_out_len(uint64_t n_in, double srcRate, double dstRate) {
    const double ratio = dstRate / srcRate;              // A:658
    const double newlen = (double)n_in * ratio;          // A:659
    if (!(newlen >= 1.0)) return 0;
    return (uint64_t)floor(newlen);                      // for i = 1, newlen
}

extern "C" double aukit_resample_position(uint64_t i, double srcRate, double dstRate) {
    const double ratio = dstRate / srcRate;
    volatile double q = ((double)i - 1.0) / ratio;       // separately rounded, A:666
    return q + 1.0;
}

extern "C" int aukit_resample_window(uint64_t n_in_total, double srcRate, double dstRate, int interpolation,
                                     uint64_t out_first, uint64_t n_out, uint64_t *in_first,
                                     uint64_t *in_count) {
    if (interpolation < 0 || interpolation > 3) return aukit_fail("bad argument #2 (invalid interpolation type)");
    if (n_out == 0 || n_in_total == 0) { *in_first = 0; *in_count = 0; return 0; }
    const double xa = aukit_resample_position(out_first + 1, srcRate, dstRate);
    const double xb = aukit_resample_position(out_first + n_out, srcRate, dstRate);
    // positions are monotone in i; taps span floor(x) + [lo, hi] (1-based)
    const int lo = interpolation == AUKIT_INTERP_SINC ? -10 : (interpolation == AUKIT_INTERP_CUBIC ? -1 : 0);
    const int hi = interpolation == AUKIT_INTERP_SINC ? 10
                 : (interpolation == AUKIT_INTERP_CUBIC ? 2 : (interpolation == AUKIT_INTERP_LINEAR ? 1 : 0));
    double fa = floor(xa) + lo, fb = floor(xb) + hi;
    if (fa < 1) fa = 1;
    if (fb > (double)n_in_total) fb = (double)n_in_total;
    if (fb < fa) fb = fa;
    *in_first = (uint64_t)fa - 1;
    *in_count = (uint64_t)(fb - fa) + 1;
    return 0;
}

extern "C" int aukit_cuda_dev_resample(aukit_ctx *ctx, const float *d_in, size_t in_stride, int channels,
                                       uint64_t n_in_total, uint64_t in_first, size_t in_avail, double srcRate,
                                       double dstRate, int interpolation, uint64_t out_first, size_t n_out,
                                       float *d_out, size_t out_stride) {
    if (!ctx) return aukit_fail("aukit_cuda: null context");
    if (interpolation < 0 || interpolation > 3) return aukit_fail("bad argument #2 (invalid interpolation type)");  // A:656
    if (channels < 1) return aukit_fail("aukit_cuda: channels < 1");
    if (n_out == 0) return 0;
    const uint64_t total_out = aukit_resample_out_len(n_in_total, srcRate, dstRate);
    if (out_first + n_out > total_out) return aukit_fail("aukit_cuda: output range exceeds floor(n_in * ratio)");
    uint64_t need_first = 0, need_count = 0;
    if (aukit_resample_window(n_in_total, srcRate, dstRate, interpolation, out_first, n_out, &need_first, &need_count))
        return -1;
    // floor(x) can reach n_in_total + 1 only for absurd ratios (nil hole in the reference)
    const double xlast = aukit_resample_position(out_first + n_out, srcRate, dstRate);
    if (floor(xlast) > (double)n_in_total) return aukit_fail("aukit_cuda: position past the end of the input");
    if (need_first < in_first || need_first + need_count > in_first + in_avail)
        return aukit_fail("aukit_cuda: input window [%llu, %llu) does not cover the needed frames [%llu, %llu)",
                          (unsigned long long)in_first, (unsigned long long)(in_first + in_avail),
                          (unsigned long long)need_first, (unsigned long long)(need_first + need_count));
    // integer rates with a short period: the polyphase kernels (pipeline_poly.cu) do the same arithmetic with
    // shared-memory tap reuse and loop-invariant weights; everything else takes the per-frame fp64 kernel below
    if (interpolation != AUKIT_INTERP_SINC) {
        const int rp = aukit_planar_resample_try(ctx, d_in, in_stride, channels, n_in_total, in_first, in_avail, srcRate, dstRate,
                                                 interpolation, out_first, n_out, d_out, out_stride);
        if (rp != 0) return rp < 0 ? -1 : 0;
        const int r = aukit_poly_resample_try(ctx, d_in, in_stride, channels, n_in_total, in_first, in_avail, srcRate, dstRate,
                                              interpolation, out_first, n_out, d_out, out_stride);
        if (r != 0) return r < 0 ? -1 : 0;
    }
    if (interpolation == AUKIT_INTERP_SINC) {
        const int rs = aukit_planar_sinc_try(ctx, d_in, in_stride, channels, n_in_total, in_first, in_avail, srcRate, dstRate, out_first, n_out,
                                             d_out, out_stride);
        if (rs != 0) return rs < 0 ? -1 : 0;
    }
    resample_args a{d_in, in_stride, channels, n_in_total, in_first, dstRate / srcRate, out_first, n_out, d_out, out_stride};
    a.y = 1.0 / a.ratio;
    a.quotient_fma_ok = aukit_quotient_fma_is_exact(a.ratio, 44) ? 1 : 0;
    a.base_d = (double)(in_first + 1);
    a.sinc_tab = nullptr; a.L = a.M = 0;
    void *tab = nullptr;
    if (interpolation == AUKIT_INTERP_SINC && srcRate == floor(srcRate) && dstRate == floor(dstRate) && srcRate >= 1 &&
        dstRate >= 1 && srcRate < 2147483648.0 && dstRate < 2147483648.0 && !getenv("AUKIT_DISABLE_SINC_TABLE")) {
        unsigned long long x = (unsigned long long)srcRate, y = (unsigned long long)dstRate;
        while (y) { const unsigned long long t = x % y; x = y; y = t; }
        const unsigned long long L = (unsigned long long)dstRate / x, M = (unsigned long long)srcRate / x;
        // n * M must stay in 64 bits and the rational fraction within the correction's reach of the reference's
        if (L <= 4096 && (double)(out_first + n_out) * (double)M < 9.0e18 && (double)(out_first + n_out) / a.ratio < 268435456.0) {
            if (aukit_dev_alloc(ctx, (size_t)L * 21 * sizeof(float2), &tab)) return -1;
            sinc_table_kernel<<<(unsigned)((L * 21 + 255) / 256), 256, 0, ctx->stream>>>(static_cast<float2 *>(tab), L);
            ctx->launches++;
            a.sinc_tab = static_cast<const float2 *>(tab);
            a.L = L; a.M = M;
        }
    }
    const int threads = 256;
    const unsigned grid = aukit_grid(n_out, threads, (size_t)ctx->num_sms * 8 * 8);
#define AUKIT_RS(MODE) \
    if (a.quotient_fma_ok) resample_kernel<MODE, true><<<grid, threads, 0, ctx->stream>>>(a); \
    else resample_kernel<MODE, false><<<grid, threads, 0, ctx->stream>>>(a)
    switch (interpolation) {
    case AUKIT_INTERP_NONE: AUKIT_RS(AUKIT_INTERP_NONE); break;
    case AUKIT_INTERP_LINEAR: AUKIT_RS(AUKIT_INTERP_LINEAR); break;
    case AUKIT_INTERP_SINC: AUKIT_RS(AUKIT_INTERP_SINC); break;
    default: AUKIT_RS(AUKIT_INTERP_CUBIC); break;
    }
#undef AUKIT_RS
    ctx->launches++;
    if (tab) aukit_dev_free(ctx, tab);
    return aukit_cuda_check(cudaGetLastError(), "resample_kernel launch");
}
