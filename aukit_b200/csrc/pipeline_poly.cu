// pipeline_poly.cu -- rational-ratio (polyphase) fast path of the fused pipeline (K10).
//
// Audio:resample evaluates x = (i-1)/ratio + 1 in fp64 per output sample (A:666).  When both
// rates are integers the positions are rational: with L = new/g, M = old/g (g = gcd), output n
// (0-based) sits at input position n*M/L, i.e. floor F(n) = (n*M) div L and phase
// j(n) = (n*M) mod L, periodic in n with period L.  The kernel exploits that:
//
//   * a CTA walks tiles of K iterations x Sp = L*m consecutive outputs; thread t owns outputs
//     base + k*Sp + t, so its phase j_t = (t*M) mod L -- and therefore its interpolation weights --
//     are loop invariants held in registers (computed once, in fp64, then narrowed);
//   * phase A of a tile converts the K*m*M + 3 input frames the tile touches (its halo
//     included) ONCE from packed bytes to float in shared memory (coalesced global loads);
//   * phase B reads 4 (cubic) / 2 (linear) / 1 (none) taps per channel from shared memory,
//     blends with FMAs, clamps (A:668), mixes to mono (A:686-687) and either max-reduces
//     (peak pass) or scales by peak/max, clamps and stores coalesced float32 (apply pass).
//
// Bit-faithfulness of positions (SURVEY finding 5): for j != 0 the reference's floor(x) equals
// F(n)+1 whenever |x_ref - x_true| < 1/L, and its fraction differs from j/L by at most
// 1.5*x*2^-52, which is below 2^-24 for x < 2^28 -- the regime this path accepts (longer
// buffers take the generic fp64 path unless the ratio is a power of two, where positions are
// exact).  For j == 0 the reference either hits the integer exactly (sample copied UNCLAMPED,
// A:667) or lands one ulp beside it (interpolated and clamped; `none` then selects the
// PREVIOUS sample).  Those outputs -- one per period -- are decided exactly: one warp per tile
// evaluates the reference's own fp64 expression for them (IEEE division) into a small shared
// table before phase B.  That table is only needed when the decision can change the result:
// interpolation `none`, or inputs that can exceed [-1, 1] (float, unsigned > 8 bit).
#include "common.cuh"
#include "pipeline.cuh"
#include "sample_formats.cuh"

#include <math.h>
#include <stdlib.h>

using namespace aukit_fmt;

namespace {

struct poly_plan {
    int L, M;                    // outputs / input frames per period (coprime)
    int m;                       // periods per iteration
    int Sp, Q;                   // outputs / input frames per iteration
    int K;                       // iterations per tile
    int nfr;                     // frames staged per tile = K*Q + 3
    int fmt;                     // (B << 8) | (KIND << 4) | BE
    unsigned long long tile0;    // first global tile index of this launch
    unsigned long long ntiles;
    double y;                    // RN(1 / ratio), for the 3-operation exact quotient (PX_BIG)
    int quotient_fma_ok;         // host-proved: fma(rem, y, q0) == RN(n / ratio) for every n in range
    float eps_r;                 // (new/old) / RN(new/old) - 1: systematic drift of the reference's x (PX_MID)
};

// how output positions are obtained
enum { PX_RATIONAL = 0,   // n*M/L; hit / near-hit indistinguishable in the result (bounded input, interpolating)
       PX_TABLE = 1,      // n*M/L plus an exact fp64 decision table for the one j == 0 output per period
       PX_BIG = 2,        // every output evaluates the reference's fp64 x exactly (positions >= 2^30.5, binade crossings)
       PX_MID = 3 };      // 2^28 <= position < 2^30.5: PX_TABLE plus the systematic drift x*eps_r to first order;
                          // what is left is the quotient's rounding, <= x * 2^-53 <= 2^-22.5

enum { HIT = 0, NEAR_BELOW = 1, NEAR_ABOVE = 2 };

template <int B, int KIND, bool BE>
__device__ __forceinline__ float conv1(const uint8_t *p) {
    if (B == 2 && KIND == K_SIGNED) {
        const uint32_t v = load_raw_aligned<2, BE>(p);
        return s16_to_float((int)(int16_t)v);
    }
    return convert<B, KIND>(load_raw_aligned<B, BE>(p), nullptr);
}

// phase A for one sample format: frames [0, nfr) of the tile -> shared floats.
// CT == 2: float2 per frame; CT == 1: float per frame; CT == 0: planar [C][nfr].
template <int B, int KIND, bool BE, int CT>
__device__ __forceinline__ void stage_tile(const pipe_args &a, const poly_plan &pl, long long g0, float *sm, int plane) {
    const int C = CT ? CT : a.channels;
    const long long n_total = (long long)a.n_total;
    const long long lo = (long long)a.in_first, hi = lo + (long long)a.in_avail;
    for (int i = threadIdx.x; i < pl.nfr; i += blockDim.x) {
        long long g = g0 + i;
        g = g < 0 ? 0 : (g >= n_total ? n_total - 1 : g);      // nil neighbours == clamped index (A:259, A:264)
        const bool ok = g >= lo && g < hi;                      // outside the shard: only masked outputs use it
        const uint8_t *p = a.in + (size_t)(g - lo) * (size_t)(C * B);
        if (CT == 2) {
            float2 v = make_float2(0.f, 0.f);
            if (ok) {
                if (B == 2 && KIND == K_SIGNED && !BE) {        // one 32-bit load per stereo frame
                    const uint32_t w = *reinterpret_cast<const uint32_t *>(p);
                    v.x = s16_to_float((int)(int16_t)(w & 0xFFFFu));
                    v.y = s16_to_float((int)(int16_t)(w >> 16));
                } else {
                    v.x = conv1<B, KIND, BE>(p);
                    v.y = conv1<B, KIND, BE>(p + B);
                }
            }
            reinterpret_cast<float2 *>(sm)[i] = v;
        } else if (CT == 1) {
            sm[i] = ok ? conv1<B, KIND, BE>(p) : 0.f;
        } else {
            for (int c = 0; c < C; c++) sm[(size_t)c * plane + i] = ok ? conv1<B, KIND, BE>(p + c * B) : 0.f;
        }
    }
}

// phase A for planar float32 input (standalone Audio:resample): value pass-through, same clamped indexing
template <int CT>
__device__ __forceinline__ void stage_planar(const pipe_args &a, const poly_plan &pl, long long g0, float *sm, int plane) {
    const int C = CT ? CT : a.channels;
    const long long n_total = (long long)a.n_total;
    const long long lo = (long long)a.in_first, hi = lo + (long long)a.in_avail;
    const float *in = reinterpret_cast<const float *>(a.in);
    for (int i = threadIdx.x; i < pl.nfr; i += blockDim.x) {
        long long g = g0 + i;
        g = g < 0 ? 0 : (g >= n_total ? n_total - 1 : g);
        const bool ok = g >= lo && g < hi;
        const size_t off = (size_t)(g - lo);
        if (CT == 2) {
            reinterpret_cast<float2 *>(sm)[i] = ok ? make_float2(in[off], in[a.in_stride + off]) : make_float2(0.f, 0.f);
        } else if (CT == 1) {
            sm[i] = ok ? in[off] : 0.f;
        } else {
            for (int c = 0; c < C; c++) sm[(size_t)c * plane + i] = ok ? in[(size_t)c * a.in_stride + off] : 0.f;
        }
    }
}

template <int CT>
__device__ __forceinline__ void stage_dispatch(const pipe_args &a, const poly_plan &pl, long long g0, float *sm, int plane) {
    if (a.planar_f32) { stage_planar<CT>(a, pl, g0, sm, plane); return; }
    switch (pl.fmt) {
#define AUKIT_STAGE(BB, KK, EE) case ((BB << 8) | (KK << 4) | EE): stage_tile<BB, KK, (EE != 0), CT>(a, pl, g0, sm, plane); break;
        AUKIT_STAGE(1, K_SIGNED, 0) AUKIT_STAGE(1, K_UNSIGNED, 0)
        AUKIT_STAGE(2, K_SIGNED, 0) AUKIT_STAGE(2, K_SIGNED, 1) AUKIT_STAGE(2, K_UNSIGNED, 0) AUKIT_STAGE(2, K_UNSIGNED, 1)
        AUKIT_STAGE(3, K_SIGNED, 0) AUKIT_STAGE(3, K_SIGNED, 1) AUKIT_STAGE(3, K_UNSIGNED, 0) AUKIT_STAGE(3, K_UNSIGNED, 1)
        AUKIT_STAGE(4, K_SIGNED, 0) AUKIT_STAGE(4, K_SIGNED, 1) AUKIT_STAGE(4, K_UNSIGNED, 0) AUKIT_STAGE(4, K_UNSIGNED, 1)
        AUKIT_STAGE(4, K_FLOAT, 0) AUKIT_STAGE(4, K_FLOAT, 1)
#undef AUKIT_STAGE
    default: break;
    }
}

// two s16 little-endian samples of one 32-bit word -> floats (A:1133 scaling, exact)
__device__ __forceinline__ float2 s16x2_to_float2(uint32_t w) {
    return make_float2(s16_to_float((int)(int16_t)(w & 0xFFFFu)), s16_to_float((int)w >> 16));
}

// CT: compile-time channel count (1, 2) or 0 = runtime.  PX: position mode (above); PX != 0 also
// keeps the reference's NaN-transparent clamp (inputs may be non-finite / out of range).
template <int CT, int MODE, bool MONO, bool APPLY, int PX>
__global__ void __launch_bounds__(512, 3) poly_kernel(pipe_args a, poly_plan pl) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *sm = reinterpret_cast<float *>(smem_raw);
    const int C = CT ? CT : a.channels;
    const int nfr_cap = pl.nfr + 16;                                    // room for the 16-byte alignment shift
    unsigned char *hit_tab = smem_raw + (size_t)nfr_cap * C * sizeof(float);   // K*m entries
    __shared__ float wm[32];

    const int t = threadIdx.x;
    const bool active = t < pl.Sp;
    // loop-invariant phase of this thread: position of output (base + t) relative to the tile base
    const long long tm = (long long)t * pl.M;
    const int off_t = (int)(tm / pl.L), j_t = (int)(tm % pl.L);
    const bool is_j0 = (j_t == 0);
    float w0 = 0.f, w1 = 1.f, w2 = 0.f, w3 = 0.f, fx = 0.f;
    float wd0 = 0.f, wd1 = 0.f, wd2 = 0.f, wd3 = 0.f;                   // d(weight)/d(fx), PX_BIG only
    const double jL = (double)j_t / (double)pl.L;                       // rational fraction of this thread
    {
        const double x = jL;
        fx = (float)x;
        if (MODE == AUKIT_INTERP_CUBIC) {                               // Catmull-Rom weights of A:265
            const double x2 = x * x, x3 = x2 * x;
            w0 = (float)(-0.5 * x3 + x2 - 0.5 * x);
            w1 = (float)(1.5 * x3 - 2.5 * x2 + 1.0);
            w2 = (float)(-1.5 * x3 + 2.0 * x2 + 0.5 * x);
            w3 = (float)(0.5 * x3 - 0.5 * x2);
            if (PX == PX_BIG || PX == PX_MID) {
                wd0 = (float)(-1.5 * x2 + 2.0 * x - 0.5);
                wd1 = (float)(4.5 * x2 - 5.0 * x);
                wd2 = (float)(-4.5 * x2 + 4.0 * x + 0.5);
                wd3 = (float)(1.5 * x2 - x);
            }
        }
    }
    float mult = 0.f;
    if (APPLY && !a.raw_out) mult = (float)(a.peak / (double)a.d_max[0]);   // A:3444
    // final clamp bounds (A:3455).  max == 0 (silence) makes every product 0 * inf = NaN, which the reference's
    // clamp lets through (A:228): NaN bounds keep it NaN through fmin/fmax as well.
    const float fin_hi = (APPLY && !a.raw_out && !(a.d_max[0] > 0.f)) ? __int_as_float(0x7FC00000) : 1.0f;
    const float fin_lo = -fin_hi;
    const float inv_cn = a.inv_cn;                                      // s / cn (A:687) as a multiply
    float mx = 0.f;
    const unsigned long long out_lo = a.out_first, out_hi = a.out_first + a.n_out;
    const unsigned long long tile_out = (unsigned long long)pl.Sp * pl.K;
    const long long n_total = (long long)a.n_total;
    const long long in_lo = (long long)a.in_first, in_hi = in_lo + (long long)a.in_avail;
    const bool s16le = !a.planar_f32 && pl.fmt == ((2 << 8) | (K_SIGNED << 4) | 0);

    auto clamp_out = [&](float v) {
        if (PX != PX_RATIONAL) return clamp_ref(v);
        return fminf(fmaxf(v, fin_lo), fin_hi);
    };
    auto clampv = [](float v) {
        if (PX != PX_RATIONAL) return clamp_ref(v);                     // NaN passes through like A:228-232
        return fminf(fmaxf(v, -1.0f), 1.0f);                            // finite inputs: same result, 2 FMNMX
    };

    for (unsigned long long tile = pl.tile0 + blockIdx.x; tile < pl.tile0 + pl.ntiles; tile += gridDim.x) {
        const unsigned long long base_out = tile * tile_out;
        const long long F0 = (long long)(tile * (unsigned long long)pl.K * (unsigned long long)pl.Q);  // floor frame of base_out
        const long long gA = F0 - 1;                                    // first frame the tile needs (p0 of output 0)
        __syncthreads();                                                // previous tile fully consumed
        // ---------------- phase A: packed frames -> shared floats (each frame converted once)
        int sh = 0;                                                     // shared index of frame gA
        bool fast = false;
        if ((CT == 1 || CT == 2) && s16le && gA >= in_lo && gA >= 0 && ((uintptr_t)a.in & 15) == 0) {
            // interior tile of 16-bit LE input: 128-bit streaming loads from a 16-byte aligned start
            const size_t boff = (size_t)(gA - in_lo) * (size_t)(2 * CT);
            const size_t a0 = boff & ~(size_t)15;
            const int shf = (int)((boff - a0) / (size_t)(2 * CT));
            constexpr int FPV = 16 / (2 * (CT ? CT : 1));               // frames per 16-byte vector
            const int nvec = (pl.nfr + shf + FPV - 1) / FPV;
            const long long last = (long long)(a0 / (size_t)(2 * CT)) + in_lo + (long long)nvec * FPV;   // one past the last frame read
            if (last <= in_hi && last <= n_total) {
                fast = true;
                sh = shf;
                const uint4 *src = reinterpret_cast<const uint4 *>(a.in + a0);
                float4 *dst = reinterpret_cast<float4 *>(sm);
                for (int v = t; v < nvec; v += blockDim.x) {
                    const uint4 w = ldg_stream(src + v);
                    const float2 f0 = s16x2_to_float2(w.x), f1 = s16x2_to_float2(w.y);
                    const float2 f2 = s16x2_to_float2(w.z), f3 = s16x2_to_float2(w.w);
                    dst[2 * v] = make_float4(f0.x, f0.y, f1.x, f1.y);
                    dst[2 * v + 1] = make_float4(f2.x, f2.y, f3.x, f3.y);
                }
            }
        }
        if ((CT == 1 || CT == 2) && a.planar_f32 && gA >= in_lo && gA >= 0 && ((uintptr_t)a.in & 15) == 0 &&
            (CT == 1 || (a.in_stride & 3) == 0)) {
            // interior tile of planar f32 input (standalone Audio:resample): 128-bit loads per row from a 16-byte
            // aligned start, interleaved into shared memory; no index clamping needed inside the signal
            const size_t foff = (size_t)(gA - in_lo);
            const size_t a0 = foff & ~(size_t)3;
            const int shf = (int)(foff - a0);
            const int nvec = (pl.nfr + shf + 3) / 4;
            const long long last = (long long)a0 + in_lo + (long long)nvec * 4;                 // one past the last frame read
            if (last <= in_hi && last <= n_total) {
                fast = true;
                sh = shf;
                const float4 *r0 = reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(a.in) + a0);
                if (CT == 1) {
                    float4 *dst = reinterpret_cast<float4 *>(sm);
                    for (int v = t; v < nvec; v += blockDim.x) dst[v] = __ldg(r0 + v);
                } else {
                    const float4 *r1 = reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(a.in) + a.in_stride + a0);
                    float4 *dst = reinterpret_cast<float4 *>(sm);
                    for (int v = t; v < nvec; v += blockDim.x) {
                        const float4 l = __ldg(r0 + v), r = __ldg(r1 + v);
                        dst[2 * v] = make_float4(l.x, r.x, l.y, r.y);
                        dst[2 * v + 1] = make_float4(l.z, r.z, l.w, r.w);
                    }
                }
            }
        }
        if (!fast) stage_dispatch<CT>(a, pl, gA, sm, nfr_cap);
        if (PX == PX_TABLE || PX == PX_MID) {
            // exact hit / near-hit decision for the tile's j == 0 outputs with the reference's own
            // fp64 expression (A:666-667); entry e <-> iteration e / m, period e % m
            for (int e = t; e < pl.K * pl.m; e += blockDim.x) {
                const unsigned long long n = base_out + (unsigned long long)(e / pl.m) * pl.Sp + (unsigned long long)(e % pl.m) * pl.L;
                const double xt = (double)(F0 + (long long)(e / pl.m) * pl.Q + (long long)(e % pl.m) * pl.M + 1);
                const double x = __dadd_rn(__ddiv_rn((double)n, a.ratio), 1.0);
                hit_tab[e] = (x == xt) ? HIT : (x < xt ? NEAR_BELOW : NEAR_ABOVE);
            }
        }
        __syncthreads();
        // ---------------- phase B: taps -> blend -> clamp -> mono -> peak / scale+store
        if (!active) continue;
        int k0 = 0, k1 = pl.K;
        if (base_out < out_lo || base_out + tile_out > out_hi) {        // first / last tile of a shard: mask
            const long long lo = (long long)out_lo - (long long)base_out - t;
            const long long hi = (long long)out_hi - (long long)base_out - t;
            k0 = lo <= 0 ? 0 : (int)((lo + pl.Sp - 1) / pl.Sp);
            k1 = hi <= 0 ? 0 : (int)((hi + pl.Sp - 1) / pl.Sp);
            if (k1 > pl.K) k1 = pl.K;
        }
        int s = k0 * pl.Q + off_t + sh;                                 // shared index of p0
        // PX_BIG: exact reference position of every output.  nd = global output index, xr = the
        // integer part of the rational position (1-based) it should be near; both advance by exact steps.
        double nd = 0.0, xr = 0.0;
        if (PX == PX_BIG) {
            nd = (double)(base_out + (unsigned long long)k0 * pl.Sp + t);
            xr = (double)(F0 + (long long)k0 * pl.Q + off_t + 1);     // integer part only: exact in fp64
        }
        float xf = 0.f;                                                 // PX_MID: rational position - 1, in fp32
        if (PX == PX_MID) xf = (float)(F0 + (long long)k0 * pl.Q + off_t) + fx;
        const float qf = (float)pl.Q;
        float *outp = nullptr;
        if (APPLY) outp = a.out + (size_t)((long long)base_out - (long long)out_lo + (long long)k0 * pl.Sp + t);
#pragma unroll 4
        for (int k = k0; k < k1; k++, s += pl.Q) {
            int st = NEAR_ABOVE;
            if ((PX == PX_TABLE || PX == PX_MID) && is_j0) st = hit_tab[k * pl.m + t / pl.L];
            float cw0 = w0, cw1 = w1, cw2 = w2, cw3 = w3, cfx = fx;
            if (PX == PX_BIG) {
                // x = (i - 1) / ratio + 1 exactly as A:666: correctly rounded quotient, then + 1
                double q;
                if (pl.quotient_fma_ok) {
                    const double q0 = __dmul_rn(nd, pl.y);
                    const double rem = __fma_rn(-q0, a.ratio, nd);
                    q = __fma_rn(rem, pl.y, q0);
                } else {
                    q = __ddiv_rn(nd, a.ratio);
                }
                const double dev = (__dadd_rn(q, 1.0) - xr) - jL;       // reference - rational position; first difference is exact
                const float dl = (float)dev;                            // |dev| <= ~x * 2^-51
                if (is_j0) st = dev == 0.0 ? HIT : (dev < 0.0 ? NEAR_BELOW : NEAR_ABOVE);
                if (MODE == AUKIT_INTERP_CUBIC) {                       // first-order update of the weights
                    cw0 = __fmaf_rn(wd0, dl, w0); cw1 = __fmaf_rn(wd1, dl, w1);
                    cw2 = __fmaf_rn(wd2, dl, w2); cw3 = __fmaf_rn(wd3, dl, w3);
                } else {
                    cfx = fx + dl;
                }
                nd += (double)pl.Sp;
                xr += (double)pl.Q;
            }
            if (PX == PX_MID) {
                const float dl = xf * pl.eps_r;                         // x_ref - x_rational, systematic part
                if (MODE == AUKIT_INTERP_CUBIC) {
                    cw0 = __fmaf_rn(wd0, dl, w0); cw1 = __fmaf_rn(wd1, dl, w1);
                    cw2 = __fmaf_rn(wd2, dl, w2); cw3 = __fmaf_rn(wd3, dl, w3);
                } else {
                    cfx = fx + dl;
                }
                xf += qf;
            }
            // one channel: blend -> clamp (A:668) / exact-hit copy (A:667)
            auto value = [&](float p0, float p1, float p2, float p3) {
                float v;
                if (MODE == AUKIT_INTERP_CUBIC) v = __fmaf_rn(cw3, p3, __fmaf_rn(cw2, p2, __fmaf_rn(cw1, p1, cw0 * p0)));
                else if (MODE == AUKIT_INTERP_LINEAR) v = __fmaf_rn(p2 - p1, cfx, p1);
                else v = p1;
                // PX_BIG, j == 0, x just BELOW the integer: the reference interpolates on the previous
                // segment [p0, p1] at fx = 1 + dev.  Linear has a slope break at the knot (cubic is C1).
                if ((PX == PX_BIG || PX == PX_MID) && MODE == AUKIT_INTERP_LINEAR && is_j0 && st == NEAR_BELOW) v = __fmaf_rn(p1 - p0, cfx, p1);
                if (PX != PX_RATIONAL && is_j0) {
                    if (st == HIT) return p1;                                              // copied unclamped, A:667
                    if (MODE == AUKIT_INTERP_NONE && st == NEAR_BELOW) return clampv(p0);  // floor(x) is one lower
                    if (PX == PX_TABLE) return clampv(p1);          // weights at j == 0 are (0, 1, 0, 0)
                }
                return clampv(v);
            };
            float acc = 0.f;
            if (CT == 2) {
                const float2 *f = reinterpret_cast<const float2 *>(sm) + s;
                const float2 z = make_float2(0.f, 0.f);
                const float2 f1 = f[1];
                const float2 f0 = (MODE == AUKIT_INTERP_CUBIC || PX != PX_RATIONAL) ? f[0] : z;
                const float2 f2 = (MODE != AUKIT_INTERP_NONE) ? f[2] : z;
                const float2 f3 = (MODE == AUKIT_INTERP_CUBIC) ? f[3] : z;
                const float vl = value(f0.x, f1.x, f2.x, f3.x), vr = value(f0.y, f1.y, f2.y, f3.y);
                if (MONO) acc = vl + vr;                                // (0 + L) + R, A:686
                else if (APPLY) {
                    if (a.raw_out) { outp[0] = vl; outp[a.out_stride] = vr; }
                    else { outp[0] = clamp_out(vl * mult); outp[a.out_stride] = clamp_out(vr * mult); }
                }
                else mx = fmaxf(mx, fmaxf(fabsf(vl), fabsf(vr)));
            } else {
                for (int c = 0; c < C; c++) {
                    const float *f = sm + (CT == 1 ? 0 : (size_t)c * nfr_cap) + s;
                    const float p1 = f[1];
                    const float p0 = (MODE == AUKIT_INTERP_CUBIC || PX != PX_RATIONAL) ? f[0] : 0.f;
                    const float p2 = (MODE != AUKIT_INTERP_NONE) ? f[2] : 0.f;
                    const float p3 = (MODE == AUKIT_INTERP_CUBIC) ? f[3] : 0.f;
                    const float v = value(p0, p1, p2, p3);
                    if (MONO) acc += v;
                    else if (APPLY) outp[(size_t)c * a.out_stride] = a.raw_out ? v : clamp_out(v * mult);
                    else mx = fmaxf(mx, fabsf(v));
                }
            }
            if (MONO) {
                const float mv = acc * inv_cn;                          // s / cn, A:687
                if (APPLY) outp[0] = clamp_out(mv * mult);              // A:3455
                else mx = fmaxf(mx, fabsf(mv));
            }
            if (APPLY) outp += pl.Sp;
        }
    }
    if (!APPLY) {
        mx = warp_max(mx);
        if ((t & 31) == 0) wm[t >> 5] = mx;
        __syncthreads();
        if (t < 32) {
            mx = t < ((blockDim.x + 31) >> 5) ? wm[t] : 0.0f;
            mx = warp_max(mx);
            if (t == 0) atomic_max_nonneg(a.d_max, mx);
        }
    }
}

long long gcd_ll(long long x, long long y) { while (y) { long long r = x % y; x = y; y = r; } return x; }

template <int CT, int MODE, bool MONO, bool APPLY>
int launch_exact(aukit_ctx *ctx, const pipe_args &a, const poly_plan &pl, int px, int threads, size_t smem) {
    auto go = [&](auto kern) -> int {
        if (aukit_cuda_check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "smem attr")) return -1;
        int occ = 0;
        if (aukit_cuda_check(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem), "occupancy")) return -1;
        if (occ < 1) return aukit_fail("aukit_cuda: polyphase kernel does not fit on an SM");
        unsigned long long g = (unsigned long long)ctx->num_sms * occ;
        if (g > pl.ntiles) g = pl.ntiles;
        kern<<<(unsigned)g, threads, smem, ctx->stream>>>(a, pl);
        ctx->launches++;
        return aukit_cuda_check(cudaGetLastError(), "poly_kernel launch");
    };
    if (px == PX_BIG) return go(poly_kernel<CT, MODE, MONO, APPLY, PX_BIG>);
    if (px == PX_MID) return go(poly_kernel<CT, MODE, MONO, APPLY, PX_MID>);
    return px == PX_TABLE ? go(poly_kernel<CT, MODE, MONO, APPLY, PX_TABLE>) : go(poly_kernel<CT, MODE, MONO, APPLY, PX_RATIONAL>);
}

template <int CT, int MODE>
int launch_flags(aukit_ctx *ctx, const pipe_args &a, const poly_plan &pl, int exact, bool apply, int threads, size_t smem) {
    if (a.mono) return apply ? launch_exact<CT, MODE, true, true>(ctx, a, pl, exact, threads, smem)
                             : launch_exact<CT, MODE, true, false>(ctx, a, pl, exact, threads, smem);
    return apply ? launch_exact<CT, MODE, false, true>(ctx, a, pl, exact, threads, smem)
                 : launch_exact<CT, MODE, false, false>(ctx, a, pl, exact, threads, smem);
}

template <int CT>
int launch_interp(aukit_ctx *ctx, const pipe_args &a, const poly_plan &pl, int interp, int exact, bool apply, int threads, size_t smem) {
    switch (interp) {
    case AUKIT_INTERP_NONE: return launch_flags<CT, AUKIT_INTERP_NONE>(ctx, a, pl, exact, apply, threads, smem);
    case AUKIT_INTERP_LINEAR: return launch_flags<CT, AUKIT_INTERP_LINEAR>(ctx, a, pl, exact, apply, threads, smem);
    default: return launch_flags<CT, AUKIT_INTERP_CUBIC>(ctx, a, pl, exact, apply, threads, smem);
    }
}

}  // namespace

int aukit_pipeline_run_try(aukit_ctx *ctx, const pipe_args &a, const aukit_pipeline_desc *p, bool apply, long long L,
                           long long M, double eps_r, bool pow2_ratio, unsigned long long *done_first,
                           unsigned long long *done_count);

static int poly_range(aukit_ctx *ctx, const pipe_args &a, const aukit_pipeline_desc *p, bool apply, bool allow_run);

int aukit_pipeline_poly_try(aukit_ctx *ctx, const pipe_args &a, const aukit_pipeline_desc *p, bool apply) {
    return poly_range(ctx, a, p, apply, true);
}

int aukit_poly_resample_try(aukit_ctx *ctx, const float *d_in, size_t in_stride, int channels, unsigned long long n_in_total,
                            unsigned long long in_first, size_t in_avail, double srcRate, double dstRate, int interpolation,
                            unsigned long long out_first, size_t n_out, float *d_out, size_t out_stride) {
    static const bool disabled = getenv("AUKIT_DISABLE_POLY") && getenv("AUKIT_DISABLE_POLY")[0] == '1';
    if (disabled) return 0;
    aukit_pipeline_desc d{};
    d.bitDepth = 32; d.dataType = AUKIT_FLOAT; d.channels = channels; d.bigEndian = 0;
    d.srcRate = srcRate; d.dstRate = dstRate; d.interpolation = interpolation; d.mono = 0;
    d.n_in_total = n_in_total; d.in_first = in_first; d.in_avail = in_avail; d.out_first = out_first; d.n_out = n_out;
    pipe_args a{};
    a.in = reinterpret_cast<const uint8_t *>(d_in);
    a.channels = channels;
    a.n_total = n_in_total; a.in_first = in_first; a.in_avail = in_avail;
    a.ratio = dstRate / srcRate;
    a.out_first = out_first; a.n_out = n_out;
    a.mono = 0; a.inv_cn = 1.0f; a.cn_pow2 = 1;
    a.d_max = nullptr; a.peak = 1.0;
    a.out = d_out; a.out_stride = out_stride;
    a.planar_f32 = 1; a.in_stride = in_stride; a.raw_out = 1;
    return poly_range(ctx, a, &d, true, false);
}

static int poly_range(aukit_ctx *ctx, const pipe_args &a, const aukit_pipeline_desc *p, bool apply, bool allow_run) {
    // AUKIT_DISABLE_POLY=1 forces the generic fp64-position kernels (used by the tests to cross-check)
    static const bool disabled = getenv("AUKIT_DISABLE_POLY") && getenv("AUKIT_DISABLE_POLY")[0] == '1';
    if (disabled) return 0;
    // integer rates only
    const double sr = p->srcRate, dr = p->dstRate;
    if (!(sr >= 1 && dr >= 1 && sr < 2147483648.0 && dr < 2147483648.0) || sr != floor(sr) || dr != floor(dr)) return 0;
    const long long g = gcd_ll((long long)sr, (long long)dr);
    const long long L = (long long)dr / g, M = (long long)sr / g;
    if (L > 512 || M > (1 << 20)) return 0;
    // position regime: rational positions are within 2^-24 of the reference's below 2^28 frames;
    // power-of-two ratios are exact at any length
    int e2 = 0;
    const bool pow2_ratio = frexp(a.ratio, &e2) == 0.5;
    const double last_pos = (double)(a.out_first + a.n_out) * (double)M / (double)L;
    const bool big = !pow2_ratio && last_pos >= 268435456.0;
    if (big && last_pos >= 1099511627776.0) return 0;                   // 2^40: beyond the proved range
    const int B = p->bitDepth / 8;
    if ((B == 2 || B == 4) && ((uintptr_t)a.in % B)) return 0;
    if (L == 1 && pow2_ratio) {
        // ratio 2^-k: every position is an exact integer, every output a copied sample (A:667): strided gather (K15)
        const int r = aukit_pipeline_decim_try(ctx, a, p, apply, M);
        if (r != 0) return r;
    }
    if (allow_run) {
        // headline shape: the run-per-lane kernel takes the interior, the kernels below the edges
        const double eps = fma(-(double)M, a.ratio, (double)L) / ((double)M * a.ratio);
        unsigned long long df = 0, dc = 0;
        // The head and tail kernels are a few thousand outputs each but ~9 us of launch + table set-up apiece: 4 of them
        // were 9 % of the step behind the interior kernel (profiles/r1_static_launches.txt).  They go to a side stream
        // forked here and joined below, and the interior kernel leaves two SMs free for them (launch_run_static), so
        // they run beside it from the start.  The peak pass's atomicMax commutes and the apply pass's output ranges
        // are disjoint, so nothing else orders them.  (Measured on one box: same stream 0.384 ms / step, side stream
        // 0.370, + two reserved SMs 0.351; head and tail on two side streams was no better, 0.357.)
        constexpr bool no_side = false;
        cudaStream_t main_stream = ctx->stream;
        if (!no_side && aukit_cuda_check(cudaEventRecord(ctx->ev_fork, main_stream), "fork event")) return -1;
        const int r = aukit_pipeline_run_try(ctx, a, p, apply, L, M, eps, pow2_ratio, &df, &dc);
        if (r < 0) return -1;
        if (r == 1) {
            int rc = 1;
            if (!no_side) {
                if (aukit_cuda_check(cudaStreamWaitEvent(ctx->side_stream, ctx->ev_fork, 0), "fork wait")) return -1;
                ctx->stream = ctx->side_stream;
            }
            if (df > a.out_first) {
                pipe_args h = a;
                h.n_out = (size_t)(df - a.out_first);
                const int e = poly_range(ctx, h, p, apply, false);
                if (e != 1) { if (e == 0) aukit_fail("aukit_cuda: no polyphase kernel for the head of the range"); rc = -1; }
            }
            const unsigned long long end = a.out_first + a.n_out;
            if (rc == 1 && df + dc < end) {
                pipe_args t = a;
                t.out_first = df + dc;
                t.n_out = (size_t)(end - (df + dc));
                if (apply) t.out = a.out + (size_t)(df + dc - a.out_first);
                const int e = poly_range(ctx, t, p, apply, false);
                if (e != 1) { if (e == 0) aukit_fail("aukit_cuda: no polyphase kernel for the tail of the range"); rc = -1; }
            }
            if (!no_side) {
                ctx->stream = main_stream;
                if (aukit_cuda_check(cudaEventRecord(ctx->ev_join, ctx->side_stream), "join event")) return -1;
                if (aukit_cuda_check(cudaStreamWaitEvent(main_stream, ctx->ev_join, 0), "join wait")) return -1;
            }
            return rc;
        }
    }
    const int kind = p->dataType == AUKIT_FLOAT ? K_FLOAT : (p->dataType == AUKIT_UNSIGNED ? K_UNSIGNED : K_SIGNED);
    const int be = (p->bigEndian && B > 1) ? 1 : 0;

    poly_plan pl{};
    pl.L = (int)L; pl.M = (int)M;
    pl.m = (int)(L >= 512 ? 1 : 512 / L);
    if ((long long)pl.m * M > 4096) pl.m = (int)(4096 / M) > 0 ? (int)(4096 / M) : 1;
    pl.Sp = pl.L * pl.m;
    pl.Q = pl.M * pl.m;
    const int threads = (pl.Sp + 31) / 32 * 32;
    if (threads > 512) return 0;
    const int C = p->channels;
    const size_t frame_bytes = sizeof(float) * (size_t)C;
    const size_t budget = 40 * 1024;
    long long K = ((long long)(budget / frame_bytes) - 3) / pl.Q;
    if (K < 1) K = 1;
    if (K > 64) K = 64;
    pl.K = (int)K;
    pl.nfr = pl.K * pl.Q + 3;
    pl.fmt = (B << 8) | (kind << 4) | be;
    size_t smem = (size_t)(pl.nfr + 16) * frame_bytes + (size_t)pl.K * pl.m + 16;
    smem = (smem + 15) / 16 * 16;
    if (smem > 200 * 1024) return 0;
    const unsigned long long tile_out = (unsigned long long)pl.Sp * pl.K;
    const unsigned long long tile_first = a.out_first / tile_out;
    const unsigned long long tile_last = (a.out_first + a.n_out - 1) / tile_out;
    // exact j == 0 decisions matter when the hit/near-hit difference is visible
    const bool unbounded = kind == K_FLOAT || (kind == K_UNSIGNED && B > 1);
    const int small_px = (unbounded || p->interpolation == AUKIT_INTERP_NONE) ? PX_TABLE : PX_RATIONAL;
    pl.y = 1.0 / a.ratio;
    pl.quotient_fma_ok = big ? (aukit_quotient_fma_is_exact(a.ratio, 41) ? 1 : 0) : 0;
    pl.eps_r = (float)(fma(-(double)M, a.ratio, (double)L) / ((double)M * a.ratio));   // L/M / ratio_d - 1, numerator exact
    // position mode of a tile = f(largest input position it touches); tiles in which x crosses a power of
    // two >= 2^28 (the "+ 1" of A:666 rounds there) are evaluated exactly
    const double frames_per_tile = (double)pl.K * (double)pl.Q;
    auto tile_mode = [&](unsigned long long tile) -> int {
        if (pow2_ratio) return small_px;
        const double lo = (double)tile * frames_per_tile, hi = lo + frames_per_tile + 4.0;
        if (hi < 268435456.0) return small_px;
        if (hi >= 1518500249.0) return PX_BIG;                          // 2^30.5
        int elo = 0, ehi = 0;
        frexp(lo > 1.0 ? lo : 1.0, &elo);
        frexp(hi, &ehi);
        return elo != ehi ? PX_BIG : PX_MID;
    };
    // tile_mode() is piecewise constant and can only change next to a position threshold: collect
    // those tile indices as breakpoints and launch one kernel per run of equal mode (usually one).
    unsigned long long cuts[32];
    int ncuts = 0;
    cuts[ncuts++] = tile_first;
    if (!pow2_ratio) {
        const double thresholds[] = {268435456.0, 536870912.0, 1073741824.0, 1518500249.0};
        for (double v : thresholds) {
            const double b = floor(v / frames_per_tile);
            for (int d = -2; d <= 2; d++) {
                const double c = b + d;
                if (c > (double)tile_first && c <= (double)tile_last) cuts[ncuts++] = (unsigned long long)c;
            }
        }
    }
    for (int i = 1; i < ncuts; i++)                                      // insertion sort, <= 21 entries
        for (int j = i; j > 0 && cuts[j] < cuts[j - 1]; j--) { unsigned long long tmp = cuts[j]; cuts[j] = cuts[j - 1]; cuts[j - 1] = tmp; }
    int rc = 0;
    int i = 0;
    while (i < ncuts && !rc) {
        const unsigned long long seg = cuts[i];
        const int mode = tile_mode(seg);
        int j = i + 1;
        while (j < ncuts && (cuts[j] == seg || tile_mode(cuts[j]) == mode)) j++;   // merge runs of equal mode
        const unsigned long long end = (j < ncuts ? cuts[j] : tile_last + 1) - 1;
        pl.tile0 = seg;
        pl.ntiles = end - seg + 1;
        if (C == 1) rc = launch_interp<1>(ctx, a, pl, p->interpolation, mode, apply, threads, smem);
        else if (C == 2) rc = launch_interp<2>(ctx, a, pl, p->interpolation, mode, apply, threads, smem);
        else rc = launch_interp<0>(ctx, a, pl, p->interpolation, mode, apply, threads, smem);
        i = j;
    }
    return rc ? -1 : 1;
}
