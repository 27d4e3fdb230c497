// resample_planar.cu -- K5: Audio:resample (A:653-673) on planar float32 with TMA-staged, double-buffered tiles.
//
// The polyphase kernel of pipeline_poly.cu stages a tile, barriers, blends, barriers: ncu (profiles/r2_k5_ncu.txt)
// showed its warps waiting at the barrier and on the staging loads (5.8 + 5.4 stalled warps per issue), 0.62 - 0.69 of
// the copy rate.  Here the same tile geometry is kept (thread t owns outputs base + k*Sp + t, so its phase
// j_t = (t*M) mod L and its weights are loop invariants; tile = K iterations, K*Q + 3 input frames), but
//   * a tile's rows arrive by bulk async copies (TMA: cp.async.bulk + mbarrier), one per channel, into a 2-deep ring:
//     the copy of tile i+1 is issued right after the barrier that ends tile i-1 and lands while tile i is blended;
//   * there is ONE __syncthreads per tile (the j == 0 decision table of the next tile is written before it);
//   * CTAs are small (one period group, <= 256 threads) so that many are resident and their phases interleave;
//   * a whole tile's outputs are K*Sp CONTIGUOUS floats of each channel row: they are staged in shared memory (lane t
//     writes float t of every iteration: conflict-free) and leave as ONE bulk async store per channel (TMA,
//     cp.async.bulk shared -> global), double-buffered -- 4-byte STG per thread gave 0.77 of the copy rate, this 0.92.
// Arithmetic, weights, the exact-hit / near-hit decisions (A:666-667) and the NaN-transparent clamp (A:228, A:668) are
// those of poly_kernel's PX_TABLE mode, bit for bit.  One launch covers the whole range of a call: the first / last tiles
// of the signal or of the caller's window are staged with plain loads and a clamped index (the nil substitutions of
// A:259 / A:264) and masked to the requested outputs.  Positions >= 2^28 frames, L > 256 and non-integer rates stay
// with pipeline_poly.cu / resample.cu (tests: this kernel == polyphase kernel == one-thread-per-frame kernel).
#include "common.cuh"
#include "pipeline.cuh"

#include <math.h>
#include <stdlib.h>

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra WAIT_%=;\n\t}"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

enum { HIT = 0, NEAR_BELOW = 1, NEAR_ABOVE = 2 };

struct prs_args {
    const float *in;              // frames [in_first, ...) of each channel row
    size_t in_stride;
    unsigned long long in_first;
    int channels;
    double ratio;
    float *out;                   // output frame out_first of each channel row
    size_t out_stride;
    unsigned long long out_first;
    int L, M, m, Sp, Q, K, nfr;
    int pitch;                    // floats between the channel rows of a staged tile (multiple of 4)
    unsigned long long tile0, ntiles;
    unsigned long long n_total;   // frames of the whole signal
    size_t in_avail, n_out;       // frames held in `in`; outputs of this call
    int bulk_out;                 // whole tiles leave through a staging row per channel and ONE bulk async store each
};

// clamp of A:228-232 in two instructions: min.NaN / max.NaN return NaN when an operand is NaN, so NaN passes through
// exactly as the reference's two failed comparisons let it; +-Inf -> +-1, -0 stays -0
__device__ __forceinline__ float clamp_nan(float v) {
    float r;
    asm("min.NaN.f32 %0, %1, 0f3F800000;\n\tmax.NaN.f32 %0, %0, 0fBF800000;" : "=f"(r) : "f"(v));
    return r;
}

// CT: compile-time channel count (1, 2) or 0 = runtime
template <int MODE, int CT>
__global__ void __launch_bounds__(256) planar_resample_kernel(prs_args a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int C = CT ? CT : a.channels;
    float *bufs = reinterpret_cast<float *>(smem_raw);                               // [2][C][pitch]
    unsigned char *hit_tab = smem_raw + (size_t)2 * C * a.pitch * sizeof(float);     // [2][K*m]
    float *ostage = reinterpret_cast<float *>(smem_raw + (((size_t)2 * C * a.pitch * sizeof(float) + (size_t)2 * a.K * a.m + 127) & ~(size_t)127));   // [2][C][Sp*K], bulk_out only
    __shared__ __align__(8) uint64_t bars[2];
    const int t = threadIdx.x;
    const bool active = t < a.Sp;
    const long long tm = (long long)t * a.M;
    const int off_t = (int)(tm / a.L), j_t = (int)(tm % a.L);
    const bool is_j0 = (j_t == 0);
    float w0 = 0.f, w1 = 1.f, w2 = 0.f, w3 = 0.f, fx = 0.f;
    {
        const double x = (double)j_t / (double)a.L;
        fx = (float)x;
        if (MODE == AUKIT_INTERP_CUBIC) {                               // Catmull-Rom weights of A:265, fp64 then narrowed
            const double x2 = x * x, x3 = x2 * x;
            w0 = (float)(-0.5 * x3 + x2 - 0.5 * x);
            w1 = (float)(1.5 * x3 - 2.5 * x2 + 1.0);
            w2 = (float)(-1.5 * x3 + 2.0 * x2 + 0.5 * x);
            w3 = (float)(0.5 * x3 - 0.5 * x2);
        }
    }
    const unsigned long long tile_out = (unsigned long long)a.Sp * a.K;
    const int ntab = a.K * a.m;
    const int K = a.K, Q = a.Q, Sp = a.Sp, pitch = a.pitch;
    const int jrow = t / a.L;                                           // which of the m periods of an iteration this thread is in

    const long long n_total = (long long)a.n_total, in_lo = (long long)a.in_first, in_hi = in_lo + (long long)a.in_avail;
    const unsigned long long out_lo = a.out_first, out_hi = a.out_first + a.n_out;
    // a tile whose staged frames (aligned start to rounded-up end) all exist in the window travels by bulk copy; the
    // first / last tiles of the signal or of the caller's window are staged with plain loads and a clamped index
    // (= the nil substitutions of A:259 / A:264)
    auto interior = [&](unsigned long long tile) {
        const long long gA = (long long)(tile * (unsigned long long)a.K * (unsigned long long)a.Q) - 1;
        if (gA < in_lo) return false;
        const long long a0 = (gA - in_lo) & ~3ll;
        const long long last = in_lo + a0 + (long long)((a.nfr + (int)((gA - in_lo) - a0) + 3) / 4) * 4;
        return last <= in_hi && last <= n_total;
    };
    // thread 0: bulk copies of one tile's rows, from a 16-byte aligned start
    auto issue = [&](unsigned long long tile, int buf) {
        const long long gA = (long long)(tile * (unsigned long long)a.K * (unsigned long long)a.Q) - 1;   // first frame the tile needs
        const size_t foff = (size_t)(gA - (long long)a.in_first);
        const size_t a0 = foff & ~(size_t)3;
        const uint32_t bytes = (uint32_t)((a.nfr + (int)(foff - a0) + 3) / 4) * 16u;
        mbar_expect_tx(&bars[buf], bytes * (uint32_t)C);
        for (int c = 0; c < C; c++)
            bulk_load(bufs + ((size_t)buf * C + c) * a.pitch, a.in + (size_t)c * a.in_stride + a0, bytes, &bars[buf]);
    };
    // exact hit / near-hit decision for a tile's j == 0 outputs with the reference's own fp64 expression (A:666-667)
    auto decide = [&](unsigned long long tile, int buf) {
        const unsigned long long base_out = tile * tile_out;
        const long long F0 = (long long)(tile * (unsigned long long)a.K * (unsigned long long)a.Q);
        for (int e = t; e < ntab; e += blockDim.x) {
            const unsigned long long n = base_out + (unsigned long long)(e / a.m) * a.Sp + (unsigned long long)(e % a.m) * a.L;
            const double xt = (double)(F0 + (long long)(e / a.m) * a.Q + (long long)(e % a.m) * a.M + 1);
            const double x = __dadd_rn(__ddiv_rn((double)n, a.ratio), 1.0);
            hit_tab[buf * ntab + e] = (x == xt) ? HIT : (x < xt ? NEAR_BELOW : NEAR_ABOVE);
        }
    };
    // one channel of one output: taps at f[0..3] (p1 = f[1]), st = decision for a j == 0 output
    auto value = [&](const float *f, int st) -> float {
        const float p1 = f[1];
        const float p0 = (MODE == AUKIT_INTERP_LINEAR) ? 0.f : f[0];    // `none` loads it unconditionally: selects, not branches, below
        float v;
        if (MODE == AUKIT_INTERP_CUBIC) v = __fmaf_rn(w3, f[3], __fmaf_rn(w2, f[2], __fmaf_rn(w1, p1, w0 * p0)));
        else if (MODE == AUKIT_INTERP_LINEAR) v = __fmaf_rn(f[2] - p1, fx, p1);
        else v = p1;
        float r = clamp_nan(v);
        if (is_j0) {
            // one output per period sits on (or one ulp beside) an input frame: decided exactly
            const float below = (MODE == AUKIT_INTERP_NONE && st == NEAR_BELOW) ? p0 : p1;   // `none`: floor(x) is one lower
            r = st == HIT ? p1 : clamp_nan(below);               // HIT: copied unclamped, A:667; else weights (0, 1, 0, 0)
        }
        return r;
    };

    unsigned long long tile = a.tile0 + blockIdx.x;
    if (tile >= a.tile0 + a.ntiles) return;
    if (t == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (interior(tile)) issue(tile, 0);
    }
    decide(tile, 0);
    __syncthreads();
    uint32_t ph0 = 0u, ph1 = 0u;
    for (int it = 0; tile < a.tile0 + a.ntiles; it++, tile += gridDim.x) {
        const int buf = it & 1;
        const unsigned long long next = tile + gridDim.x;
        const bool has_next = next < a.tile0 + a.ntiles;
        // the other buffer (tile it - 1) and its table are free: every thread passed the barrier that ended that tile
        if (has_next) {
            if (t == 0 && interior(next)) issue(next, buf ^ 1);
            decide(next, buf ^ 1);
        }
        const long long gA = (long long)(tile * (unsigned long long)a.K * (unsigned long long)a.Q) - 1;
        int sh;                                                                       // staged index of frame gA
        const bool whole = tile * tile_out >= out_lo && (tile + 1) * tile_out <= out_hi;
        if (interior(tile)) {
            mbar_wait(&bars[buf], buf ? ph1 : ph0);
            if (buf) ph1 ^= 1u; else ph0 ^= 1u;
            sh = (int)((size_t)(gA - in_lo) & 3);
        } else {
            sh = 0;
            for (int c = 0; c < C; c++) {
                float *dst = bufs + ((size_t)buf * C + c) * pitch;
                const float *rowp = a.in + (size_t)c * a.in_stride;
                for (int f = t; f < a.nfr; f += blockDim.x) {
                    long long gi = gA + f;
                    gi = gi < 0 ? 0 : (gi >= n_total ? n_total - 1 : gi);
                    dst[f] = (gi >= in_lo && gi < in_hi) ? rowp[gi - in_lo] : 0.0f;   // outside the window: not needed by any output of this call
                }
            }
            __syncthreads();
        }
        const bool bulk = a.bulk_out && whole;
        const int tile_n = Sp * K;
        if (active && bulk) {
            // whole tile: outputs go to a staging row per channel (consecutive lanes, consecutive floats) ...
            float *os = ostage + (size_t)buf * C * tile_n + t;
            const float *f = bufs + (size_t)buf * C * pitch + off_t + sh;
            const unsigned char *ht = hit_tab + buf * ntab + jrow;
#pragma unroll 4
            for (int k = 0; k < K; k++) {
                const int st = ht[k * a.m];                          // read by every thread (one byte, no branch); only j == 0 threads use it
                if (CT == 1) {
                    os[0] = value(f, st);
                } else if (CT == 2) {
                    const float vl = value(f, st), vr = value(f + pitch, st);
                    os[0] = vl;
                    os[tile_n] = vr;
                } else {
                    for (int c = 0; c < C; c++) os[c * tile_n] = value(f + c * pitch, st);
                }
                f += Q;
                os += Sp;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");            // my staging writes before the copy engine's reads
        } else if (active) {
            const float *f = bufs + (size_t)buf * C * pitch + off_t + sh;            // taps of output k = 0, channel 0
            const unsigned char *ht = hit_tab + buf * ntab + jrow;
            const unsigned long long o0 = tile * tile_out + t;                        // global index of this thread's output k = 0
            float *outp = a.out + (size_t)((long long)o0 - (long long)a.out_first);
#pragma unroll 4
            for (int k = 0; k < K; k++) {
                if (!whole) {                                                         // first / last tile of a range
                    const unsigned long long o = o0 + (unsigned long long)k * Sp;
                    if (o < out_lo || o >= out_hi) { f += Q; outp += Sp; continue; }
                }
                const int st = ht[k * a.m];                          // read by every thread (one byte, no branch); only j == 0 threads use it
                if (CT == 1) {
                    outp[0] = value(f, st);
                } else if (CT == 2) {
                    const float vl = value(f, st), vr = value(f + pitch, st);
                    outp[0] = vl;
                    outp[a.out_stride] = vr;
                } else {
                    for (int c = 0; c < C; c++) outp[(size_t)c * a.out_stride] = value(f + c * pitch, st);
                }
                f += Q;
                outp += Sp;
            }
        }
        // the staging rows written one tile ago must have been read before the NEXT tile overwrites them
        if (t == 0 && a.bulk_out) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncthreads();                                                // tile done: its buffer and table may be refilled
        if (t == 0 && bulk) {
            // ... and leave as ONE bulk async store per channel (TMA): tile_n contiguous floats of the channel's row
            const float *src = ostage + (size_t)buf * C * tile_n;
            float *dst = a.out + (size_t)(tile * tile_out - a.out_first);
            for (int c = 0; c < C; c++)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + (size_t)c * a.out_stride),
                             "r"(smem_u32(src + (size_t)c * tile_n)), "r"((uint32_t)tile_n * 4u) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (t == 0 && a.bulk_out) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ---- sinc (A:267-281): 21 taps per output.  Same tile geometry and TMA ring; a thread's 21 weights and their derivatives
// (fp64 sinc at its rational phase j_t / L, narrowed) are loop invariants, corrected per output to first order for the
// drift of the reference's fp64 position (x_ref - n*M/L = x * eps_r) -- the same table-plus-correction the
// one-thread-per-frame kernel uses, minus its per-output table loads and global taps (45 G samples/s).  This kernel takes
// the WHOLE output range of a call, edge tiles included (frames outside the signal are staged as zeros: the reference
// skips those taps, A:271; a zero tap adds exactly nothing), so a range gives the same bits however it is sharded.
struct sinc_args {
    prs_args g;
    float eps_r;
};

template <int CT>
__global__ void __launch_bounds__(256) planar_sinc_kernel(sinc_args sa) {
    const prs_args &a = sa.g;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int C = CT ? CT : a.channels;
    float *bufs = reinterpret_cast<float *>(smem_raw);                               // [2][C][pitch]
    unsigned char *hit_tab = smem_raw + (size_t)2 * C * a.pitch * sizeof(float);     // [2][K*m]
    __shared__ __align__(8) uint64_t bars[2];
    const int t = threadIdx.x;
    const bool active = t < a.Sp;
    const long long tm = (long long)t * a.M;
    const int off_t = (int)(tm / a.L), j_t = (int)(tm % a.L);
    const bool is_j0 = (j_t == 0);
    float w[21], wd[21];
    const double fxd = (double)j_t / (double)a.L;
#pragma unroll
    for (int k = 0; k < 21; k++) {                                      // A:273-276 at fx = j / L, and d/dfx
        const double px = 3.14159265358979323846 * (fxd - (double)(k - 10));
        if (px == 0.0) { w[k] = 1.0f; wd[k] = 0.0f; }
        else {
            const double sn = sin(px), cs = cos(px);
            w[k] = (float)(sn / px);
            wd[k] = (float)(3.14159265358979323846 * (cs * px - sn) / (px * px));
        }
    }
    const float fx = (float)fxd;
    const unsigned long long tile_out = (unsigned long long)a.Sp * a.K;
    const int ntab = a.K * a.m;
    const int K = a.K, Q = a.Q, Sp = a.Sp, pitch = a.pitch;
    const int jrow = t / a.L;
    const long long n_total = (long long)a.n_total, in_lo = (long long)a.in_first, in_hi = in_lo + (long long)a.in_avail;
    const unsigned long long out_lo = a.out_first, out_hi = a.out_first + a.n_out;

    auto first_frame = [&](unsigned long long tile) { return (long long)(tile * (unsigned long long)a.K * (unsigned long long)a.Q) - 10; };
    // a tile whose staged frames (aligned start to rounded-up end) all exist in the window travels by bulk copy
    auto interior = [&](unsigned long long tile) {
        const long long gA = first_frame(tile);
        if (gA < in_lo) return false;
        const long long a0 = (gA - in_lo) & ~3ll;
        const long long last = in_lo + a0 + (long long)((a.nfr + (int)((gA - in_lo) - a0) + 3) / 4) * 4;
        return last <= in_hi && last <= n_total;
    };
    auto issue = [&](unsigned long long tile, int buf) {
        const size_t foff = (size_t)(first_frame(tile) - in_lo);
        const size_t a0 = foff & ~(size_t)3;
        const uint32_t bytes = (uint32_t)((a.nfr + (int)(foff - a0) + 3) / 4) * 16u;
        mbar_expect_tx(&bars[buf], bytes * (uint32_t)C);
        for (int c = 0; c < C; c++)
            bulk_load(bufs + ((size_t)buf * C + c) * a.pitch, a.in + (size_t)c * a.in_stride + a0, bytes, &bars[buf]);
    };
    auto decide = [&](unsigned long long tile, int buf) {
        const unsigned long long base_out = tile * tile_out;
        const long long F0 = (long long)(tile * (unsigned long long)a.K * (unsigned long long)a.Q);
        for (int e = t; e < ntab; e += blockDim.x) {
            const unsigned long long n = base_out + (unsigned long long)(e / a.m) * a.Sp + (unsigned long long)(e % a.m) * a.L;
            const double xt = (double)(F0 + (long long)(e / a.m) * a.Q + (long long)(e % a.m) * a.M + 1);
            const double x = __dadd_rn(__ddiv_rn((double)n, a.ratio), 1.0);
            hit_tab[buf * ntab + e] = (x == xt) ? HIT : (x < xt ? NEAR_BELOW : NEAR_ABOVE);
        }
    };

    unsigned long long tile = a.tile0 + blockIdx.x;
    if (tile >= a.tile0 + a.ntiles) return;
    if (t == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (interior(tile)) issue(tile, 0);
    }
    decide(tile, 0);
    __syncthreads();
    uint32_t ph0 = 0u, ph1 = 0u;
    for (int it = 0; tile < a.tile0 + a.ntiles; it++, tile += gridDim.x) {
        const int buf = it & 1;
        const unsigned long long next = tile + gridDim.x;
        if (next < a.tile0 + a.ntiles) {
            if (t == 0 && interior(next)) issue(next, buf ^ 1);
            decide(next, buf ^ 1);
        }
        const long long gA = first_frame(tile);
        int sh;
        if (interior(tile)) {
            mbar_wait(&bars[buf], buf ? ph1 : ph0);
            if (buf) ph1 ^= 1u; else ph0 ^= 1u;
            sh = (int)((size_t)(gA - in_lo) & 3);
        } else {
            // edge tile: plain loads; frames outside the signal are zeros (taps the reference skips, A:271)
            sh = 0;
            for (int c = 0; c < C; c++) {
                float *dst = bufs + ((size_t)buf * C + c) * pitch;
                const float *row = a.in + (size_t)c * a.in_stride;
                for (int f = t; f < a.nfr; f += blockDim.x) {
                    const long long gi = gA + f;
                    dst[f] = (gi >= 0 && gi < n_total && gi >= in_lo && gi < in_hi) ? row[gi - in_lo] : 0.0f;
                }
            }
            __syncthreads();
        }
        if (active) {
            const float *f = bufs + (size_t)buf * C * pitch + off_t + sh;            // tap k = -10 of output k = 0, channel 0
            const unsigned char *ht = hit_tab + buf * ntab + jrow;
            const unsigned long long o0 = tile * tile_out + t;                        // global index of this thread's output k = 0
            float *outp = a.out + (size_t)((long long)o0 - (long long)a.out_first);
            float xf = (float)((long long)(tile * (unsigned long long)a.K * (unsigned long long)a.Q) + off_t) + fx;   // rational position
            const float qf = (float)Q;
            for (int k = 0; k < K; k++, f += Q, outp += Sp, xf += qf) {
                const unsigned long long o = o0 + (unsigned long long)k * Sp;
                if (o < out_lo || o >= out_hi) continue;                              // first / last tile of a range
                const float dl = xf * sa.eps_r;                                       // x_ref - x_rational
                float cw[21];
#pragma unroll
                for (int q = 0; q < 21; q++) cw[q] = __fmaf_rn(wd[q], dl, w[q]);
                const bool hit = is_j0 && ht[k * a.m] == HIT;
                if (CT == 2) {
                    float sl = 0.f, sr = 0.f;
#pragma unroll
                    for (int q = 0; q < 21; q++) { sl = __fmaf_rn(f[q], cw[q], sl); sr = __fmaf_rn(f[pitch + q], cw[q], sr); }
                    outp[0] = hit ? f[10] : clamp_nan(sl);                            // exact hit: copied unclamped, A:667
                    outp[a.out_stride] = hit ? f[pitch + 10] : clamp_nan(sr);
                } else {
                    for (int c = 0; c < C; c++) {
                        const float *fc = f + c * pitch;
                        float sum = 0.f;
#pragma unroll
                        for (int q = 0; q < 21; q++) sum = __fmaf_rn(fc[q], cw[q], sum);
                        outp[(size_t)c * a.out_stride] = hit ? fc[10] : clamp_nan(sum);
                    }
                }
            }
        }
        __syncthreads();
    }
}

long long gcd_ll(long long x, long long y) { while (y) { long long r = x % y; x = y; y = r; } return x; }

}  // namespace

// Returns 1 when the range was produced, 0 when this path does not apply, -1 on error.
int aukit_planar_resample_try(aukit_ctx *ctx, const float *d_in, size_t in_stride, int channels, unsigned long long n_in_total,
                              unsigned long long in_first, size_t in_avail, double srcRate, double dstRate, int interpolation,
                              unsigned long long out_first, size_t n_out, float *d_out, size_t out_stride) {
    // AUKIT_DISABLE_PLANAR=1 (diagnostic, tests): everything through the polyphase / per-frame kernels
    const char *dis = getenv("AUKIT_DISABLE_PLANAR");
    const bool disabled = dis && dis[0] == '1';
    if (disabled || interpolation < AUKIT_INTERP_NONE || interpolation > AUKIT_INTERP_CUBIC) return 0;
    if (channels < 1 || channels > 8 || ((uintptr_t)d_in & 15) || (channels > 1 && (in_stride & 3))) return 0;
    const double sr = srcRate, dr = dstRate;
    if (!(sr >= 1 && dr >= 1 && sr < 2147483648.0 && dr < 2147483648.0) || sr != floor(sr) || dr != floor(dr)) return 0;
    const long long g = gcd_ll((long long)sr, (long long)dr);
    const long long L = (long long)dr / g, M = (long long)sr / g;
    if (L < 2 || L > 256 || M > 4096) return 0;                         // L == 1: the strided-gather kernels (pipeline_decim.cu)
    const double ratio = dstRate / srcRate;
    if ((double)(out_first + n_out) * (double)M / (double)L >= 268435456.0) return 0;   // rational-position regime only
    prs_args a{};
    a.L = (int)L; a.M = (int)M;
    a.m = (int)(256 / L);
    if ((long long)a.m * M > 4096) a.m = (int)(4096 / M);
    if (a.m < 1) a.m = 1;
    a.Sp = a.L * a.m;
    a.Q = a.M * a.m;
    const int threads = (a.Sp + 31) / 32 * 32;
    const size_t budget = 12 * 1024;                                     // per input buffer (the output staging doubles it): small tiles, many resident CTAs
    long long K = ((long long)(budget / (sizeof(float) * (size_t)channels)) - 12) / a.Q;
    if (K < 1) K = 1;
    if (K > 64) K = 64;
    a.K = (int)K;
    a.nfr = a.K * a.Q + 3;
    a.pitch = (a.nfr + 3 + 3) / 4 * 4 + 4;
    const unsigned long long tile_out = (unsigned long long)a.Sp * a.K;
    // bulk stores need 16-byte aligned tile rows in global memory
    a.bulk_out = (((uintptr_t)d_out & 15) == 0 && (channels == 1 || (out_stride & 3) == 0) && (tile_out & 3) == 0 && (out_first & 3) == 0) ? 1 : 0;
    size_t smem = (size_t)2 * channels * a.pitch * sizeof(float) + (size_t)2 * a.K * a.m + 16;
    if (a.bulk_out) smem = ((smem + 127) & ~(size_t)127) + (size_t)2 * channels * tile_out * sizeof(float);
    if (smem > 100 * 1024) return 0;
    // positions past the end of the signal (absurd ratios) are the caller's / the per-frame kernel's business
    if (n_out < 4 * tile_out) return 0;                                 // short ranges: the polyphase kernel's tiles are as good
    a.in = d_in; a.in_stride = in_stride; a.in_first = in_first; a.channels = channels; a.ratio = ratio;
    a.out = d_out; a.out_stride = out_stride; a.out_first = out_first;
    a.tile0 = out_first / tile_out;
    a.ntiles = (out_first + n_out - 1) / tile_out - a.tile0 + 1;
    a.n_total = n_in_total; a.in_avail = in_avail; a.n_out = n_out;
    auto go = [&](auto kern) -> int {
        if (aukit_cuda_check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "smem attr")) return -1;
        int occ = 0;
        if (aukit_cuda_check(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem), "occupancy")) return -1;
        if (occ < 1) return aukit_fail("aukit_cuda: planar resample kernel does not fit on an SM");
        unsigned long long grid = (unsigned long long)ctx->num_sms * occ;
        if (grid > a.ntiles) grid = a.ntiles;
        kern<<<(unsigned)grid, threads, smem, ctx->stream>>>(a);
        ctx->launches++;
        return aukit_cuda_check(cudaGetLastError(), "planar_resample_kernel launch");
    };
    int rc;
#define AUKIT_PRS(MODE) (channels == 1 ? go(planar_resample_kernel<MODE, 1>) : channels == 2 ? go(planar_resample_kernel<MODE, 2>) \
                                                                                              : go(planar_resample_kernel<MODE, 0>))
    if (interpolation == AUKIT_INTERP_NONE) rc = AUKIT_PRS(AUKIT_INTERP_NONE);
    else if (interpolation == AUKIT_INTERP_LINEAR) rc = AUKIT_PRS(AUKIT_INTERP_LINEAR);
    else rc = AUKIT_PRS(AUKIT_INTERP_CUBIC);
#undef AUKIT_PRS
    return rc ? -1 : 1;
}

// interpolate.sinc on planar float32, the whole range of a call.  Returns 1 handled, 0 not applicable, -1 error.
int aukit_planar_sinc_try(aukit_ctx *ctx, const float *d_in, size_t in_stride, int channels, unsigned long long n_in_total,
                          unsigned long long in_first, size_t in_avail, double srcRate, double dstRate,
                          unsigned long long out_first, size_t n_out, float *d_out, size_t out_stride) {
    const char *dis = getenv("AUKIT_DISABLE_PLANAR");
    if (dis && dis[0] == '1') return 0;
    if (channels < 1 || channels > 8 || ((uintptr_t)d_in & 15) || (channels > 1 && (in_stride & 3))) return 0;
    const double sr = srcRate, dr = dstRate;
    if (!(sr >= 1 && dr >= 1 && sr < 2147483648.0 && dr < 2147483648.0) || sr != floor(sr) || dr != floor(dr)) return 0;
    const long long g = gcd_ll((long long)sr, (long long)dr);
    const long long L = (long long)dr / g, M = (long long)sr / g;
    if (L > 256 || M > 4096) return 0;
    const double ratio = dstRate / srcRate;
    // the first-order correction covers the systematic drift; what is left (the quotient's rounding, <= x * 2^-53) must stay
    // negligible against 2^-20 over 21 taps: positions below 2^27 frames
    if ((double)(out_first + n_out) * (double)M / (double)L >= 134217728.0) return 0;
    sinc_args sa{};
    prs_args &a = sa.g;
    a.L = (int)L; a.M = (int)M;
    a.m = (int)(128 / L);
    if (a.m < 1) a.m = 1;
    if ((long long)a.m * M > 4096) a.m = (int)(4096 / M) > 0 ? (int)(4096 / M) : 1;
    a.Sp = a.L * a.m;
    a.Q = a.M * a.m;
    const int threads = (a.Sp + 31) / 32 * 32;
    if (threads > 256) return 0;
    const size_t budget = 24 * 1024;
    long long K = ((long long)(budget / (sizeof(float) * (size_t)channels)) - 32) / a.Q;
    if (K < 1) K = 1;
    if (K > 64) K = 64;
    a.K = (int)K;
    a.nfr = a.K * a.Q + 21;
    a.pitch = (a.nfr + 3 + 3) / 4 * 4 + 4;
    const size_t smem = (size_t)2 * channels * a.pitch * sizeof(float) + (size_t)2 * a.K * a.m + 16;
    if (smem > 160 * 1024) return 0;
    const unsigned long long tile_out = (unsigned long long)a.Sp * a.K;
    a.in = d_in; a.in_stride = in_stride; a.in_first = in_first; a.channels = channels; a.ratio = ratio;
    a.out = d_out; a.out_stride = out_stride; a.out_first = out_first;
    a.tile0 = out_first / tile_out;
    a.ntiles = (out_first + n_out - 1) / tile_out - a.tile0 + 1;
    a.n_total = n_in_total; a.in_avail = in_avail; a.n_out = n_out;
    sa.eps_r = (float)(fma(-(double)M, ratio, (double)L) / ((double)M * ratio));
    auto go = [&](auto kern) -> int {
        if (aukit_cuda_check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "smem attr")) return -1;
        int occ = 0;
        if (aukit_cuda_check(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem), "occupancy")) return -1;
        if (occ < 1) return aukit_fail("aukit_cuda: planar sinc kernel does not fit on an SM");
        unsigned long long grid = (unsigned long long)ctx->num_sms * occ;
        if (grid > a.ntiles) grid = a.ntiles;
        kern<<<(unsigned)grid, threads, smem, ctx->stream>>>(sa);
        ctx->launches++;
        return aukit_cuda_check(cudaGetLastError(), "planar_sinc_kernel launch");
    };
    const int rc = channels == 1 ? go(planar_sinc_kernel<1>) : channels == 2 ? go(planar_sinc_kernel<2>) : go(planar_sinc_kernel<0>);
    return rc ? -1 : 1;
}
