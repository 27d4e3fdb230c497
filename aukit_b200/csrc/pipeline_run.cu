// pipeline_run.cu -- "run per lane" kernels of the fused pipeline (K10) for the headline case:
// 16-bit little-endian STEREO input, cubic interpolation, mono mixdown, integer rates with
// L = new/gcd a multiple of 32 and M = old/gcd odd (44.1 / 22.05 / 11.025 kHz -> 48 kHz).
//
// Two kernels live here:
//   run_static_kernel  (second half of the file)  44.1 -> 48 kHz: L = 160, M = 147 as template parameters, one class
//                      of half periods per warp, the whole half period a straight line of code, a CTA-wide ring of
//                      TMA-filled frame buffers.  This is the kernel bench.py measures.
//   run_kernel         (first half)  the scripted predecessor, kept for 22.05 / 11.025 -> 48 kHz and for A/B runs
//                      (AUKIT_RUN_STATIC=0): both classes of half periods in one warp, a byte script per frame.
//
// Why "run per lane" at all: pipeline_poly.cu maps lanes to consecutive outputs, so every output re-reads
// its 4 taps from shared memory (32 B of the 128 B/clk/SM) -- measured 65 % shared-memory pipe,
// 60 % issue, 42 instructions per output.  Here a lane owns a RUN of consecutive outputs of one PERIOD of the
// rational resampling pattern (period p of a tile = outputs [p*L, (p+1)*L), input frames [p*M - 1, (p+1)*M + 2)).
// All lanes of a class are therefore at the SAME phase at the same step:
//   * the 4 Catmull-Rom weights of a step are uniform: one broadcast LDS.128 from a table
//     built once per CTA (in fp64, narrowed), not per-lane registers or per-lane reads;
//   * the taps slide in registers: each input frame is loaded from shared memory and converted
//     ONCE per lane (0.92 loads per output instead of 4), with row stride M words (odd => the 32
//     lanes hit 32 different banks);
//   * whether a new input frame yields 1 or 2 outputs is uniform too (a byte script per frame in run_kernel,
//     compile-time constants in run_static_kernel); the register slots are addressed statically, so no moves.
// Input tiles (contiguous frames) are staged by ONE bulk async copy each (TMA, cp.async.bulk + mbarrier).
// Outputs are transposed through a small XOR-swizzled shared staging tile and written as full 64-byte row
// segments (st.global.L1::no_allocate.v4), so stores stay coalesced although a lane's outputs are L samples
// apart from its neighbour's.
//
// Positions: rational (n*M/L) plus a drift term that is constant over a segment of a fixed global grid (delta = x*eps_r, see
// pipeline_poly.cu / DESIGN.md 3.2); the grid keeps delta within 11 % of its
// true value.  Tiles that touch the ends of the signal or of the shard, other formats / modes and
// exactness-sensitive cases stay on pipeline_poly.cu's kernels.
#include "common.cuh"
#include "pipeline.cuh"

#include <math.h>

namespace {

struct run_plan {
    int L, M;                     // outputs / input frames per lane (one period)
    int raw_words;                // 32-bit words of raw input per warp buffer (multiple of 32)
    int nwarps;
    float delta;                  // drift of the reference's position in this launch's range (frames)
    unsigned long long tile0;     // first warp tile (32*L outputs each) of this launch
    unsigned long long ntiles;
    int nbuf;                     // static kernel: frame buffers in the CTA's ring
    int nbuf2;                    // ... of a CTA whose tiles span two drift segments (a second weight table is resident)
    // static kernel: the launch's tile range as up to RUN_MAXSEG segments of constant drift (the weight table of a
    // segment is rebuilt in place by the first warp that reaches it -- one launch per pass instead of one per segment,
    // which cost a drain + refill of the persistent kernel each, ~8 us)
    int nseg;
    unsigned long long seg_end[8];   // first tile (global index) past segment k
    float seg_delta[8];
};
constexpr int RUN_MAXSEG = 8;

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
// global -> shared bulk async copy (TMA), completion on an mbarrier
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- packed f32x2 helpers (sm_100 FFMA2: one issue slot for both channels)
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float x, float y) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float &x, float &y) { asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// One stereo s16 frame -> (L, R) in the SCALED domain v * 2^15: u + max(u, 0) / 32767 is the correctly
// rounded s * 32768 / 32767, i.e. 2^15 times the exactly rounded sample of A:1133 (same proof as
// common.cuh::s16_to_float, scaled by a power of two).  The 2^-15 is folded into the final scale.
// ALU_MAX: max(u, 0) as an FMNMX on the ALU pipe instead (for kernels whose FMA pipe is the busier one).
template <bool ALU_MAX = false>
__device__ __forceinline__ f32x2 cvt_frame(uint32_t w) {
    const float ul = (float)(int)(int16_t)(w & 0xFFFFu), ur = (float)((int)w >> 16);
    constexpr float c = 0.5f / 32767.0f;
    // max(u, 0) = (u + |u|) / 2, evaluated as an FADD with an |.| operand (FMA pipe, exact) instead of an
    // FMNMX (ALU pipe, the busier one here); the 1/2 is folded into c
    if (ALU_MAX) return fma2(pack2(fmaxf(ul, 0.f), fmaxf(ur, 0.f)), pack2(2.0f * c, 2.0f * c), pack2(ul, ur));
    return fma2(pack2(ul + fabsf(ul), ur + fabsf(ur)), pack2(c, c), pack2(ul, ur));
}

// clamp(v, -32768, 32768) of A:668 in the scaled domain as ONE instruction: sign(v) * min(|v|, 32768)
// (min.xorsign.abs: the result takes the XOR of the operands' signs, the bound is positive).  Inputs are finite here.
__device__ __forceinline__ float clamp_scaled(float v) {
    float r;
    asm("min.xorsign.abs.f32 %0, %1, %2;" : "=f"(r) : "f"(v), "f"(32768.0f));
    return r;
}

constexpr int STAGE_COLS = 16;        // outputs per lane per flush (64 bytes per lane-row)
constexpr int STAGE_RING = 32;        // staged columns per lane (ring): the two classes drift by a few outputs
constexpr int STAGE_STRIDE = 36;      // floats per staging row: 16-byte aligned rows
constexpr int PERIODS = 16;           // periods per warp tile: two lanes (half-periods) per period

// The whole warp flushes 16 staged outputs of every lane-row: transpose back so that 4 lanes write 64
// contiguous bytes of one row.  Row r = lane index that produced it (class r >> 4, period r & 15); columns are
// XOR-swizzled by 4 * ((r >> 3) & 3) so that the column writes of 32 lanes hit 32 banks.
// Kept out of line on purpose: inlined, the compiler hoists the address computations into every output step.
__device__ __noinline__ void flush_stage(const float *stage, float *g_tile, int L, int col0, int lane) {
    __syncwarp();
    const int q = lane & 3;
#pragma unroll
    for (int it = 0; it < 4; it++) {
        const int r = (lane >> 2) + 8 * it;
        const float4 v = *reinterpret_cast<const float4 *>(stage + r * STAGE_STRIDE + (col0 & (STAGE_RING - 1)) + 4 * (q ^ ((r >> 3) & 3)));
        stg_stream(reinterpret_cast<float4 *>(g_tile + (size_t)(r & 15) * L + (size_t)(r >> 4) * (L >> 1) + col0 + 4 * q), v);
    }
    __syncwarp();
}

// Lane (h, p): class h = lane >> 4 owns HALF of period p = lane & 15 of the warp tile: outputs
// [h*L/2, (h+1)*L/2) of the period.  All lanes of a class are at the same phase at the same step, the two
// classes half a period apart (two weight addresses per LDS.128, two script bytes).  Halving the run per lane
// halves the shared memory per warp, which is what bounds the number of resident warps.
// CMIN = floor(L / M): every new input frame yields CMIN outputs, some one more (bit flags per 8 frames).
// CLAMP1: keep the final clamp(v * peak/max, -1, 1) of A:3455.  |v| <= max for every output, so for
// peakAmplitude < 1 - 2^-20 it can never act and the apply pass drops it (two ALU-pipe instructions per output).
template <bool APPLY, int CMIN, bool CLAMP1>
__global__ void __launch_bounds__(768, 1) run_kernel(pipe_args a, run_plan rp) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int L = rp.L, M = rp.M, LH = L / 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = lane >> 4, pp = lane & 15;
    // per-class geometry
    const int e_begin = h * LH;
    const int fbase = (int)(((long long)e_begin * M) / L);                 // floor position of the class's first output
    const int nsteps0 = (int)(((long long)(LH - 1) * M) / L) + 1;         // distinct floors of class 0 / class 1
    const int nsteps1 = (int)(((long long)(L - 1) * M) / L) - (int)(((long long)LH * M) / L) + 1;
    const int NSmin = nsteps0 < nsteps1 ? nsteps0 : nsteps1, NSmax = nsteps0 < nsteps1 ? nsteps1 : nsteps0;
    const int G = NSmin / 8;                                               // full groups of 8 steps both classes share
    const int tail_steps = NSmax - 8 * G;                                  // <= 15, handled with explicit counts
    const int SCR = 64;                                                    // bytes per class: flags[G] then tail counts
    // layout: weights[L] float4 | script[2][SCR] | mbar[nwarps] | per-warp: raw words | staging
    float4 *W = reinterpret_cast<float4 *>(smem);
    unsigned char *script = smem + (size_t)L * 16;
    uint64_t *bars = reinterpret_cast<uint64_t *>(script + 2 * SCR);
    unsigned char *warp_base = reinterpret_cast<unsigned char *>(bars) + (((size_t)rp.nwarps * 8 + 127) & ~(size_t)127);
    const size_t per_warp = (size_t)rp.raw_words * 4 + (APPLY ? 32 * STAGE_STRIDE * 4 : 0);
    uint32_t *raw = reinterpret_cast<uint32_t *>(warp_base + (size_t)warp * per_warp);
    float *stage = reinterpret_cast<float *>(raw + rp.raw_words);

    // ---- one-time tables (fp64 weights of A:265 at fraction j/L + delta, narrowed)
    for (int e = threadIdx.x; e < L; e += blockDim.x) {
        const int j = (int)(((long long)e * M) % L);
        const double x = (double)j / (double)L + (double)rp.delta;
        const double x2 = x * x, x3 = x2 * x;
        W[e] = make_float4((float)(-0.5 * x3 + x2 - 0.5 * x), (float)(1.5 * x3 - 2.5 * x2 + 1.0),
                           (float)(-1.5 * x3 + 2.0 * x2 + 0.5 * x), (float)(0.5 * x3 - 0.5 * x2));
    }
    // Output scale of the apply pass: peak/max (A:3444) * 2^-15 (sample scale) * 1/2 (mono mean, A:687).  It stays a
    // separate multiply (not folded into the weights) so that these outputs are bit-identical to the polyphase
    // kernel's, which is what makes a time-sharded run equal a single pass bit for bit.  max == 0 (silence): the
    // reference computes 0 * inf = NaN for every sample and its clamp lets NaN through; a NaN scale (and NaN clamp
    // bounds: fmin/fmax of two NaNs stay NaN) reproduces that.
    float mult = 0.f, one_hi = 1.0f;
    if (APPLY) {
        const float mx0 = a.d_max[0];
        mult = mx0 > 0.f ? (float)(a.peak / (double)mx0) * (1.0f / 65536.0f) : __int_as_float(0x7FC00000);
        if (!(mx0 > 0.f)) one_hi = mult;
    }
    const float one_lo = -one_hi;
    // count(cls, s) = outputs of class cls whose floor position is fbase_cls + s  (s = step index, 0-based)
    auto count_at = [&](int cls, int s) -> int {
        const int eb = cls * LH, ee = eb + LH;
        const long long f = ((long long)eb * M) / L + s;
        long long lo = (f * L + M - 1) / M, hi = ((f + 1) * L + M - 1) / M;  // e with floor(e*M/L) == f
        if (lo < eb) lo = eb;
        if (hi > ee) hi = ee;
        return hi > lo ? (int)(hi - lo) : 0;
    };
    for (int i = threadIdx.x; i < 2 * SCR; i += blockDim.x) {
        const int cls = i / SCR, g = i % SCR;
        int v = 0;
        if (g < G) {
            for (int k = 0; k < 8; k++) v |= (count_at(cls, 8 * g + k) > CMIN) << k;
        } else if (g - G < tail_steps) {
            v = count_at(cls, 8 * G + (g - G));
        }
        script[i] = (unsigned char)v;
    }
    if (lane == 0) mbar_init(&bars[warp], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    float mx = 0.f;
    uint32_t parity = 0;
    const unsigned long long warps_total = (unsigned long long)gridDim.x * rp.nwarps;
    const unsigned char *myscript = script + h * SCR;
    float *my_stage = stage + lane * STAGE_STRIDE;
    const int swz = 4 * ((lane >> 3) & 3);

    for (unsigned long long tile = rp.tile0 + (unsigned long long)blockIdx.x * rp.nwarps + warp; tile < rp.tile0 + rp.ntiles;
         tile += warps_total) {
        const unsigned long long out0 = tile * (unsigned long long)(PERIODS * L);          // first output of the warp tile
        const long long gA = (long long)(tile * (unsigned long long)(PERIODS * M)) - 1;    // first input frame needed
        const size_t boff = (size_t)(gA - (long long)a.in_first) * 4;
        const size_t a0 = boff & ~(size_t)15;
        const int sh = (int)((boff - a0) >> 2);
        const uint32_t bytes = (uint32_t)(((size_t)(PERIODS * M + 3 + sh) * 4 + 15) & ~(size_t)15);
        __syncwarp();
        if (lane == 0) {
            fence_async_smem();                       // earlier generic reads of this buffer vs the async write
            mbar_expect_tx(&bars[warp], bytes);
            bulk_load(raw, a.in + a0, bytes, &bars[warp]);
        }
        mbar_wait(&bars[warp], parity);
        parity ^= 1;
        // local frame i of this lane = row[i] = global frame (tile start + p*M + fbase - 1 + i)
        const uint32_t *row = raw + sh + pp * M + fbase;

        // Converted frames live in 8 register slots, frame f in slot f % 8; the loop is unrolled over 8
        // frames so every slot index is static.  Frame f + 3 is loaded and converted while the outputs
        // of frame f are produced (software pipelining: the load -> convert chain is off the critical path).
        f32x2 c0 = cvt_frame(row[0]), c1 = cvt_frame(row[1]), c2 = cvt_frame(row[2]), c3 = cvt_frame(row[3]);
        f32x2 c4 = cvt_frame(row[4]), c5 = cvt_frame(row[5]), c6 = 0, c7 = 0;
        int er = 0;                                   // outputs produced so far by this lane (relative to e_begin)
        const float4 *wp = W + e_begin;
        int nfl = 0;                                  // flushes done in this tile (16 outputs per lane each)
        float *out_tile = nullptr;
        if (APPLY) out_tile = a.out + (size_t)(out0 - a.out_first);

        // one output of this lane's half period: the weights *wp are the same for every lane of the class
        auto emit = [&](f32x2 p0, f32x2 p1, f32x2 p2, f32x2 p3) {
            const float4 w = *wp++;
            f32x2 acc = mul2(p0, pack2(w.x, w.x));
            acc = fma2(p1, pack2(w.y, w.y), acc);
            acc = fma2(p2, pack2(w.z, w.z), acc);
            acc = fma2(p3, pack2(w.w, w.w), acc);
            float vl, vr;
            unpack2(acc, vl, vr);
            vl = clamp_scaled(vl);                        // A:668 in the scaled domain (inputs finite, |v| <= 1)
            vr = clamp_scaled(vr);
            const float sum = vl + vr;                    // (0 + L) + R, A:686; the /2 is in the final scale
            if (APPLY) {
                const float o = sum * mult;
                my_stage[(er & (STAGE_RING - 1)) ^ swz] = CLAMP1 ? fminf(fmaxf(o, one_lo), one_hi) : o;   // A:3455
            } else {
                mx = fmaxf(mx, fabsf(sum));
            }
            er++;
        };
        // uniform point: flush once BOTH classes have 16 outputs staged (they drift by a few outputs at most;
        // the ring holds 32).  Called after every 8-frame group (CMIN == 1: <= 16 new outputs), after every
        // step when a step can produce more, and after every tail step.
        auto maybe_flush = [&]() {
            const int pend = er - STAGE_COLS * nfl;
            const int other = __shfl_xor_sync(0xffffffffu, pend, 16);
            if ((pend < other ? pend : other) >= STAGE_COLS) {
                flush_stage(stage, out_tile, L, STAGE_COLS * nfl, lane);
                nfl++;
            }
        };
        // step for frame F (p3 = frame F): prefetch-convert frame F + 3 into NEXT, then CMIN (+1) outputs
#define AUKIT_RUN_STEP(NEXT, P0, P1, P2, P3, F, EXTRA)                       \
        NEXT = cvt_frame(row[(F) + 3]);                                      \
        _Pragma("unroll")                                                    \
        for (int c = 0; c < CMIN; c++) emit(P0, P1, P2, P3);                 \
        if (EXTRA) emit(P0, P1, P2, P3);                                     \
        if (APPLY && CMIN > 1) maybe_flush();
        int F = 3;
#pragma unroll 1
        for (int g = 0; g < G; g++, F += 8) {
            const int fl = myscript[g];
            AUKIT_RUN_STEP(c6, c0, c1, c2, c3, F, fl & 1)
            AUKIT_RUN_STEP(c7, c1, c2, c3, c4, F + 1, fl & 2)
            AUKIT_RUN_STEP(c0, c2, c3, c4, c5, F + 2, fl & 4)
            AUKIT_RUN_STEP(c1, c3, c4, c5, c6, F + 3, fl & 8)
            AUKIT_RUN_STEP(c2, c4, c5, c6, c7, F + 4, fl & 16)
            AUKIT_RUN_STEP(c3, c5, c6, c7, c0, F + 5, fl & 32)
            AUKIT_RUN_STEP(c4, c6, c7, c0, c1, F + 6, fl & 64)
            AUKIT_RUN_STEP(c5, c7, c0, c1, c2, F + 7, fl & 128)
            if (APPLY && CMIN == 1) maybe_flush();
        }
#undef AUKIT_RUN_STEP
        // tail: the remaining steps (up to 15), with explicit per-class counts (may be 0 for one class)
#define AUKIT_RUN_TAIL(NEXT, P0, P1, P2, P3, K)                              \
        if ((K) < tail_steps) {                                              \
            NEXT = cvt_frame(row[F + (K) + 3]);                              \
            _Pragma("unroll 1")                                              \
            for (int c = myscript[G + (K)]; c > 0; c--) emit(P0, P1, P2, P3); \
            if (APPLY) maybe_flush();                                        \
        }
        AUKIT_RUN_TAIL(c6, c0, c1, c2, c3, 0)
        AUKIT_RUN_TAIL(c7, c1, c2, c3, c4, 1)
        AUKIT_RUN_TAIL(c0, c2, c3, c4, c5, 2)
        AUKIT_RUN_TAIL(c1, c3, c4, c5, c6, 3)
        AUKIT_RUN_TAIL(c2, c4, c5, c6, c7, 4)
        AUKIT_RUN_TAIL(c3, c5, c6, c7, c0, 5)
        AUKIT_RUN_TAIL(c4, c6, c7, c0, c1, 6)
        AUKIT_RUN_TAIL(c5, c7, c0, c1, c2, 7)
        AUKIT_RUN_TAIL(c6, c0, c1, c2, c3, 8)
        AUKIT_RUN_TAIL(c7, c1, c2, c3, c4, 9)
        AUKIT_RUN_TAIL(c0, c2, c3, c4, c5, 10)
        AUKIT_RUN_TAIL(c1, c3, c4, c5, c6, 11)
        AUKIT_RUN_TAIL(c2, c4, c5, c6, c7, 12)
        AUKIT_RUN_TAIL(c3, c5, c6, c7, c0, 13)
        AUKIT_RUN_TAIL(c4, c6, c7, c0, c1, 14)
#undef AUKIT_RUN_TAIL
        if (APPLY) {
            // every lane has now produced its L/2 outputs (a multiple of 16): drain the staging ring
            for (; STAGE_COLS * nfl < LH; nfl++) flush_stage(stage, out_tile, L, STAGE_COLS * nfl, lane);
        }
    }
    if (!APPLY) {
        __shared__ float wm[32];
        mx = warp_max(mx) * (1.0f / 65536.0f);         // back from the scaled domain: 2^-15, and /2 for the mono mean
        if (lane == 0) wm[warp] = mx;
        __syncthreads();
        if (threadIdx.x < 32) {
            mx = threadIdx.x < rp.nwarps ? wm[threadIdx.x] : 0.0f;
            mx = warp_max(mx);
            if (threadIdx.x == 0) atomic_max_nonneg(a.d_max, mx);
        }
    }
}

template <bool APPLY>
int launch_run(aukit_ctx *ctx, const pipe_args &a, run_plan rp) {
    const int L = rp.L, M = rp.M;
    rp.raw_words = ((PERIODS * M + 3 + 3 + 24) + 31) / 32 * 32;  // tile + halo + alignment shift + prefetch slack
    const size_t fixed = (size_t)L * 16 + 2 * 64;
    const size_t per_warp = (size_t)rp.raw_words * 4 + (APPLY ? 32 * STAGE_STRIDE * 4 : 0);
    const size_t budget = 224 * 1024;
    int nw = (int)((budget - fixed - 256 - 256) / per_warp);
    if (nw > 24) nw = 24;
    if (nw < 2) return 0;                                         // not worth it: let the caller fall back
    rp.nwarps = nw;
    const size_t smem = fixed + (((size_t)nw * 8 + 127) & ~(size_t)127) + (size_t)nw * per_warp + 128;
    const int cmin = L / M;
    // the final clamp to +-1 can only act when |peakAmplitude| is (about) 1 or more (negative peaks included: normalize(a, -2))
    const bool clamp1 = APPLY && !(fabs(a.peak) < 1.0 - 9.5367431640625e-07);
    auto kern = cmin == 1 ? run_kernel<APPLY, 1, false> : (cmin == 2 ? run_kernel<APPLY, 2, false> : run_kernel<APPLY, 4, false>);
    if (clamp1) kern = cmin == 1 ? run_kernel<APPLY, 1, true> : (cmin == 2 ? run_kernel<APPLY, 2, true> : run_kernel<APPLY, 4, true>);
    if (aukit_cuda_check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "smem attr")) return -1;
    unsigned long long g = (rp.ntiles + nw - 1) / nw;
    if (g > (unsigned long long)ctx->num_sms) g = ctx->num_sms;
    kern<<<(unsigned)g, nw * 32, smem, ctx->stream>>>(a, rp);
    ctx->launches++;
    return aukit_cuda_check(cudaGetLastError(), "run_kernel launch") ? -1 : 1;
}


// =====================================================================================================
// Static variant for the canonical 44.1 -> 48 kHz ratio (L = 160, M = 147).
//
// What bounds run_kernel above is instruction issue: per input frame a flag test, a divergent
// branch (the two half-period classes share a warp) with its BSSY/BSYNC pair and pointer bumps, and per output
// four FMNMX for the per-channel clamp of A:668 -- and because every output sits in its own basic block the
// compiler reuses the same registers for consecutive outputs, so they cannot overlap.  Here
//   * a warp is ONE class: warps 2k / 2k+1 share a 32-period input tile and take the first / second half of
//     every period (lane = period), so the whole half period is warp-uniform and, with L and M template
//     parameters, a straight line: which frame yields one or two outputs, every weight address, every staging
//     slot and every flush point is a compile-time constant (no script, no branches, no pointer arithmetic), and
//     ptxas interleaves neighbouring outputs;
//   * the clamp of A:668 is checked, not applied: one FMNMX3 per output tracks max(|l|, |r|) of the unclamped
//     channel values; only if a tile saw a value beyond +-1 (cubic overshoot of a near-full-scale signal) is the
//     tile redone by the clamping twin of the same code, and the warp stays on the twin from then on.  Where no
//     clamp acts the two are the same arithmetic, so results stay bit-identical to the polyphase kernel's;
//   * four outputs leave in one STS.128 into a 16-column staging tile whose XOR swizzle makes both that store
//     and the transposed LDS.128 of the flush conflict-free.
constexpr int SPERIODS = 32;          // periods per warp-pair tile (lane = period)
constexpr int SSTAGE_WORDS = 32 * 16; // staging floats per warp: 32 rows x 16 columns

template <int L, int M, int CLS>
struct half_geom {
    static constexpr int LH = L / 2;
    static constexpr int EB = CLS * LH;                                     // first output of the half (within the period)
    static constexpr int FBASE = (int)(((long long)EB * M) / L);            // floor position of that output
    static constexpr int NSTEPS = (int)(((long long)(EB + LH - 1) * M) / L) - FBASE + 1;
    // outputs of this half whose floor position is FBASE + s
    static constexpr int count(int s) {
        const long long f = FBASE + s;
        long long lo = (f * L + M - 1) / M, hi = ((f + 1) * L + M - 1) / M;
        if (lo < EB) lo = EB;
        if (hi > EB + LH) hi = EB + LH;
        return hi > lo ? (int)(hi - lo) : 0;
    }
    static constexpr int before(int s) {
        int n = 0;
        for (int k = 0; k < s; k++) n += count(k);
        return n;
    }
};

// Loop state of a warp lives in SHARED memory on purpose, read through volatile accesses where it is used.  Left in
// registers, ptxas spills exactly these values around the straight-line body (they are only needed between tiles),
// and with ~220 KB of shared memory carved out the L1 behind local memory is a few KB: every reload was an L2 round
// trip per tile (ncu: the instructions after them carried 12 % of all stall samples).
struct __align__(16) warp_state {
    unsigned long long dst;    // APPLY: global address of output 0 of the current tile
    int i;                     // the warp's current tile of the CTA's sequence (pair, pair + npairs, ...)
    uint32_t stage_off;        // byte offset of this warp's staging tile from the dynamic shared memory base
    int cls, slot;             // half-period class; frame buffer holding tile i
    int seg;                   // drift segment of the warp's current tile
};
struct __align__(16) slot_state {       // one frame buffer of the CTA's ring
    unsigned long long full;   // mbarrier: the bulk copy of the buffer's current tile has landed
    int done;                  // warps that finished with the buffer (two per use)
    int issued;                // uses whose bulk copy has been issued (guards the parity wait against a two-phase lead)
};
struct __align__(16) cta_state {
    unsigned long long src0, src_step;   // global address of the CTA's first bulk copy; bytes between consecutive tiles
    unsigned long long dst0, dst_step;   // the same for the outputs (bytes)
    uint32_t bytes;                      // size of one bulk copy
    int sh;                              // frames between the 16-byte aligned copy and the tile's first frame
    int n, nbuf, npairs;                 // tiles of this CTA; frame buffers; warp pairs
    uint32_t bufs_off, buf_bytes;        // ring: offset from the dynamic shared memory base, pitch
    float mult, one_hi;
    int start_twin;                      // the call already knows that the channel clamp acts: skip the checking code
    unsigned long long tile_first, tile_step;   // global index of the CTA's tile 0; tiles between consecutive ones
    int wclaim[2], wtag[2];              // weight-table ring: segment being built into / available in slot (seg & 1)
    int nseg;
};

struct half_ctx {
    const uint32_t *row;       // this lane's first frame (floor position FBASE - 1 of its period)
    const float4 *Wc;          // weights of the half's first output
    float mult, one_lo, one_hi;
    uint32_t stage_x;          // shared address of this lane's staging row, XOR-swizzle of the column quad folded in
    const float *fsrc;         // flush: this lane's 16-byte piece of staging row lane >> 2
    uint32_t loff;             // flush: offset of that piece from the tile's output 0 (floats)
    volatile warp_state *ws;
};

// 16 staged columns of all 32 rows -> global: 4 lanes write 64 contiguous bytes of one period's half.
// Staging traffic (STS.128 in semit, the warp barriers, LDS.128 and the global stores here) is written as volatile asm
// WITHOUT a memory clobber: those statements keep their order among themselves, but the compiler may move the
// weight / frame loads of the following outputs across them (with clobbers every group of four outputs began by
// waiting for its own LDS: short-scoreboard was 23 % of the apply pass's stall samples).
template <int L>
__device__ __forceinline__ void sflush(const half_ctx &hc, int col0) {
    asm volatile("bar.warp.sync 0xffffffff;");
    float *fdst = reinterpret_cast<float *>(hc.ws->dst) + hc.loff;
    const uint32_t fs = smem_u32(hc.fsrc);
    float4 v[4];
#pragma unroll
    for (int it = 0; it < 4; it++)
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[it].x), "=f"(v[it].y), "=f"(v[it].z), "=f"(v[it].w) : "r"(fs + it * 8 * 64));
    asm volatile("bar.warp.sync 0xffffffff;");
#pragma unroll
    for (int it = 0; it < 4; it++)
        asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                     :: "l"(fdst + (size_t)it * 8 * L + col0), "f"(v[it].x), "f"(v[it].y), "f"(v[it].z), "f"(v[it].w));
}

template <bool APPLY, bool CLAMPCH, bool CLAMP1, int L, int E>
__device__ __forceinline__ void semit(const half_ctx &hc, f32x2 p0, f32x2 p1, f32x2 p2, f32x2 p3, float (&o)[4], float &mx, float &chk) {
    const float4 w = hc.Wc[E];
    f32x2 acc = mul2(p0, pack2(w.x, w.x));
    acc = fma2(p1, pack2(w.y, w.y), acc);
    acc = fma2(p2, pack2(w.z, w.z), acc);
    acc = fma2(p3, pack2(w.w, w.w), acc);
    float vl, vr;
    unpack2(acc, vl, vr);
    if (CLAMPCH) {
        vl = clamp_scaled(vl);                            // A:668 in the scaled domain, one FMNMX per channel
        vr = clamp_scaled(vr);
    } else {
        chk = fmaxf(chk, fmaxf(fabsf(vl), fabsf(vr)));    // one FMNMX3: did A:668 have anything to do?
    }
    const float sum = vl + vr;                            // (0 + L) + R, A:686; the /2 is in the final scale
    if (APPLY) {
        const float v = sum * hc.mult;
        o[E & 3] = CLAMP1 ? fminf(fmaxf(v, hc.one_lo), hc.one_hi) : v;   // A:3455
        if ((E & 3) == 3)   // column quad (E >> 2) & 3 of this lane's row, quad index XORed with (lane >> 1) & 3
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(hc.stage_x ^ (uint32_t)(((E >> 2) & 3) << 4)), "f"(o[0]), "f"(o[1]),
                         "f"(o[2]), "f"(o[3]));
        if ((E & 15) == 15) sflush<L>(hc, E - 15);
    } else {
        mx = fmaxf(mx, fabsf(sum));
    }
}

// Returns true when the CHECKING variant gave up early: right after output 15 of the half (the first flush point) the
// warp votes on chk; with a near-full-scale signal the clamp of A:668 acts within the first few outputs of some lane,
// so the tile is handed to the clamping twin after a fifth of its work instead of after all of it.
template <bool APPLY, bool CLAMPCH, bool CLAMP1, bool CVTA, int L, int M, int CLS, int S>
__device__ __forceinline__ bool ssteps(const half_ctx &hc, f32x2 (&c)[8], float (&o)[4], float &mx, float &chk) {
    using G = half_geom<L, M, CLS>;
    if constexpr (S < G::NSTEPS) {
        if constexpr (S + 6 <= G::NSTEPS + 2) c[(S + 6) & 7] = cvt_frame<CVTA>(hc.row[S + 6]);
        constexpr int E0 = G::before(S), N = G::count(S);
        static_assert(N <= 5, "at most CMIN + 1 outputs per input frame");
        if constexpr (N > 0) semit<APPLY, CLAMPCH, CLAMP1, L, E0>(hc, c[S & 7], c[(S + 1) & 7], c[(S + 2) & 7], c[(S + 3) & 7], o, mx, chk);
        if constexpr (N > 1) semit<APPLY, CLAMPCH, CLAMP1, L, E0 + 1>(hc, c[S & 7], c[(S + 1) & 7], c[(S + 2) & 7], c[(S + 3) & 7], o, mx, chk);
        if constexpr (N > 2) semit<APPLY, CLAMPCH, CLAMP1, L, E0 + 2>(hc, c[S & 7], c[(S + 1) & 7], c[(S + 2) & 7], c[(S + 3) & 7], o, mx, chk);
        if constexpr (N > 3) semit<APPLY, CLAMPCH, CLAMP1, L, E0 + 3>(hc, c[S & 7], c[(S + 1) & 7], c[(S + 2) & 7], c[(S + 3) & 7], o, mx, chk);
        if constexpr (N > 4) semit<APPLY, CLAMPCH, CLAMP1, L, E0 + 4>(hc, c[S & 7], c[(S + 1) & 7], c[(S + 2) & 7], c[(S + 3) & 7], o, mx, chk);
        if constexpr (!CLAMPCH && E0 <= 15 && 15 < E0 + N) {
            if (__any_sync(0xffffffffu, chk > 32768.0f)) return true;
        }
        return ssteps<APPLY, CLAMPCH, CLAMP1, CVTA, L, M, CLS, S + 1>(hc, c, o, mx, chk);
    }
    return false;
}

// one half period of one lane, start to end; returns max |l + r| of its outputs, chk = max |channel value|
template <bool APPLY, bool CLAMPCH, bool CLAMP1, bool CVTA, int L, int M, int CLS>
__device__ __forceinline__ float shalf(const half_ctx &hc, float &chk, bool &gave_up) {
    using G = half_geom<L, M, CLS>;
    static_assert(G::before(G::NSTEPS) == L / 2, "every output of the half is produced");
    static_assert((L / 2) % 16 == 0, "whole flush groups");
    f32x2 c[8];
#pragma unroll
    for (int i = 0; i < 6; i++) c[i] = cvt_frame<CVTA>(hc.row[i]);
    c[6] = c[7] = 0;
    float o[4] = {0.f, 0.f, 0.f, 0.f};
    float mx = 0.f;
    gave_up = ssteps<APPLY, CLAMPCH, CLAMP1, CVTA, L, M, CLS, 0>(hc, c, o, mx, chk);
    return mx;
}

__device__ __forceinline__ void mbar_wait32(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra WAIT_%=;\n\t}"
        ::"r"(bar), "r"(parity) : "memory");
}

constexpr int SMAX_BUFS = 12;

template <bool APPLY, bool CLAMP1, bool CVTA, int L, int M>
__global__ void __launch_bounds__(APPLY ? 512 : 576, 1) run_static_kernel(pipe_args a, run_plan rp) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ warp_state wst[24];
    __shared__ slot_state sst[SMAX_BUFS];
    __shared__ cta_state cst;
    // The CTA walks tiles  tile0 + blockIdx.x + i * gridDim.x,  i = 0 .. n-1.  Tile i lives in frame buffer i % nbuf and
    // is consumed by warp pair i % npairs (warp 2p: first halves of its 32 periods, warp 2p+1: second halves).  With
    // nbuf = npairs + 2 (apply) or + 4 (peak) a pair's next tile was requested most of a tile time earlier: whichever of a buffer's two
    // warps finishes LAST issues the bulk copy of tile i + nbuf into it -- no producer warp, no pair barrier, and
    // nobody waits for DRAM unless DRAM is the bottleneck.
    {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, pair = warp >> 1, cls = warp & 1, npairs = rp.nwarps >> 1;
        // Tiles of this CTA.  One drift segment in the launch: tile0 + blockIdx.x + i * gridDim.x (neighbouring CTAs stream
        // neighbouring tiles).  Several segments: a contiguous chunk per CTA, so that only the CTA whose chunk holds a
        // segment boundary needs two weight tables -- every other CTA keeps the frame buffer a second table would cost
        // (the peak pass lost 5 % to that on far time shards).
        // (peak pass only: in the apply pass 146 distant output streams cost more than the buffer gains, 0.224 -> 0.234 ms)
        const bool blocked = rp.nseg > 1 && !APPLY;
        const unsigned long long c0 = blocked ? rp.ntiles * blockIdx.x / gridDim.x : blockIdx.x;
        const unsigned long long c1 = blocked ? rp.ntiles * (blockIdx.x + 1) / gridDim.x : rp.ntiles;
        const unsigned long long cta_first = rp.tile0 + c0, cta_step = blocked ? 1ull : (unsigned long long)gridDim.x;
        const int cta_n = blocked ? (int)(c1 - c0) : (c0 < rp.ntiles ? (int)((rp.ntiles - c0 + gridDim.x - 1) / gridDim.x) : 0);
        int seg0 = 0;
        while (seg0 + 1 < rp.nseg && cta_first >= rp.seg_end[seg0]) seg0++;
        const bool spans = rp.nseg > 1 && cta_n > 0 && seg0 + 1 < rp.nseg &&
                           cta_first + (unsigned long long)(cta_n - 1) * cta_step >= rp.seg_end[seg0];
        // layout: weights[nW][L] float4 (nW = 2 when this CTA's tiles span drift segments) | nbuf frame buffers | staging
        float4 *W = reinterpret_cast<float4 *>(smem);
        const int nW = (APPLY ? rp.nseg > 1 : spans) ? 2 : 1;          // apply pass: strided tiles, every CTA meets every segment
        const int nbuf = (!APPLY && spans) ? rp.nbuf2 : rp.nbuf;
        const uint32_t bufs_off = (uint32_t)(nW * L) * 16, buf_bytes = (uint32_t)rp.raw_words * 4;
        for (int e = threadIdx.x; e < L; e += blockDim.x) {     // fp64 weights of A:265 at fraction j/L + delta, narrowed
            const int j = (int)(((long long)e * M) % L);
            const double x = (double)j / (double)L + (double)rp.seg_delta[seg0];
            const double x2 = x * x, x3 = x2 * x;
            W[(seg0 & (nW - 1)) * L + e] = make_float4((float)(-0.5 * x3 + x2 - 0.5 * x), (float)(1.5 * x3 - 2.5 * x2 + 1.0),
                                                       (float)(-1.5 * x3 + 2.0 * x2 + 0.5 * x), (float)(0.5 * x3 - 0.5 * x2));
        }
        if (lane == 0) {
            warp_state &w = wst[warp];
            w.dst = 0ull; w.i = pair; w.cls = cls; w.slot = 0; w.seg = seg0;
            w.stage_off = bufs_off + (uint32_t)nbuf * buf_bytes + (uint32_t)warp * SSTAGE_WORDS * 4;
        }
        if (threadIdx.x == 0) {
            // the byte offset of a tile inside its 16-byte aligned bulk copy (sh) is the same for every tile: the tile
            // pitch is a multiple of 16 bytes
            const unsigned long long first = cta_first;
            const size_t boff = (size_t)((long long)(first * (unsigned long long)(SPERIODS * M)) - 1 - (long long)a.in_first) * 4;
            const int sh = (int)((boff & 15) >> 2);
            const int n = cta_n;
            const uint32_t bytes = (uint32_t)(((size_t)(SPERIODS * M + 3 + sh) * 4 + 15) & ~(size_t)15);
            cst.src0 = (unsigned long long)(uintptr_t)(a.in + (boff & ~(size_t)15));
            cst.src_step = cta_step * (SPERIODS * M * 4);
            cst.dst0 = APPLY ? (unsigned long long)(uintptr_t)(a.out + (size_t)(first * (unsigned long long)(SPERIODS * L) - a.out_first)) : 0ull;
            cst.dst_step = cta_step * (SPERIODS * L * 4);
            cst.bytes = bytes; cst.sh = sh; cst.n = n; cst.nbuf = nbuf; cst.npairs = npairs;
            cst.bufs_off = bufs_off; cst.buf_bytes = buf_bytes;
            cst.tile_first = first; cst.tile_step = cta_step;
            cst.nseg = nW == 2 ? rp.nseg : 1;                  // one resident table: the segment logic of fetch() is off
            cst.wclaim[0] = cst.wclaim[1] = cst.wtag[0] = cst.wtag[1] = -1;
            cst.wclaim[seg0 & 1] = cst.wtag[seg0 & 1] = seg0;
            float mult = 0.f, one_hi = 1.0f;
            if (APPLY) {                                        // same scale and silence rule as run_kernel
                const float mx0 = a.d_max[0];
                mult = mx0 > 0.f ? (float)(a.peak / (double)mx0) * (1.0f / 65536.0f) : __int_as_float(0x7FC00000);
                if (!(mx0 > 0.f)) one_hi = mult;
            }
            cst.mult = mult; cst.one_hi = one_hi;
            cst.start_twin = (a.hint && *reinterpret_cast<volatile int *>(a.hint) == a.epoch) ? 1 : 0;
            for (int j = 0; j < nbuf; j++) {
                mbar_init(reinterpret_cast<uint64_t *>(&sst[j].full), 1);
                sst[j].done = 0;
                sst[j].issued = j < n ? 1 : 0;
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            for (int j = 0; j < nbuf && j < n; j++) {           // fill the ring
                mbar_expect_tx(reinterpret_cast<uint64_t *>(&sst[j].full), bytes);
                bulk_load(smem + bufs_off + (size_t)j * buf_bytes, reinterpret_cast<const void *>(cst.src0 + (unsigned long long)j * cst.src_step),
                          bytes, reinterpret_cast<uint64_t *>(&sst[j].full));
            }
        }
        __syncthreads();
    }

    volatile cta_state *cs = &cst;
    // (the state pointer is recomputed from a volatile read of %tid wherever it is needed, so that it is not a
    // loop-carried register either)
    auto my_state = [&]() -> volatile warp_state * {
        uint32_t t;
        asm volatile("mov.u32 %0, %%tid.x;" : "=r"(t));
        return &wst[t >> 5];
    };
    half_ctx hc;
    // Wait for the warp's next tile; false when it has none left.  Everything the body needs is rebuilt here from the
    // lane id and the shared state (a handful of integer instructions per 1300-instruction tile): values carried in
    // registers across the body are the ones ptxas spills.
    auto fetch = [&]() -> bool {
        volatile warp_state *ws = my_state();
        const int i = ws->i;
        if (i >= cs->n) return false;
        uint32_t lane;
        asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
        const int cls = ws->cls, nbuf = cs->nbuf;
        const int slot = i % nbuf, use = i / nbuf;
        volatile slot_state *ss = &sst[slot];
        if (lane == 0) {
            ws->slot = slot;
            if (APPLY) ws->dst = cs->dst0 + (unsigned long long)i * cs->dst_step;
        }
        const uint32_t raw_off = cs->bufs_off + (uint32_t)slot * cs->buf_bytes, stage_off = ws->stage_off;
        hc.row = reinterpret_cast<const uint32_t *>(smem + raw_off) + cs->sh + lane * M + (cls ? half_geom<L, M, 1>::FBASE : 0);
        // drift segment of this tile; the first warp to reach a new segment rebuilds the table in slot (seg & 1) -- the
        // other slot still serves warps on the previous segment (a warp is never a whole segment behind: segments are
        // thousands of tiles long, the ring a dozen)
        int wslot = 0;
        if (cs->nseg > 1) {
            const unsigned long long tile = cs->tile_first + (unsigned long long)i * cs->tile_step;
            int seg = ws->seg;
            while (seg + 1 < cs->nseg && tile >= rp.seg_end[seg]) seg++;
            wslot = seg & 1;
            if (seg != ws->seg) {
                int owner = 0;
                if (lane == 0) {
                    ws->seg = seg;
                    owner = atomicMax(const_cast<int *>(&cs->wclaim[wslot]), seg) < seg;
                }
                owner = __shfl_sync(0xffffffffu, owner, 0);
                if (owner) {
                    float4 *Wn = reinterpret_cast<float4 *>(smem) + wslot * L;
                    for (int e = (int)lane; e < L; e += 32) {
                        const int j = (int)(((long long)e * M) % L);
                        const double x = (double)j / (double)L + (double)rp.seg_delta[seg];
                        const double x2 = x * x, x3 = x2 * x;
                        Wn[e] = make_float4((float)(-0.5 * x3 + x2 - 0.5 * x), (float)(1.5 * x3 - 2.5 * x2 + 1.0),
                                            (float)(-1.5 * x3 + 2.0 * x2 + 0.5 * x), (float)(0.5 * x3 - 0.5 * x2));
                    }
                    __syncwarp();
                    __threadfence_block();
                    if (lane == 0) cs->wtag[wslot] = seg;
                } else {
                    while (cs->wtag[wslot] < seg) __nanosleep(64);
                    __threadfence_block();
                }
            }
        }
        hc.Wc = reinterpret_cast<const float4 *>(smem) + wslot * L + cls * (L / 2);
        hc.mult = cs->mult;
        hc.one_hi = cs->one_hi;
        hc.one_lo = -hc.one_hi;
        hc.stage_x = smem_u32(smem + stage_off + lane * 64) ^ (((lane >> 1) & 3u) << 4);
        hc.fsrc = reinterpret_cast<const float *>(smem + stage_off) + (lane >> 2) * 16 + 4 * ((lane & 3) ^ ((lane >> 3) & 3));
        hc.loff = (lane >> 2) * L + cls * (L / 2) + 4 * (lane & 3);
        hc.ws = ws;
        // a parity wait cannot tell phase u from phase u - 2: make sure the copy of THIS use has been issued first
        // (only a warp more than a whole ring ahead of the slowest one ever spins here)
        while (ss->issued <= use) __nanosleep(64);
        mbar_wait32(smem_u32(const_cast<unsigned long long *>(&ss->full)), (uint32_t)use & 1u);
        __syncwarp();                                 // lane 0's ws->dst / ws->slot visible to the warp
        return true;
    };
    // After a tile: hand the buffer back.  The second of its two warps to get here requests tile i + nbuf into it.
    auto release = [&]() {
        volatile warp_state *ws = my_state();
        __syncwarp();                                 // every lane has read its frames
        uint32_t lane;
        asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
        if (lane == 0) {
            const int i = ws->i, slot = ws->slot, nbuf = cs->nbuf;
            ws->i = i + cs->npairs;
            __threadfence_block();
            const int before = atomicAdd(const_cast<int *>(&sst[slot].done), 1);
            if ((before & 1) && i + nbuf < cs->n) {
                __threadfence_block();
                fence_async_smem();                   // both warps' generic reads of the buffer vs the async write
                const uint32_t bar = smem_u32(const_cast<unsigned long long *>(&sst[slot].full)), bytes = cs->bytes;
                sst[slot].issued = (i + nbuf) / nbuf + 1;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(smem + cs->bufs_off + (uint32_t)slot * cs->buf_bytes)),
                               "l"(cs->src0 + (unsigned long long)(i + nbuf) * cs->src_step), "r"(bytes), "r"(bar) : "memory");
            }
        }
        __syncwarp();
    };
    // Phase 1: the checking code.  It ends at the first tile in which the clamp of A:668 acted somewhere ...
    // (not entered at all when an earlier launch of the same call -- the peak pass before this apply pass, or an
    // earlier split of the same pass -- already found that the signal needs the clamp: a.hint / a.epoch)
    float mx = 0.f;
    bool bad = false;                              // true: a fetched tile is waiting for the clamping twin
    if (!cs->start_twin) {
        while (fetch()) {
            float chk = 0.f;
            bool gave_up = false;
            const float mt = hc.ws->cls ? shalf<APPLY, false, CLAMP1, CVTA, L, M, 1>(hc, chk, gave_up) : shalf<APPLY, false, CLAMP1, CVTA, L, M, 0>(hc, chk, gave_up);
            if (gave_up || __any_sync(0xffffffffu, chk > 32768.0f)) {
                bad = true;
                if (a.hint && (threadIdx.x & 31) == 0) atomicMax(a.hint, a.epoch);
                break;
            }
            mx = fmaxf(mx, mt);
            release();
        }
    } else {
        bad = fetch();
    }
    // ... phase 2: that tile again (its frames are still in the buffer) and every later one with the clamping twin.
    // Two loops rather than a call inside one: nothing is live across a call.
    if (bad) {
        do {
            float chk = 0.f;
            bool gave_up = false;
            const float mt = hc.ws->cls ? shalf<APPLY, true, CLAMP1, CVTA, L, M, 1>(hc, chk, gave_up) : shalf<APPLY, true, CLAMP1, CVTA, L, M, 0>(hc, chk, gave_up);
            mx = fmaxf(mx, mt);
            release();
        } while (fetch());
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (!APPLY) {
        __shared__ float wm[32];
        mx = warp_max(mx) * (1.0f / 65536.0f);         // back from the scaled domain: 2^-15, and /2 for the mono mean
        if (lane == 0) wm[warp] = mx;
        __syncthreads();
        if (threadIdx.x < 32) {
            mx = threadIdx.x < rp.nwarps ? wm[threadIdx.x] : 0.0f;
            mx = warp_max(mx);
            if (threadIdx.x == 0) atomic_max_nonneg(a.d_max, mx);
        }
    }
}

template <bool APPLY, int L, int M>
int launch_run_static(aukit_ctx *ctx, const pipe_args &a, run_plan rp) {
    rp.raw_words = ((SPERIODS * M + 3 + 3) + 31) / 32 * 32;       // tile + halo + alignment shift; 128-byte pitch
    // peak pass: one weight table, a CTA whose chunk spans two segments trades a buffer for the second (nbuf2);
    // apply pass: two tables in every CTA of a multi-segment launch
    const size_t fixed = (size_t)((APPLY && rp.nseg > 1) ? 2 : 1) * L * 16 + 128;
    const size_t buf = (size_t)rp.raw_words * 4, stage = APPLY ? 2 * SSTAGE_WORDS * 4 : 0;   // per buffer; per pair
    const size_t budget = 227 * 1024 - 2048;                       // opt-in maximum minus this kernel's static shared memory
    // buffers beyond one per pair = tiles in flight while every pair computes.  Measured: the peak pass (no staging, no
    // output stream) gains 9 % from 8 pairs + 4 over 9 + 2 (0.152 -> 0.139 ms); the apply pass is flat from 8 + 2 to 6 + 4
    // Warp pairs first (they are what issues instructions), spare buffers with what is left: 8 pairs + 2 (apply) or
    // + 4 (peak).
    const int want_spare = APPLY ? 2 : 4, cap = APPLY ? 8 : 9;     // cap = launch bounds
    int np = 0, spare = want_spare;
    for (; spare >= 0; spare--) {
        np = (int)((budget - fixed - (size_t)spare * buf) / (buf + stage));
        if (np > cap) np = cap;
        if (np >= 8 || spare == 0) break;
    }
    if (np < 1) return 0;
    int nbuf = (int)((budget - fixed - (size_t)np * stage) / buf);
    if (nbuf > np + spare) nbuf = np + spare;
    if (nbuf > SMAX_BUFS) nbuf = SMAX_BUFS;
    if (nbuf < np) return 0;
    rp.nwarps = 2 * np;
    rp.nbuf = nbuf;
    rp.nbuf2 = nbuf;
    const size_t smem = fixed + (size_t)nbuf * buf + (size_t)np * stage;
    if (!APPLY && rp.nseg > 1) {
        while (rp.nbuf2 > np && fixed + (size_t)L * 16 + (size_t)rp.nbuf2 * buf + (size_t)np * stage > smem) rp.nbuf2--;
        if (fixed + (size_t)L * 16 + (size_t)rp.nbuf2 * buf + (size_t)np * stage > smem) return 0;
    }
    // the final clamp to +-1 can only act when |peakAmplitude| is (about) 1 or more (negative peaks included: normalize(a, -2))
    const bool clamp1 = APPLY && !(fabs(a.peak) < 1.0 - 9.5367431640625e-07);
    // max(u, 0) of the sample conversion rides on the ALU pipe (CVTA): the FMA pipe is the busier one here
    // (measured 0.352 vs 0.371 ms / step, profiles/r2_noise_cvt{1,0}.json)
    auto kern = clamp1 ? run_static_kernel<APPLY, APPLY, true, L, M> : run_static_kernel<APPLY, false, true, L, M>;
    if (aukit_cuda_check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "smem attr")) return -1;
    // Two SMs stay free for the polyphase edge kernels running beside this one on the side stream (pipeline_poly.cu):
    // this kernel's CTAs take a whole SM's shared memory, so otherwise the edges could only start as it drains.
    const int reserve = 2;
    unsigned long long g = (rp.ntiles + np - 1) / np;
    if (g > (unsigned long long)(ctx->num_sms - reserve)) g = ctx->num_sms - reserve;
    kern<<<(unsigned)g, rp.nwarps * 32, smem, ctx->stream>>>(a, rp);
    ctx->launches++;
    return aukit_cuda_check(cudaGetLastError(), "run_static_kernel launch") ? -1 : 1;
}

}  // namespace

// Tries the run-per-lane kernel on the interior of [a.out_first, a.out_first + a.n_out).  On success
// returns 1 and sets [*done_first, *done_first + *done_count) to the outputs it produced (whole
// warp tiles); the caller covers the rest with the polyphase kernels.  Returns 0 if not applicable.
int aukit_pipeline_run_try(aukit_ctx *ctx, const pipe_args &a, const aukit_pipeline_desc *p, bool apply, long long L,
                           long long M, double eps_r, bool pow2_ratio, unsigned long long *done_first,
                           unsigned long long *done_count) {
    static const bool disabled = getenv("AUKIT_DISABLE_RUN") && getenv("AUKIT_DISABLE_RUN")[0] == '1';
    // AUKIT_RUN_APPLY=0 keeps the apply pass on the polyphase kernel (A/B measurements, profiles/r1_*)
    static const bool no_run_apply = getenv("AUKIT_RUN_APPLY") && getenv("AUKIT_RUN_APPLY")[0] == '0';
    if (disabled || (apply && no_run_apply)) return 0;
    if (p->bitDepth != 16 || p->dataType != AUKIT_SIGNED || p->bigEndian || p->channels != 2) return 0;
    if (p->interpolation != AUKIT_INTERP_CUBIC || !p->mono) return 0;
    if (L % 32 != 0 || L > 1024 || (M & 1) == 0 || M > 2048) return 0;
    if ((L / 2 - 1) * M / L + 1 - 8 * (((L / 2 - 1) * M / L + 1) / 8) + 2 > 15) return 0;   // tail steps must fit the unrolled tail
    if (L / M != 1 && L / M != 2 && L / M != 4) return 0;          // outputs per input frame: CMIN or CMIN + 1
    if (((uintptr_t)a.in & 15) != 0) return 0;
    if (apply && ((((uintptr_t)a.out) & 15) != 0 || (a.out_first & 3) != 0)) return 0;   // 16-byte aligned bulk stores
    // AUKIT_RUN_STATIC=0 keeps the canonical ratio on the scripted kernel (A/B measurements)
    static const bool no_static = getenv("AUKIT_RUN_STATIC") && getenv("AUKIT_RUN_STATIC")[0] == '0';
    const bool use_static = !no_static && L == 160 && M == 147;
    const unsigned long long periods = use_static ? SPERIODS : PERIODS;
    const unsigned long long tile_out = periods * L, tile_in = periods * M;
    // interior warp tiles: fully inside the output range, every tap inside [0, n_total) and inside the shard window
    unsigned long long t_lo = (a.out_first + tile_out - 1) / tile_out;
    if (t_lo == 0) t_lo = 1;                                       // tile 0 needs frame -1 (clamped): poly path
    unsigned long long t_hi = (a.out_first + a.n_out) / tile_out;  // exclusive
    const unsigned long long in_lo = a.in_first, in_hi = a.in_first + a.in_avail;
    while (t_lo < t_hi && t_lo * tile_in < in_lo + 1 + 4) t_lo++;  // frame t*tile_in - 1 (and the 16-byte round-down) must exist
    while (t_hi > t_lo && ((t_hi - 1) * tile_in + tile_in + 2 + 8 > in_hi || (t_hi - 1) * tile_in + tile_in + 2 >= a.n_total)) t_hi--;
    if (t_hi <= t_lo || t_hi - t_lo < 16) return 0;
    if (!pow2_ratio && (double)(t_hi * tile_in) >= 1518500249.0) return 0;              // beyond 2^30.5: exact-position kernels
    run_plan rp{};
    rp.L = (int)L; rp.M = (int)M;
    // The drift x*eps_r (<= 2^-22.5 frames) is baked into the weight table, constant over a SEGMENT of tiles.  The
    // segments form a fixed GLOBAL grid -- [0, b0) with no drift (below 2^28 / 1.25 frames it is under 2^-25), then
    // [b_k, b_k+1) with b_k+1 = 1.25 b_k and the drift of the segment's middle (within 11 % of the true value: an error
    // below 2^-25.5 in the position, < 6e-8 in the output) -- so what an output frame gets does not depend on how the
    // buffer was sharded: any two shardings give the same bits.
    const unsigned long long b0 = (unsigned long long)(268435456.0 / 1.25 / (double)tile_in);
    unsigned long long t = t_lo;
    int rc = 1;
    rp.nseg = 0;
    unsigned long long batch_first = t_lo;
    unsigned long long seg_lo = 0, seg_hi = pow2_ratio ? ~0ull : b0;     // grid cell containing t
    while (t >= seg_hi) { seg_lo = seg_hi; seg_hi = seg_hi + seg_hi / 4 + 1; }
    while (t < t_hi && rc == 1) {
        const unsigned long long t_end = seg_hi < t_hi ? seg_hi : t_hi;
        const float delta = (pow2_ratio || seg_lo == 0) ? 0.f : (float)(0.5 * ((double)(seg_lo * tile_in) + (double)(seg_hi * tile_in)) * eps_r);
        if (use_static) {
            // the straight-line kernel takes the whole range in ONE launch: segments of constant drift, tables rebuilt in place
            rp.seg_end[rp.nseg] = t_end;
            rp.seg_delta[rp.nseg] = delta;
            rp.nseg++;
            if (rp.nseg == RUN_MAXSEG || t_end == t_hi) {
                rp.tile0 = batch_first;
                rp.ntiles = t_end - batch_first;
                rp.delta = rp.seg_delta[0];
                rc = apply ? launch_run_static<true, 160, 147>(ctx, a, rp) : launch_run_static<false, 160, 147>(ctx, a, rp);
                rp.nseg = 0;
                batch_first = t_end;
            }
        } else {
            rp.tile0 = t;
            rp.ntiles = t_end - t;
            rp.delta = delta;
            rp.nseg = 0;
            rc = apply ? launch_run<true>(ctx, a, rp) : launch_run<false>(ctx, a, rp);
        }
        t = t_end;
        if (t >= seg_hi) { seg_lo = seg_hi; seg_hi = seg_hi + seg_hi / 4 + 1; }
    }
    if (rc != 1) return rc;
    *done_first = t_lo * tile_out;
    *done_count = (t_hi - t_lo) * tile_out;
    return 1;
}
