// lowpass.cu -- K11 effects.lowpass (A:3586-3598): in-place one-pole IIR per channel,
//     a = 1 - exp(-(frequency / sampleRate) * 2 * pi);   d[i] = d[i-1] + a * (d[i] - d[i-1])   for i = 2..#d
// (d[1] is left as it is).  auplay.lua:30 calls it right after effects.normalize: SURVEY 8(f) rank 1.
//
// The recurrence y[i] = (1-a) y[i-1] + a x[i] is linear, so a tile's effect on the state is the pair
// (P, S) = ((1-a)^len, end state from a zero start) and tiles combine associatively.  One pass over
// HBM (4 B read + 4 B written per sample), single kernel, chained with a decoupled look-back:
//   * a CTA claims tiles of 8192 samples in order (atomic ticket => every predecessor is running or done);
//   * the tile is staged through shared memory (coalesced 16-byte accesses both ways); thread t owns 32
//     consecutive samples and runs the reference's own step on them, first from a zero state (-> S_t);
//   * S_t are combined by a warp-shuffle scan with the constant ratio (1-a)^32, then across the 8 warps;
//   * warp 0 publishes the tile aggregate, looks back over the predecessors' aggregates / inclusive states
//     (32 at a time, one 16-byte {value, flag} load each, stopping early once (1-a)^k has decayed below
//     2^-80) and publishes the inclusive state;
//   * the next tile's loads are issued before any of this, so they are in flight meanwhile;
//   * every thread re-runs its 32 steps from its true incoming state and the tile is written back.
// Arithmetic is fp64 like the reference's (Lua numbers): in fp32 the rounding error of a low cut-off is
// amplified by 1/a and would leave the 2^-20 tolerance; the kernel stays HBM-bound either way (6 fp64 ops
// per sample).  Samples are narrowed to f32 only when stored.
//
// Large buffers with an ordinary cut-off take the BLOCKED variant of the same kernel instead of the look-back: when
// (1-a)^8192 < 2^-80 (any cut-off above ~50 Hz at 48 kHz) the state entering a tile is, to fp64 precision, a function of
// the previous tile alone -- the very truncation the look-back applies.  Each CTA then owns one contiguous chunk of tiles
// and carries the state from tile to tile in a register; the state entering a chunk is the zero-start aggregate of the
// tile before it, computed by a small pre-pass (one tile per chunk, read before anything is overwritten).  No tickets, no
// slots to clear, no polling: the CTAs never talk to each other.  Lower cut-offs (down to ~10 Hz) work the same way with a
// pre-pass over the 2 - 8 tiles whose combined weight is below 2^-80, while the chunks are long enough to amortise it.
#include "common.cuh"

#include <math.h>

namespace {

constexpr int LP_THREADS = 256;
constexpr int LP_PER = 32;                       // consecutive samples per thread
constexpr int LP_TILE = LP_THREADS * LP_PER;     // 8192
constexpr int LP_ROW = LP_PER + 4;               // padded row: conflict-free 16-byte accesses both ways
constexpr size_t LP_SMEM = 2 * (size_t)LP_THREADS * LP_ROW * sizeof(float);
constexpr int LP_CTAS_PER_SM = 3;                // resident CTAs: tiles in flight hide the look-back's L2 round trip

// Per (channel, tile): ONE 16-byte slot that holds either {aggregate, 1} or {inclusive state, 2}; it is
// written and read with single 128-bit accesses, so a look-back costs one L2 round trip per window.
struct __align__(16) lp_slot { double v; long long flag; };

__device__ __forceinline__ lp_slot ld_slot(const lp_slot *p) {
    lp_slot r;
    // one .b128 access: single-copy atomic, so the value and its flag can never be seen torn
    asm volatile("{\n\t.reg .b128 q;\n\tld.relaxed.gpu.global.b128 q, [%2];\n\tmov.b128 {%0, %1}, q;\n\t}"
                 : "=l"(*reinterpret_cast<long long *>(&r.v)), "=l"(r.flag) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void st_slot(lp_slot *p, double v, long long flag) {
    asm volatile("{\n\t.reg .b128 q;\n\tmov.b128 q, {%1, %2};\n\tst.relaxed.gpu.global.b128 [%0], q;\n\t}"
                 ::"l"(p), "l"(__double_as_longlong(v)), "l"(flag) : "memory");
}
__device__ __forceinline__ double shfl_up_d(double v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }

// Asynchronous global -> shared copy of one tile into its padded layout (16 bytes per cp.async, no
// registers held while the loads are in flight); ragged tails are filled with plain loads and zeros.
__device__ __forceinline__ void lp_fetch_tile(float *tile, const float *base, int cnt, int t) {
#pragma unroll
    for (int k = 0; k < LP_PER / 4; k++) {
        const int i4 = (k * LP_THREADS + t) * 4;                       // first sample of this float4
        float *dst = &tile[(i4 / LP_PER) * LP_ROW + (i4 % LP_PER)];
        if (i4 + 3 < cnt) {
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(base + i4) : "memory");
        } else {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i4 < cnt) v.x = base[i4];
            if (i4 + 1 < cnt) v.y = base[i4 + 1];
            if (i4 + 2 < cnt) v.z = base[i4 + 2];
            *reinterpret_cast<float4 *>(dst) = v;
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// x^e for a small non-negative integer e by binary exponentiation: a handful of DMULs where pow() is a ~250-instruction
// call (six of them per thread made the prologue a few per cent of the kernel).  The few-ulp difference from pow() is
// far below the f32 rounding of the stored samples.
__device__ __forceinline__ double lp_ipow(double x, unsigned e) {
    double r = 1.0;
    while (e) {
        if (e & 1u) r *= x;
        x *= x;
        e >>= 1;
    }
    return r;
}

// one step of the reference's loop (A:3593-3594 / A:3613-3615); `first` = the channel's first sample, which both
// effects leave as it is
template <bool HIGH>
__device__ __forceinline__ double lp_step(double y, float x, double &xp, bool first, double a) {
    if (!HIGH) return y + a * ((double)x - y);
    const double xd = (double)x;
    const double r = first ? xd : a * ((y + xd) - xp);
    xp = xd;
    return r;
}

// Zero-start run of thread t's LP_PER samples, then the block-wide combine with the constant ratio pt per thread.
// Returns the state entering thread t when nothing enters the tile; warp_tot[] holds the warps' inclusive totals
// (valid after the barrier inside).  Samples past the end of the channel are zeros: they only decay the state, which
// nothing reads afterwards.
template <bool HIGH>
__device__ __forceinline__ double lp_zero_scan(const float *tile, int t, double a, double xprev0, bool chan_first, double pt,
                                               double p_warp, const double *pt_pow, double *warp_tot) {
    const int lane = t & 31, warp = t >> 5;
    double s = 0.0, xp = xprev0;
#pragma unroll
    for (int k = 0; k < LP_PER / 4; k++) {
        const float4 v = *reinterpret_cast<const float4 *>(&tile[t * LP_ROW + 4 * k]);
        s = lp_step<HIGH>(s, v.x, xp, chan_first && k == 0, a);
        s = lp_step<HIGH>(s, v.y, xp, false, a);
        s = lp_step<HIGH>(s, v.z, xp, false, a);
        s = lp_step<HIGH>(s, v.w, xp, false, a);
    }
    double inc = s;
    double r = pt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double up = shfl_up_d(inc, d);
        if (lane >= d) inc = fma(r, up, inc);
        r *= r;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    double wprev = 0.0;                                            // state entering this warp (zero tile carry)
    for (int w = 0; w < warp; w++) wprev = fma(p_warp, wprev, warp_tot[w]);
    double exc = shfl_up_d(inc, 1);                                // inclusive prefix of thread t-1
    if (lane == 0) exc = 0.0;
    return fma(pt_pow[lane], wprev, exc);
}

__device__ __forceinline__ double lp_tile_aggregate(const double *warp_tot, double p_warp) {
    double agg = 0.0;
    for (int w = 0; w < LP_THREADS / 32; w++) agg = fma(p_warp, agg, warp_tot[w]);
    return agg;
}

// What enters chunk q's first tile: the state and, for the high-pass, the last INPUT sample of the tile before it.
struct __align__(16) lp_chunk_in { double state, xlast; };

// Pre-pass of the blocked variant: fills chunk_in[q] (zeros where a chunk starts a channel: the kernel applies the
// channel-start rule itself) and resets the per-channel poison marks.  The state entering a chunk is that of a run from
// zero over the `warm` tiles before it -- `warm` chosen by the host so that ratio^(8192 warm) < 2^-80: whatever came
// earlier weighs less than an fp64 ulp -- or from the channel's first tile if that is nearer, which is exact.  Reads the
// untouched input; the tiles before a chunk are always full ones of the same channel.
template <bool HIGH>
__global__ void __launch_bounds__(LP_THREADS)
lp_chunk_states(const float *__restrict__ data, size_t stride, int channels, double a, double b, unsigned long long tiles_per_ch,
                unsigned long long total, unsigned long long chunk, int warm, lp_chunk_in *__restrict__ chunk_in,
                unsigned long long *__restrict__ poison) {
    __shared__ __align__(16) float tile[LP_THREADS * LP_ROW];
    __shared__ double pt_pow[LP_THREADS];
    __shared__ double warp_tot[LP_THREADS / 32];
    const int t = threadIdx.x;
    if (blockIdx.x == 0)
        for (int c = t; c < channels; c += LP_THREADS) poison[c] = ~0ull;
    const unsigned long long first = (unsigned long long)blockIdx.x * chunk;
    const unsigned long long tl0 = first % tiles_per_ch, ch = first / tiles_per_ch;
    if (first >= total || tl0 == 0) {
        if (t == 0) chunk_in[blockIdx.x] = lp_chunk_in{0.0, 0.0};
        return;
    }
    const double pt = lp_ipow(b, LP_PER);
    pt_pow[t] = lp_ipow(pt, (unsigned)t);
    const double p_warp = lp_ipow(pt, 32), p_tile = lp_ipow(pt, LP_THREADS);
    double state = 0.0;                                                 // thread 0
    for (unsigned long long tl = tl0 < (unsigned long long)warm ? 0 : tl0 - (unsigned long long)warm; tl < tl0; tl++) {
        const float *base = data + (size_t)ch * stride + (size_t)tl * LP_TILE;
        __syncthreads();                                                // the previous tile (and warp_tot) has been read
        lp_fetch_tile(tile, base, LP_TILE, t);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        double xprev0 = 0.0;
        if (HIGH) xprev0 = t > 0 ? (double)tile[(t - 1) * LP_ROW + LP_PER - 1] : (tl > 0 ? (double)base[-1] : 0.0);
        lp_zero_scan<HIGH>(tile, t, a, xprev0, HIGH && tl == 0 && t == 0, pt, p_warp, pt_pow, warp_tot);
        if (t == 0) {
            const double carry = tl == 0 ? (HIGH ? 0.0 : (double)tile[0]) : state;     // A:3591, as in the kernel below
            state = fma(p_tile, carry, lp_tile_aggregate(warp_tot, p_warp));
        }
    }
    if (t == 0) chunk_in[blockIdx.x] = lp_chunk_in{state, (double)tile[(LP_THREADS - 1) * LP_ROW + LP_PER - 1]};
}

// HIGH = false: effects.lowpass, y = y + a (x - y), per-step ratio b = 1 - a.
// HIGH = true:  effects.highpass (A:3605-3618), y = a ((y + x) - x_prev), per-step ratio b = a, y[1] = x[1]; the
//               previous INPUT sample across a tile boundary comes from `xb` (saved before anything is overwritten).
// BLOCKED:      CTA q owns tiles [q * chunk, (q + 1) * chunk) and carries the state itself (chunk_in[q] enters its
//               first tile; the high-pass keeps each tile's last input sample for the next one in shared memory, no
//               `xb`); otherwise tiles are claimed by ticket and chained by the look-back.
template <bool HIGH, bool BLOCKED>
__global__ void __launch_bounds__(LP_THREADS, LP_CTAS_PER_SM)
lowpass_kernel(float *__restrict__ data, size_t stride, int channels, size_t n, double a, double b,
               lp_slot *slots, unsigned long long *ticket, unsigned long long tiles_per_ch, const float *__restrict__ xb,
               unsigned long long *poison, const lp_chunk_in *__restrict__ chunk_in, unsigned long long chunk) {
    extern __shared__ __align__(16) float lp_dyn[];                     // two tile buffers (double buffered: see the loop)
    float *const tiles[2] = {lp_dyn, lp_dyn + LP_THREADS * LP_ROW};
    __shared__ double pt_pow[LP_THREADS];        // (ratio^LP_PER)^t
    __shared__ double warp_tot[LP_THREADS / 32];
    __shared__ double s_carry;
    __shared__ double s_xlast[2];                // BLOCKED high-pass: last input sample of the previous tile (slot it & 1)
    __shared__ unsigned long long s_ticket[2];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const double pt = lp_ipow(b, LP_PER);
    pt_pow[t] = lp_ipow(pt, (unsigned)t);
    const double p_warp = lp_ipow(pt, 32), p_tile = lp_ipow(pt, LP_THREADS);
    const double pl = lp_ipow(p_tile, (unsigned)lane); // look-back weight of the lane-th predecessor
    const double p_tile32 = lp_ipow(p_tile, 32);
    const unsigned long long total = tiles_per_ch * (unsigned long long)channels;
    auto tile_base = [&](unsigned long long id, int &cnt) -> float * {
        const unsigned long long ch = id / tiles_per_ch, tl = id % tiles_per_ch;
        const size_t left = n - (size_t)tl * LP_TILE;
        cnt = left < (size_t)LP_TILE ? (int)left : LP_TILE;
        return data + (size_t)ch * stride + (size_t)tl * LP_TILE;
    };
    // prologue: the first tile (BLOCKED: of this CTA's chunk; otherwise claimed by ticket) and its loads
    unsigned long long id, end_id = total;
    double state = 0.0;                              // BLOCKED, thread 0 only: state entering the current tile
    if (BLOCKED) {
        id = (unsigned long long)blockIdx.x * chunk;
        if (id + chunk < total) end_id = id + chunk;
        if (t == 0 && id < total) {
            const lp_chunk_in ci = chunk_in[blockIdx.x];
            state = ci.state;
            s_xlast[0] = ci.xlast;               // read by this thread only before the next write to the slot
        }
    } else {
        if (t == 0) s_ticket[0] = atomicAdd(ticket, 1ull);
        __syncthreads();
        id = s_ticket[0];
    }
    int cnt = 0;
    float *base = nullptr;
    if (id < end_id) { base = tile_base(id, cnt); lp_fetch_tile(tiles[0], base, cnt, t); }
    for (int it = 0; id < end_id; it++) {
        float *tile = tiles[it & 1];
        // ---- this tile was fetched during the previous iteration; claim the next one
        if (!BLOCKED && t == 0) s_ticket[(it + 1) & 1] = atomicAdd(ticket, 1ull);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        // ---- next tile's loads are in flight while this one is processed (the other buffer was last read
        // before the barrier above)
        const unsigned long long next_id = BLOCKED ? id + 1 : s_ticket[(it + 1) & 1];
        int next_cnt = 0;
        float *next_base = nullptr;
        if (next_id < end_id) { next_base = tile_base(next_id, next_cnt); lp_fetch_tile(tiles[(it + 1) & 1], next_base, next_cnt, t); }
        const int ch = (int)(id / tiles_per_ch);
        const unsigned long long tl = id % tiles_per_ch;
        // input sample just before this thread's first one (highpass only); read now, the rows are overwritten later
        double xprev0 = 0.0;
        if (HIGH) xprev0 = t > 0 ? (double)tile[(t - 1) * LP_ROW + LP_PER - 1] : (tl > 0 ? (BLOCKED ? s_xlast[it & 1] : (double)xb[id]) : 0.0);
        if (HIGH && BLOCKED && t == LP_THREADS - 1) s_xlast[(it + 1) & 1] = (double)tile[t * LP_ROW + LP_PER - 1];
        const bool chan_first = HIGH && tl == 0 && t == 0;
        // ---- zero-start run of this thread's samples + inclusive scan over the block
        const double enter0 = lp_zero_scan<HIGH>(tile, t, a, xprev0, chan_first, pt, p_warp, pt_pow, warp_tot);
        // ---- tile aggregate + the state entering the tile (warp 0)
        if (warp == 0) {
            const double agg = lp_tile_aggregate(warp_tot, p_warp);
            double carry;
            if (tl == 0) {
                // A:3591: d[1] is untouched, which is what a state equal to d[1] gives (l + a*(l - l) = l);
                // highpass handles its first sample inside lp_step(), so nothing enters the first tile
                carry = HIGH ? 0.0 : (double)tile[0];
            } else if (BLOCKED) {
                carry = state;
            } else {
                if (lane == 0) st_slot(&slots[id], agg, 1);
                double acc = 0.0, scale = 1.0;
                long long back = (long long)tl - 1;                    // nearest predecessor tile of this channel
                for (;;) {
                    const long long j = back - lane;
                    lp_slot sv;
                    sv.v = 0.0; sv.flag = 2;
                    // only poll predecessors whose weight can still matter: with (1-a)^8192 tiny (any cut-off above
                    // a few Hz) that is the nearest one or two, and a tile then waits for those alone instead of
                    // for the slowest of 32 (which locks all CTAs into step; measured 2.4x slower)
                    const bool live = j >= 0 && pl * scale >= 8.3e-25;
                    if (live) {
                        const lp_slot *sp = &slots[(unsigned long long)ch * tiles_per_ch + (unsigned long long)j];
                        do { sv = ld_slot(sp); } while (sv.flag == 0);
                    }
                    // first lane holding an inclusive state (tile 0 always does), or the first lane before tile 0
                    const unsigned stop_mask = __ballot_sync(0xffffffffu, sv.flag == 2);
                    const int stop = __ffs(stop_mask) - 1;             // -1: only aggregates in this window
                    double term = (live && (stop < 0 || lane <= stop)) ? pl * sv.v : 0.0;
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) term += __shfl_xor_sync(0xffffffffu, term, d);
                    acc = fma(scale, term, acc);
                    if (stop >= 0) break;
                    scale *= p_tile32;
                    back -= 32;
                    if (scale < 8.3e-25) break;                        // 2^-80: older tiles no longer matter
                }
                carry = acc;
            }
            if (lane == 0) {
                const double incl = fma(p_tile, carry, agg);
                if (BLOCKED) state = incl;
                else st_slot(&slots[id], incl, 2);
                // a non-finite state never leaves the reference's recurrence (A:3592-3595); the bounded look-back
                // above (and the one-tile memory of a chunk start) can miss it, so the first such tile of a channel is
                // recorded for lp_poison_fix, which reads the state from the tile's slot
                if (!isfinite(incl)) {
                    atomicMin(&poison[ch], tl);
                    // flag 3: the tile's last input sample is not finite either (still unmodified here); lp_poison_fix
                    // needs to know for the high-pass, and the blocked variant has no `xb` to look it up in
                    if (BLOCKED) st_slot(&slots[id], incl, (HIGH && !isfinite(tile[(LP_THREADS - 1) * LP_ROW + LP_PER - 1])) ? 3 : 2);
                }
                s_carry = carry;
            }
        }
        __syncthreads();
        // ---- true run: state entering thread t = pt^t * (tile carry) + (state entering t with a zero tile carry)
        double y = fma(pt_pow[t], s_carry, enter0);
        double xp = xprev0;
#pragma unroll
        for (int k = 0; k < LP_PER / 4; k++) {
            float4 v = *reinterpret_cast<const float4 *>(&tile[t * LP_ROW + 4 * k]);
            y = lp_step<HIGH>(y, v.x, xp, chan_first && k == 0, a); v.x = (float)y;
            y = lp_step<HIGH>(y, v.y, xp, false, a); v.y = (float)y;
            y = lp_step<HIGH>(y, v.z, xp, false, a); v.z = (float)y;
            y = lp_step<HIGH>(y, v.w, xp, false, a); v.w = (float)y;
            *reinterpret_cast<float4 *>(&tile[t * LP_ROW + 4 * k]) = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < LP_PER / 4; k++) {
            const int i4 = (k * LP_THREADS + t) * 4;
            const float4 v = *reinterpret_cast<const float4 *>(&tile[(i4 / LP_PER) * LP_ROW + (i4 % LP_PER)]);
            if (i4 + 3 < cnt) stg_stream(reinterpret_cast<float4 *>(base + i4), v);
            else {
                if (i4 < cnt) base[i4] = v.x;
                if (i4 + 1 < cnt) base[i4 + 1] = v.y;
                if (i4 + 2 < cnt) base[i4 + 2] = v.z;
            }
        }
        id = next_id; base = next_base; cnt = next_cnt;
    }
}

}  // namespace

// x[tile * LP_TILE - 1] of every tile after the first, per channel: the highpass step needs the previous INPUT sample
// and by the time a tile runs its predecessor may already have been overwritten in place
__global__ void lp_save_boundaries(const float *__restrict__ data, size_t stride, unsigned long long tiles_per_ch,
                                   unsigned long long total, float *__restrict__ xb) {
    const unsigned long long id = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= total) return;
    const unsigned long long ch = id / tiles_per_ch, tl = id % tiles_per_ch;
    xb[id] = tl ? data[(size_t)ch * stride + (size_t)tl * LP_TILE - 1] : 0.0f;
}

// Non-finite input (float PCM can carry NaN / Inf): once the reference's state is non-finite every later sample of the
// channel is too -- NaN for the low-pass (Inf + a (x - Inf) = NaN one step later), NaN or the same Inf for the
// high-pass (a ((Inf + x) - x') = Inf).  Tiles further than the look-back's 2^-80 horizon behind such a tile never
// poll it, so after the scan every tile past the first poisoned one of its channel is overwritten.  No-op (one
// 8-byte read per CTA) for finite audio.
template <bool HIGH>
__global__ void lp_poison_fix(float *__restrict__ data, size_t stride, int channels, size_t n, const lp_slot *slots,
                              unsigned long long tiles_per_ch, const unsigned long long *poison, const float *__restrict__ xb) {
    // xb == nullptr: blocked variant, the slot's flag says whether tile T's last input sample was finite
    for (int ch = 0; ch < channels; ch++) {
        const unsigned long long T = poison[ch];
        if (T >= tiles_per_ch || T + 1 >= tiles_per_ch) continue;          // none (all ones), or nothing after it
        const lp_slot sl = slots[(unsigned long long)ch * tiles_per_ch + T];
        const double st = sl.v;
        // an infinite high-pass state stays infinite only while the inputs are finite: if it came from an infinite LAST
        // input sample of tile T, the next step is a ((Inf + x) - Inf) = NaN
        const bool keep_inf = HIGH && isinf(st) && (xb ? isfinite(xb[(unsigned long long)ch * tiles_per_ch + T + 1]) : sl.flag == 2);
        const float fill = keep_inf ? (float)st : __int_as_float(0x7FC00000);
        float *row = data + (size_t)ch * stride;
        for (size_t i = (size_t)(T + 1) * LP_TILE + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
             i += (size_t)gridDim.x * blockDim.x)
            row[i] = fill;
    }
}

template <bool HIGH, bool BLOCKED>
static void lp_launch(aukit_ctx *ctx, unsigned grid, float *d, size_t stride, int channels, size_t n, double a, double ratio,
                      lp_slot *slots, unsigned long long *ticket, unsigned long long tiles, const float *xb,
                      unsigned long long *poison, const lp_chunk_in *chunk_in, unsigned long long chunk) {
    cudaFuncSetAttribute(lowpass_kernel<HIGH, BLOCKED>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(lowpass_kernel<HIGH, BLOCKED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LP_SMEM);
    lowpass_kernel<HIGH, BLOCKED><<<grid, LP_THREADS, LP_SMEM, ctx->stream>>>(d, stride, channels, n, a, ratio, slots, ticket, tiles, xb,
                                                                              poison, chunk_in, chunk);
    lp_poison_fix<HIGH><<<ctx->num_sms, 256, 0, ctx->stream>>>(d, stride, channels, n, slots, tiles, poison, xb);
    ctx->launches += 2;
}

static int lp_run(aukit_ctx *ctx, float *d, size_t stride, int channels, size_t n, double a, double ratio, bool high) {
    const unsigned long long tiles = (n + LP_TILE - 1) / LP_TILE, total = tiles * (unsigned long long)channels;
    const unsigned long long cap = (unsigned long long)ctx->num_sms * LP_CTAS_PER_SM;
    // blocked chunks (see the top of the file).  `warm` = tiles of memory: the fewest whose combined weight
    // |ratio|^(8192 warm) is below 2^-80, the look-back's own cut (1 for any cut-off above ~50 Hz at 48 kHz, 2 from ~26 Hz,
    // 5 from ~10 Hz).  Each chunk's pre-pass reads that many extra tiles, so the chunks must be long enough for that to cost
    // less than the look-back does (measured with warm = 1: chunks of 4 - 5 tiles are 4 % SLOWER than the look-back, 44 tiles
    // 18 % faster): at least 8 tiles, and four times the warm-up.  NaN or |ratio| >= 1 never qualifies.
    int warm = 0;
    for (int w = 1; w <= 8 && !warm; w++)
        if (pow(fabs(ratio), (double)LP_TILE * w) < 8.3e-25) warm = w;
    const bool blocked = warm > 0 && total >= (unsigned long long)(warm > 2 ? 4 * warm : 8) * cap;
    const unsigned long long chunk = blocked ? (total + cap - 1) / cap : 0;
    const unsigned long long nchunks = blocked ? (total + chunk - 1) / chunk : 0;
    // scratch: one 16-byte slot per (channel, tile) + the ticket counter (+ the boundary samples for highpass, + the
    // states entering the chunks)
    void *scratch = nullptr;
    const size_t slot_bytes = (size_t)total * sizeof(lp_slot);
    const size_t xb_bytes = (high && !blocked) ? (((size_t)total * sizeof(float) + 15) & ~(size_t)15) : 0;
    const size_t poison_bytes = (((size_t)channels * sizeof(unsigned long long)) + 15) & ~(size_t)15;
    const size_t chunk_bytes = (size_t)nchunks * sizeof(lp_chunk_in);
    if (aukit_dev_alloc(ctx, slot_bytes + 16 + poison_bytes + xb_bytes + chunk_bytes, &scratch)) return -1;
    lp_slot *slots = static_cast<lp_slot *>(scratch);
    unsigned long long *ticket = reinterpret_cast<unsigned long long *>(static_cast<char *>(scratch) + slot_bytes);
    unsigned long long *poison = ticket + 2;
    float *xb = xb_bytes ? reinterpret_cast<float *>(static_cast<char *>(scratch) + slot_bytes + 16 + poison_bytes) : nullptr;
    lp_chunk_in *chunk_in = reinterpret_cast<lp_chunk_in *>(static_cast<char *>(scratch) + slot_bytes + 16 + poison_bytes + xb_bytes);
    // the slots are flags only for the look-back; the blocked variant writes (and lp_poison_fix reads) a slot only where
    // the state went non-finite, and its pre-pass also resets the poison marks
    int rc = blocked ? 0 : aukit_cuda_check(cudaMemsetAsync(scratch, 0, slot_bytes + 16, ctx->stream), "memset");
    if (!rc && !blocked) rc = aukit_cuda_check(cudaMemsetAsync(poison, 0xFF, poison_bytes, ctx->stream), "memset");
    if (!rc && xb) {
        lp_save_boundaries<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(d, stride, tiles, total, xb);
        ctx->launches++;
    }
    if (!rc && blocked) {
        if (high) lp_chunk_states<true><<<(unsigned)nchunks, LP_THREADS, 0, ctx->stream>>>(d, stride, channels, a, ratio, tiles, total, chunk, warm, chunk_in, poison);
        else lp_chunk_states<false><<<(unsigned)nchunks, LP_THREADS, 0, ctx->stream>>>(d, stride, channels, a, ratio, tiles, total, chunk, warm, chunk_in, poison);
        ctx->launches++;
    }
    if (!rc) {
        const unsigned g = (unsigned)(blocked ? nchunks : (total < cap ? total : cap));
        if (high) {
            if (blocked) lp_launch<true, true>(ctx, g, d, stride, channels, n, a, ratio, slots, ticket, tiles, xb, poison, chunk_in, chunk);
            else lp_launch<true, false>(ctx, g, d, stride, channels, n, a, ratio, slots, ticket, tiles, xb, poison, chunk_in, chunk);
        } else {
            if (blocked) lp_launch<false, true>(ctx, g, d, stride, channels, n, a, ratio, slots, ticket, tiles, xb, poison, chunk_in, chunk);
            else lp_launch<false, false>(ctx, g, d, stride, channels, n, a, ratio, slots, ticket, tiles, xb, poison, chunk_in, chunk);
        }
        rc = aukit_cuda_check(cudaGetLastError(), "lowpass_kernel launch");
    }
    aukit_dev_free(ctx, scratch);
    return rc;
}

extern "C" int aukit_cuda_dev_lowpass(aukit_ctx *ctx, float *d, size_t stride, int channels, size_t n, double frequency,
                                      double sampleRate) {
    if (!ctx) return aukit_fail("aukit_cuda: null context");
    if (channels < 1 || n < 2) return 0;                                // for i = 2, #d: nothing to do
    if (((uintptr_t)d & 15) != 0 || (channels > 1 && stride % 4 != 0))
        return aukit_fail("aukit_cuda: lowpass needs 16-byte aligned channel rows");
    const double a = 1.0 - exp(-(frequency / sampleRate) * 2.0 * 3.14159265358979323846);      // A:3589
    return lp_run(ctx, d, stride, channels, n, a, 1.0 - a, false);
}

extern "C" int aukit_cuda_dev_highpass(aukit_ctx *ctx, float *d, size_t stride, int channels, size_t n, double frequency,
                                       double sampleRate) {
    if (!ctx) return aukit_fail("aukit_cuda: null context");
    if (channels < 1 || n < 2) return 0;
    if (((uintptr_t)d & 15) != 0 || (channels > 1 && stride % 4 != 0))
        return aukit_fail("aukit_cuda: highpass needs 16-byte aligned channel rows");
    const double a = 1.0 / (2.0 * 3.14159265358979323846 * (frequency / sampleRate) + 1.0);       // A:3608
    return lp_run(ctx, d, stride, channels, n, a, a, true);
}

extern "C" int aukit_cuda_lowpass(aukit_ctx *ctx, aukit_audio *au, double frequency) {
    if (!ctx || !au) return aukit_fail("aukit_cuda: null argument");
    return aukit_cuda_dev_lowpass(ctx, au->data, au->stride, au->channels, au->frames, frequency, au->sampleRate);
}

extern "C" int aukit_cuda_highpass(aukit_ctx *ctx, aukit_audio *au, double frequency) {
    if (!ctx || !au) return aukit_fail("aukit_cuda: null argument");
    return aukit_cuda_dev_highpass(ctx, au->data, au->stride, au->channels, au->frames, frequency, au->sampleRate);
}
