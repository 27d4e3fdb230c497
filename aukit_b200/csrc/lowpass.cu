// lowpass.cu -- K11 effects.lowpass (A:3586-3598): in-place one-pole IIR per channel,
//     a = 1 - exp(-(frequency / sampleRate) * 2 * pi);   d[i] = d[i-1] + a * (d[i] - d[i-1])   for i = 2..#d
// (d[1] is left as it is).  auplay.lua:30 calls it right after effects.normalize: SURVEY 8(f) rank 1.
//
// The recurrence y[i] = (1-a) y[i-1] + a x[i] is linear, so a tile's effect on the state is the pair
// (P, S) = ((1-a)^len, end state from a zero start) and tiles combine associatively.  One pass over
// HBM (4 B read + 4 B written per sample), single kernel, chained with a decoupled look-back:
//   * a CTA claims tiles of 4096 samples in order (atomic ticket => every predecessor is running or done);
//   * the tile is staged through shared memory (coalesced 16-byte accesses both ways); thread t owns 16
//     consecutive samples and runs the reference's own step on them, first from a zero state (-> S_t);
//   * S_t are combined by a warp-shuffle scan with the constant ratio (1-a)^16, then across the 8 warps;
//   * warp 0 publishes the tile aggregate, looks back over the predecessors' aggregates / inclusive states
//     (32 at a time, stopping early once (1-a)^k has decayed below 2^-80) and publishes the inclusive state;
//   * every thread re-runs its 16 steps from its true incoming state and the tile is written back.
// Arithmetic is fp64 like the reference's (Lua numbers): in fp32 the rounding error of a low cut-off is
// amplified by 1/a and would leave the 2^-20 tolerance; the kernel stays HBM-bound either way (6 fp64 ops
// per sample).  Samples are narrowed to f32 only when stored.
#include "common.cuh"

#include <math.h>

namespace {

constexpr int LP_THREADS = 256;
constexpr int LP_PER = 16;                       // consecutive samples per thread
constexpr int LP_TILE = LP_THREADS * LP_PER;     // 4096
constexpr int LP_ROW = LP_PER + 4;               // padded row: conflict-free 16-byte accesses both ways

struct lp_state {                                // per (channel, tile)
    double agg;                                  // end state of the tile from a zero start
    double incl;                                 // true end state
};

__device__ __forceinline__ int ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ double shfl_up_d(double v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }

__global__ void __launch_bounds__(LP_THREADS)
lowpass_kernel(float *__restrict__ data, size_t stride, int channels, size_t n, double a, double b,
               lp_state *st, int *flags, unsigned long long *ticket, unsigned long long tiles_per_ch) {
    __shared__ __align__(16) float tile[LP_THREADS * LP_ROW];
    __shared__ double pt_pow[LP_THREADS];        // ((1-a)^16)^t
    __shared__ double warp_tot[LP_THREADS / 32];
    __shared__ double s_carry;
    __shared__ unsigned long long s_ticket;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const double pt = pow(b, (double)LP_PER);
    pt_pow[t] = pow(pt, (double)t);
    const double p_warp = pow(pt, 32.0), p_tile = pow(pt, (double)LP_THREADS);
    const double pl = pow(p_tile, (double)lane); // look-back weight of the lane-th predecessor
    const double p_tile32 = pow(p_tile, 32.0);
    const unsigned long long total = tiles_per_ch * (unsigned long long)channels;
    for (;;) {
        __syncthreads();
        if (t == 0) s_ticket = atomicAdd(ticket, 1ull);
        __syncthreads();
        const unsigned long long id = s_ticket;
        if (id >= total) break;
        const int ch = (int)(id / tiles_per_ch);
        const unsigned long long tl = id % tiles_per_ch;
        float *base = data + (size_t)ch * stride + (size_t)tl * LP_TILE;
        const size_t left = n - (size_t)tl * LP_TILE;
        const int cnt = left < (size_t)LP_TILE ? (int)left : LP_TILE;
        // ---- stage the tile (rows are 16-byte aligned: stride % 4 == 0 is checked by the host)
#pragma unroll
        for (int k = 0; k < LP_PER / 4; k++) {
            const int i4 = (k * LP_THREADS + t) * 4;                   // first sample of this float4
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i4 + 3 < cnt) v = *reinterpret_cast<const float4 *>(base + i4);
            else {
                if (i4 < cnt) v.x = base[i4];
                if (i4 + 1 < cnt) v.y = base[i4 + 1];
                if (i4 + 2 < cnt) v.z = base[i4 + 2];
            }
            *reinterpret_cast<float4 *>(&tile[(i4 >> 4) * LP_ROW + (i4 & 15)]) = v;
        }
        __syncthreads();
        float x[LP_PER];
#pragma unroll
        for (int k = 0; k < LP_PER / 4; k++) {
            const float4 v = *reinterpret_cast<const float4 *>(&tile[t * LP_ROW + 4 * k]);
            x[4 * k] = v.x; x[4 * k + 1] = v.y; x[4 * k + 2] = v.z; x[4 * k + 3] = v.w;
        }
        // ---- zero-start run of this thread's samples (the reference's step, A:3593-3594).  Samples past the
        // end of the channel are zeros: they only decay the state, which nothing reads afterwards.
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < LP_PER; k++) s = s + a * ((double)x[k] - s);
        // ---- inclusive scan over the block with ratio pt per thread
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
        // exclusive prefix of this thread = inclusive prefix of thread t-1
        double exc = shfl_up_d(inc, 1);
        if (lane == 0) exc = 0.0;
        // state entering thread t (zero tile carry) = pt^lane * wprev + exc
        const double enter0 = fma(pt_pow[lane], wprev, exc);
        // ---- tile aggregate + look-back (warp 0)
        if (warp == 0) {
            double agg = 0.0;
            for (int w = 0; w < LP_THREADS / 32; w++) agg = fma(p_warp, agg, warp_tot[w]);
            double carry;
            if (tl == 0) {
                // A:3591: d[1] is untouched, which is what a state equal to d[1] gives (l + a*(l - l) = l)
                carry = (double)tile[0];
            } else {
                if (lane == 0) {
                    st[id].agg = agg;
                    st_release(&flags[id], 1);
                }
                double acc = 0.0, scale = 1.0;
                long long back = (long long)tl - 1;                    // nearest predecessor tile of this channel
                carry = 0.0;
                for (;;) {
                    const long long j = back - lane;
                    int f = 2;
                    if (j >= 0) {
                        const int *fp = &flags[(unsigned long long)ch * tiles_per_ch + (unsigned long long)j];
                        do { f = ld_acquire(fp); } while (f == 0);
                    }
                    const unsigned incl_mask = __ballot_sync(0xffffffffu, j >= 0 && f == 2);
                    const unsigned none_mask = __ballot_sync(0xffffffffu, j < 0);
                    // first lane holding an inclusive state, or the first lane before tile 0
                    const int stop = __ffs(incl_mask | none_mask) - 1;    // -1: neither in this window
                    double term = 0.0;
                    if (j >= 0 && (stop < 0 || lane <= stop)) {
                        const lp_state sv = st[(unsigned long long)ch * tiles_per_ch + (unsigned long long)j];
                        term = pl * ((stop >= 0 && lane == stop && f == 2) ? sv.incl : sv.agg);
                    }
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
                st[id].incl = fma(p_tile, carry, agg);
                st_release(&flags[id], 2);
                s_carry = carry;
            }
        }
        __syncthreads();
        // ---- true run from the incoming state, written back through shared memory
        // state entering thread t = pt^t * (tile carry) + (state entering t with a zero tile carry)
        double y = fma(pt_pow[t], s_carry, enter0);
        float o[LP_PER];
#pragma unroll
        for (int k = 0; k < LP_PER; k++) {
            y = y + a * ((double)x[k] - y);
            o[k] = (float)y;
        }
#pragma unroll
        for (int k = 0; k < LP_PER / 4; k++)
            *reinterpret_cast<float4 *>(&tile[t * LP_ROW + 4 * k]) = make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < LP_PER / 4; k++) {
            const int i4 = (k * LP_THREADS + t) * 4;
            const float4 v = *reinterpret_cast<const float4 *>(&tile[(i4 >> 4) * LP_ROW + (i4 & 15)]);
            if (i4 + 3 < cnt) stg_stream(reinterpret_cast<float4 *>(base + i4), v);
            else {
                if (i4 < cnt) base[i4] = v.x;
                if (i4 + 1 < cnt) base[i4 + 1] = v.y;
                if (i4 + 2 < cnt) base[i4 + 2] = v.z;
            }
        }
    }
}

}  // namespace

extern "C" int aukit_cuda_dev_lowpass(aukit_ctx *ctx, float *d, size_t stride, int channels, size_t n, double frequency,
                                      double sampleRate) {
    if (!ctx) return aukit_fail("aukit_cuda: null context");
    if (channels < 1 || n < 2) return 0;                                // for i = 2, #d: nothing to do
    if (((uintptr_t)d & 15) != 0 || (channels > 1 && stride % 4 != 0))
        return aukit_fail("aukit_cuda: lowpass needs 16-byte aligned channel rows");
    const double a = 1.0 - exp(-(frequency / sampleRate) * 2.0 * 3.14159265358979323846);      // A:3589
    const double b = 1.0 - a;
    const unsigned long long tiles = (n + LP_TILE - 1) / LP_TILE, total = tiles * (unsigned long long)channels;
    // scratch: states, flags, ticket
    void *scratch = nullptr;
    const size_t st_bytes = (size_t)total * sizeof(lp_state), fl_bytes = ((size_t)total * sizeof(int) + 15) & ~(size_t)15;
    if (aukit_dev_alloc(ctx, st_bytes + fl_bytes + 16, &scratch)) return -1;
    lp_state *st = static_cast<lp_state *>(scratch);
    int *flags = reinterpret_cast<int *>(static_cast<char *>(scratch) + st_bytes);
    unsigned long long *ticket = reinterpret_cast<unsigned long long *>(static_cast<char *>(scratch) + st_bytes + fl_bytes);
    int rc = aukit_cuda_check(cudaMemsetAsync(flags, 0, fl_bytes + 16, ctx->stream), "memset");
    if (!rc) {
        unsigned long long g = total;
        const unsigned long long cap = (unsigned long long)ctx->num_sms * 8;
        if (g > cap) g = cap;
        lowpass_kernel<<<(unsigned)g, LP_THREADS, 0, ctx->stream>>>(d, stride, channels, n, a, b, st, flags, ticket, tiles);
        ctx->launches++;
        rc = aukit_cuda_check(cudaGetLastError(), "lowpass_kernel launch");
    }
    aukit_dev_free(ctx, scratch);
    return rc;
}

extern "C" int aukit_cuda_lowpass(aukit_ctx *ctx, aukit_audio *au, double frequency) {
    if (!ctx || !au) return aukit_fail("aukit_cuda: null argument");
    return aukit_cuda_dev_lowpass(ctx, au->data, au->stride, au->channels, au->frames, frequency, au->sampleRate);
}
