// adpcm.cu -- K3 ima_adpcm_wav_decode + K4 ms_adpcm_decode (+ headerless aukit.adpcm).
//
// Replaces aukit.wav's IMA block framing (A:1509-1548) + aukit.adpcm's step (A:1246-1257),
// and aukit.msadpcm (A:1283-1353).  Blocks are self-contained (the header is the full
// predictor state), so the unit of parallelism is one serial chain per (block, channel):
// lane <-> (block, channel) with channel fastest, so the lanes of a warp read whole
// 4*C-byte IMA groups (one 32 B sector for 8 channels) and each chain writes its 8 decoded
// samples as two float4 = one full 32 B sector.  The step table lives in shared memory.
// Integer arithmetic is bit-exact with the reference, including its dialect quirks:
//   - diff = ((n&7)*step >> 2) + (step >> 3)                      (A:1252)
//   - header predictor is state only, never emitted                (A:1513-1541)
//   - LITERAL mono masks the header step index with 0x0F           (A:1544)
//   - LITERAL mono MS-ADPCM re-reads block 1's header every block  (A:1331)
//   - MS prediction uses floor division by 256                     (A:1321)
//   - MS delta has no upper bound; the reference carries it as a double.  The kernel runs
//     32-bit integers while delta < 2^31 and switches that chain to the same fp64
//     operations the reference performs once it grows past that.
#include "common.cuh"
#include <cstdlib>

namespace {

__constant__ int c_ima_steps[89] = {
    7,     8,     9,     10,    11,    12,    13,    14,    16,    17,    19,    21,    23,
    25,    28,    31,    34,    37,    41,    45,    50,    55,    60,    66,    73,    80,
    88,    97,    107,   118,   130,   143,   157,   173,   190,   209,   230,   253,   279,
    307,   337,   371,   408,   449,   494,   544,   598,   658,   724,   796,   876,   963,
    1060,  1166,  1282,  1411,  1552,  1707,  1878,  2066,  2272,  2499,  2749,  3024,  3327,
    3660,  4026,  4428,  4871,  5358,  5894,  6484,  7132,  7845,  8630,  9493,  10442, 11487,
    12635, 13899, 15289, 16818, 18500, 20350, 22385, 24623, 27086, 29794, 32767};

// A:1250-1254 for one nibble
__device__ __forceinline__ int ima_step(int nib, int &pred, int &idx, const int *steps) {
    const int step = steps[idx];
    const int t = nib & 7;
    idx += (t < 4) ? -1 : (2 * t - 6);          // ima_index_table, A:156-159
    idx = min(max(idx, 0), 88);
    const int diff = ((t * step) >> 2) + (step >> 3);
    pred = (nib & 8) ? pred - diff : pred + diff;
    pred = min(max(pred, -32768), 32767);
    return pred;
}

__device__ __forceinline__ uint32_t load_u32_any(const uint8_t *p, bool aligned) {
    if (aligned) return *reinterpret_cast<const uint32_t *>(p);
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

enum { IMA_GENERAL = 0, IMA_LITERAL_MONO = 2 };

// One thread per (block, channel) chain.
//   general / literal stereo: 4*C-byte header, then `groups` groups of 4*C bytes; chain c
//     owns the little-endian uint32 at +4c of each group, nibble k = bits 4k..4k+3.
//   literal mono: 4-byte header, then every remaining byte of the block (low nibble first);
//     the last block may be short (A:1546 str_sub).
__global__ void __launch_bounds__(128)
ima_wav_kernel(const uint8_t *__restrict__ data, size_t nbytes, int blockAlign, int C, int mode,
               size_t nblocks, size_t spb, int groups, float *__restrict__ out, size_t stride,
               int *status, int word_aligned, int out_aligned) {
    __shared__ int steps[89];
    for (int i = threadIdx.x; i < 89; i += blockDim.x) steps[i] = c_ima_steps[i];
    __syncthreads();
    const size_t nchains = nblocks * (size_t)C;
    for (size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x; id < nchains;
         id += (size_t)gridDim.x * blockDim.x) {
        const size_t b = id / (size_t)C;
        const int c = (int)(id % (size_t)C);
        const size_t start = b * (size_t)blockAlign;
        const uint8_t *hp = data + start + 4 * (size_t)c;
        int pred = (int)(int16_t)((uint32_t)hp[0] | ((uint32_t)hp[1] << 8));
        int idx = hp[2];
        if (mode == IMA_LITERAL_MONO) idx &= 0x0F;                     // A:1544
        if (idx > 88) { atomicOr(status, AUKIT_DEVERR_IMA_INDEX); idx = 88; }
        float *o = out + (size_t)c * stride + b * spb;
        const bool vec_store = out_aligned && (spb % 4 == 0);
        if (mode != IMA_LITERAL_MONO) {
            const size_t hdr = 4 * (size_t)C;
            const uint8_t *gp = data + start + hdr + 4 * (size_t)c;
            uint32_t w = groups > 0 ? load_u32_any(gp, word_aligned) : 0u;
            for (int g = 0; g < groups; g++, o += 8) {
                gp += hdr;
                const uint32_t wn = (g + 1 < groups) ? load_u32_any(gp, word_aligned) : 0u;   // prefetch: the chain is serial
                float v[8];
#pragma unroll
                for (int k = 0; k < 8; k++) v[k] = s16_to_float(ima_step((w >> (4 * k)) & 0xF, pred, idx, steps));
                if (vec_store) {
                    stg_stream(reinterpret_cast<float4 *>(o), make_float4(v[0], v[1], v[2], v[3]));
                    stg_stream(reinterpret_cast<float4 *>(o) + 1, make_float4(v[4], v[5], v[6], v[7]));
                } else {
#pragma unroll
                    for (int k = 0; k < 8; k++) o[k] = v[k];
                }
                w = wn;
            }
        } else {
            size_t end = start + (size_t)blockAlign;
            if (end > nbytes) end = nbytes;
            const uint8_t *bp = data + start + 4;
            size_t nb = end > start + 4 ? end - (start + 4) : 0;
            while (nb >= 4) {
                const uint32_t w = load_u32_any(bp, word_aligned);
                float v[8];
#pragma unroll
                for (int k = 0; k < 8; k++) v[k] = s16_to_float(ima_step((w >> (4 * k)) & 0xF, pred, idx, steps));
                if (vec_store) {
                    stg_stream(reinterpret_cast<float4 *>(o), make_float4(v[0], v[1], v[2], v[3]));
                    stg_stream(reinterpret_cast<float4 *>(o) + 1, make_float4(v[4], v[5], v[6], v[7]));
                } else {
#pragma unroll
                    for (int k = 0; k < 8; k++) o[k] = v[k];
                }
                bp += 4; nb -= 4; o += 8;
            }
            for (; nb > 0; nb--, bp++, o += 2) {
                const int byte = *bp;
                o[0] = s16_to_float(ima_step(byte & 0xF, pred, idx, steps));
                o[1] = s16_to_float(ima_step(byte >> 4, pred, idx, steps));
            }
        }
    }
}

// Headerless aukit.adpcm on a nibble string (A:1183-1274): one chain per channel.
// interleaved: nibble m (0-based, stream order) belongs to channel m % C; otherwise channel j
// owns nibbles [j*len, (j+1)*len).  Serial by construction (no block headers to restart from).
__global__ void adpcm_stream_kernel(const uint8_t *__restrict__ data, size_t len, int C, int topFirst,
                                    int interleaved, const int *__restrict__ pred0,
                                    const int *__restrict__ idx0, float *__restrict__ out, size_t stride) {
    __shared__ int steps[89];
    for (int i = threadIdx.x; i < 89; i += blockDim.x) steps[i] = c_ima_steps[i];
    __syncthreads();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    int pred = pred0 ? pred0[c] : 0, idx = idx0 ? idx0[c] : 0;
    for (size_t i = 0; i < len; i++) {
        const size_t m = interleaved ? i * (size_t)C + c : (size_t)c * len + i;
        const int byte = data[m >> 1];
        const bool first = (m & 1) == 0;
        const int nib = (first == (topFirst != 0)) ? (byte >> 4) : (byte & 0xF);
        out[(size_t)c * stride + i] = s16_to_float(ima_step(nib, pred, idx, steps));
    }
}

__constant__ int c_ms_adapt[16] = {230, 230, 230, 230, 307, 409, 512, 614,      // nibble 0..7
                                   768, 614, 512, 409, 307, 230, 230, 230};     // nibble 8..15 = -8..-1

struct ms_coefs { int c1[256], c2[256]; int n; };   // 2052 bytes: passed by value as a __grid_constant__ kernel parameter

// One thread per (block, channel) chain.  Nibble stream after the 7*C-byte header is
// high-nibble-first; nibble m belongs to channel m % C (A:1317-1347 for C = 1, 2).
__global__ void __launch_bounds__(128)
ms_adpcm_kernel(const uint8_t *__restrict__ data, int blockAlign, int C, int literal_mono,
                size_t nblocks, size_t spb, const __grid_constant__ ms_coefs coefs,
                float *__restrict__ out, size_t stride, int *status, int vec_ok) {
    __shared__ int adapt[16];
    if (threadIdx.x < 16) adapt[threadIdx.x] = c_ms_adapt[threadIdx.x];
    __syncthreads();
    const size_t nchains = nblocks * (size_t)C;
    const int ncoef = coefs.n;
    for (size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x; id < nchains;
         id += (size_t)gridDim.x * blockDim.x) {
        const size_t b = id / (size_t)C;
        const int c = (int)(id % (size_t)C);
        const size_t start = b * (size_t)blockAlign;
        const uint8_t *hp = data + (literal_mono ? 0 : start);         // A:1331: block 1's header
        int pi = hp[c];
        if (pi >= ncoef) { atomicOr(status, AUKIT_DEVERR_MS_PREDICTOR); pi = 0; }
        const int c1 = coefs.c1[pi], c2 = coefs.c2[pi];
        auto rd16 = [&](size_t off) { return (int)(int16_t)((uint32_t)hp[off] | ((uint32_t)hp[off + 1] << 8)); };
        int delta = rd16((size_t)C + 2 * (size_t)c);
        int s1 = rd16(3 * (size_t)C + 2 * (size_t)c);
        int s2 = rd16(5 * (size_t)C + 2 * (size_t)c);
        float *o = out + (size_t)c * stride + b * spb;
        o[0] = s16_to_float(s2);                                        // A:1312-1315
        o[1] = s16_to_float(s1);
        const uint8_t *np = data + start + 7 * (size_t)C;
        bool big = false;
        double ds1 = 0, ds2 = 0, dd = 0;                                // fp64 mirror once delta >= 2^31
        // Integer fast path in 32 bits: s1*c1 + s2*c2 fits when |c1| + |c2| <= 65535; nib * delta is evaluated
        // with delta saturated at 2^24 (any larger delta drives the clamp to the same side: |lin| < 2^24); the
        // delta update uses 32 bits while |delta| < 2^21 and 64 bits above.
        const bool narrow = (abs(c1) + abs(c2)) <= 65535;
        const bool evenC = (C & 1) == 0;
        const uint8_t *bp = np + (c >> 1);                              // even C: one byte per sample, C/2 apart
        const int bstride = C >> 1, hi_nib = (c & 1) == 0;
        // one sample of this chain (A:1319-1324 / A:1338-1347)
        auto step = [&](size_t k) -> float {
            int byte;
            bool hi;
            if (evenC) { byte = bp[k * (size_t)bstride]; hi = hi_nib; }
            else { const size_t m = k * (size_t)C + (size_t)c; byte = np[m >> 1]; hi = (m & 1) == 0; }
            const int un = hi ? (byte >> 4) : (byte & 0xF);
            const int nib = un >= 8 ? un - 16 : un;                     // A:1319-1320
            if (!big) {
                int p;
                if (narrow) {
                    const int lin = (s1 * c1 + s2 * c2) >> 8;           // floor(/256), A:1321
                    const int dsat = delta > (1 << 24) ? (1 << 24) : delta;
                    p = lin + nib * dsat;
                } else {
                    const long long lin = ((long long)s1 * c1 + (long long)s2 * c2) >> 8;
                    long long pl = lin + (long long)nib * delta;
                    p = pl < -32768 ? -32768 : (pl > 32767 ? 32767 : (int)pl);
                }
                p = p < -32768 ? -32768 : (p > 32767 ? 32767 : p);
                s2 = s1; s1 = p;
                if (delta < (1 << 21) && delta > -(1 << 21)) {
                    const int nd = (adapt[un] * delta) >> 8;            // A:1324
                    delta = nd < 16 ? 16 : nd;
                } else {
                    long long nd = ((long long)adapt[un] * delta) >> 8;
                    if (nd < 16) nd = 16;
                    if (nd >= (1ll << 31)) { big = true; ds1 = (double)s1; ds2 = (double)s2; dd = (double)nd; }
                    else delta = (int)nd;
                }
                return s16_to_float(p);
            }
            // the reference's own double arithmetic (Lua numbers), A:1321-1324
            double p = floor(__dadd_rn(__dmul_rn(ds1, (double)c1), __dmul_rn(ds2, (double)c2)) / 256.0);
            p = __dadd_rn(p, __dmul_rn((double)nib, dd));
            p = p < -32768.0 ? -32768.0 : (p > 32767.0 ? 32767.0 : p);   // NaN passes, A:228
            ds2 = ds1; ds1 = p;
            const double nd = floor(__dmul_rn((double)adapt[un], dd) / 256.0);
            dd = (16.0 > nd) ? 16.0 : nd;                                  // math.max(nd, 16)
            return (float)(p / (p < 0 ? 32768.0 : 32767.0));
        };
        // stores: each lane owns its own output row, so 4-byte stores would touch 32 sectors per warp
        // instruction; group 4 samples into one 16-byte store once the row position is 16-byte aligned
        const size_t nk = spb - 2;
        float *os = o + 2;
        size_t k = 0;
        if (vec_ok) {
            const size_t head = (4 - ((b * spb + 2) & 3)) & 3;
            for (; k < head && k < nk; k++) os[k] = step(k);
            for (; k + 4 <= nk; k += 4) {
                float4 v;
                v.x = step(k); v.y = step(k + 1); v.z = step(k + 2); v.w = step(k + 3);
                stg_stream(reinterpret_cast<float4 *>(os + k), v);
            }
        }
        for (; k < nk; k++) os[k] = step(k);
    }
}

// ------------------------------------------------------------------ warp-tiled variants (the fast paths)
// The chain-per-lane kernels above store 16 bytes per lane per instruction into 32 different rows:
// 32 sector requests per STG, which fills the LSU queue (ncu: lg_throttle + mio_throttle were the top
// stalls).  The tiled kernels keep lane <-> chain but stage 16 samples per chain in shared memory
// (32 rows x 64 B per warp, XOR-swizzled so both the row-wise STS.128 and the transposed LDS.128 are
// conflict-free) and flush with each quarter-warp writing two whole 64 B row pieces: 16 full-sector
// requests per STG instead of 32 half-sector ones.

constexpr size_t ROW_NONE = ~(size_t)0;

// Per-warp staging tile: 32 rows (one per chain) x NQ float4.  Row-wise STS.128 (lane = row, fixed column)
// and transposed LDS.128 (consecutive lanes = consecutive columns of a row) are both conflict-free under
// the XOR swizzle below.  NQ = 4, 8 or 16: 64, 128 or 256 bytes per row per flush.
template <int NQ> __device__ __forceinline__ int stage_swz(int r) { return NQ == 4 ? ((r >> 1) & 3) : (r & 7); }

template <int NQ>
__device__ __forceinline__ void stage_put(float4 *st, int lane, int q, float4 v) {
    st[lane * NQ + (q ^ stage_swz<NQ>(lane))] = v;
}

// nq = valid float4 columns (1..NQ); rowb[r] = output float index of row r's block start (ROW_NONE = no chain)
template <int NQ>
__device__ __forceinline__ void stage_flush(const float4 *st, float *out, const size_t *rowb, size_t col0, int lane, int nq) {
    constexpr int RPI = 32 / NQ;                                        // rows per instruction
    const int q = lane % NQ;
#pragma unroll
    for (int i = 0; i < NQ; i++) {
        const int r = i * RPI + lane / NQ;
        const float4 v = st[r * NQ + (q ^ stage_swz<NQ>(r))];
        const size_t base = rowb[r];
        if (q < nq && base != ROW_NONE) stg_stream(reinterpret_cast<float4 *>(out + base + col0) + q, v);
    }
}

// IMA transition table: entry(idx, nib) = float bits of (signed diff * 2^-16) with next_idx in the low 7
// bits (the float needs 16 significant bits, so its low 8 mantissa bits are free).  One LDS.32 replaces
// the step lookup, the diff arithmetic (A:1252), the sign select and the index update + clamp
// (A:1250-1254).  1424 entries (5.6 KB), rebuilt per CTA.  32-bit entries on purpose: with 8-byte
// entries the random (idx, nibble) accesses cost ~5 shared-memory wavefronts per sample and the kernel
// became shared-memory bound (ncu, r1); 4-byte ones cost ~2.3.
//
// The predictor is carried as u = (pred + 32768) * 2^-16 in [0, 65535/65536]: u + d is exact in fp32
// (both are multiples of 2^-16 below 2), the lower clamp (A:1253) is the saturate of the add, the upper
// one a single min, and the output p / (p < 0 and 32768 or 32767) follows without an int->float convert:
//   p * 2^-15 = 2u - 1 (exact),  max(p, 0) * 2^-16 = saturate(u - 1/2) (exact),
//   out = fma(max(p,0) * 2^-16, 2^16 / (32767 * 32768), p * 2^-15)  -- the same FMA s16_to_float does.
constexpr int IMA_TAB = 89 * 16;

__device__ __forceinline__ void ima_build_tab(uint32_t *tab) {
    for (int e = threadIdx.x; e < IMA_TAB; e += blockDim.x) {
        const int idx = e >> 4, nib = e & 15, t = nib & 7;
        const int step = c_ima_steps[idx];
        const int diff = ((t * step) >> 2) + (step >> 3);
        int nidx = idx + ((t < 4) ? -1 : (2 * t - 6));
        nidx = min(max(nidx, 0), 88);
        tab[e] = __float_as_uint((float)((nib & 8) ? -diff : diff) * (1.0f / 65536.0f)) | (uint32_t)nidx;
    }
}

// state: u, and e = the previous table entry (its low 7 bits = current step index); x4 = nibble * 4
__device__ __forceinline__ float ima_tab_step(const char *tab, uint32_t x4, float &u, uint32_t &e) {
    e = *reinterpret_cast<const uint32_t *>(tab + ((e & 0x7F) * 64 + x4));
    u = fminf(__saturatef(__fadd_rn(u, __uint_as_float(e & 0xFFFFFF80u))), 65535.0f / 65536.0f);
    const float lo = __fmaf_rn(u, 2.0f, -1.0f);
    return __fmaf_rn(__saturatef(__fadd_rn(u, -0.5f)), (1.0f / (32767.0f * 32768.0f)) * 65536.0f, lo);
}

__device__ __forceinline__ void ima_word(const char *tab, uint32_t w, float &u, uint32_t &e, float4 &a, float4 &b) {
    a.x = ima_tab_step(tab, (w << 2) & 0x3C, u, e);
    a.y = ima_tab_step(tab, (w >> 2) & 0x3C, u, e);
    a.z = ima_tab_step(tab, (w >> 6) & 0x3C, u, e);
    a.w = ima_tab_step(tab, (w >> 10) & 0x3C, u, e);
    b.x = ima_tab_step(tab, (w >> 14) & 0x3C, u, e);
    b.y = ima_tab_step(tab, (w >> 18) & 0x3C, u, e);
    b.z = ima_tab_step(tab, (w >> 22) & 0x3C, u, e);
    b.w = ima_tab_step(tab, (w >> 26) & 0x3C, u, e);
}

// ---- IMA, general / literal-stereo block layout, output segments aligned in absolute address.  A row (one block of one channel, spb = 8 * groups floats)
// starts on a 32-byte boundary but rarely on a 256-byte one (config 4: 8160-byte rows), so 256-byte flushes that
// start at the row start straddle 128-byte lines: measured 0.73 of the copy rate, against 0.82 for the same kernel on
// rows that are a multiple of 256 bytes.  Here a lane's iteration i decodes the words of ITS row that fall into the
// row's i-th 256-byte ALIGNED segment: words 8 i - G0 + s, s = 0..7, with G0 = (row address % 256) / 32 -- the lanes of
// a warp run up to 7 words apart through their blocks, and every interior flush writes whole aligned segments.
// Iteration 0 (short by G0 words) and the last one or two (partial) take a per-word loop.
// Requires 4-byte aligned input, 32-byte aligned rows (out % 32 == 0, stride % 8 == 0).
__global__ void __launch_bounds__(128, 5)
ima_wav_aligned_kernel(const uint8_t *__restrict__ data, int blockAlign, int C, size_t nblocks, size_t spb,
                       int groups, float *__restrict__ out, size_t stride, int *status) {
    constexpr int NQ = 16, GP = 8;
    __shared__ __align__(16) uint32_t tab[IMA_TAB];
    __shared__ float4 stage_all[4][32 * NQ];
    __shared__ uintptr_t rowp_all[4][32];
    ima_build_tab(tab);
    __syncthreads();
    const char *tb = reinterpret_cast<const char *>(tab);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4 *st = stage_all[warp];
    uintptr_t *rowp = rowp_all[warp];
    const size_t nchains = nblocks * (size_t)C;
    const size_t ntiles = (nchains + 31) / 32;
    const size_t hdr = 4 * (size_t)C;
    const int M = groups / GP;                                           // iterations 1 .. M - 1 are full for every lane
    const int niter = (groups + 7 + GP - 1) / GP;                        // segments a row can touch (G0 <= 7)
    const int q = lane & 15;
    for (size_t tile = (size_t)blockIdx.x * 4 + warp; tile < ntiles; tile += (size_t)gridDim.x * 4) {
        const size_t id = min(tile * 32 + lane, nchains - 1);            // surplus lanes shadow the last chain
        const size_t b = id / (size_t)C;
        const int c = (int)(id % (size_t)C);
        const uintptr_t orow = (uintptr_t)(out + (size_t)c * stride + b * spb);
        const int G0 = (int)((orow & 255) >> 5);
        rowp[lane] = tile * 32 + lane < nchains ? orow : 0;
        const uint8_t *gp = data + b * (size_t)blockAlign + 4 * (size_t)c;
        const uint32_t hw = *reinterpret_cast<const uint32_t *>(gp);
        float pred = (float)((int)(int16_t)(hw & 0xFFFF) + 32768) * (1.0f / 65536.0f);      // u, see above
        int idx = (hw >> 16) & 0xFF;
        if (idx > 88) { atomicOr(status, AUKIT_DEVERR_IMA_INDEX); idx = 88; }
        uint32_t row = (uint32_t)idx;                                    // "previous entry": only its index bits matter
        gp += hdr;                                                       // word 0 of this chain
        uint32_t w[GP];
        // words of iteration 1 (if it is a full one): 8 - G0 + s
        if (M > 1) {
#pragma unroll
            for (int g = 0; g < GP; g++) w[g] = *reinterpret_cast<const uint32_t *>(gp + (size_t)(GP - G0 + g) * hdr);
        }
        for (int i = 0; i < niter; i++) {
            const int w0 = GP * i - G0;                                  // word in slot 0
            if (i >= 1 && i < M) {
                uint32_t n[GP];
#pragma unroll
                for (int g = 0; g < GP; g++)                             // prefetch: the chain itself is serial
                    n[g] = (i + 1 < M) ? *reinterpret_cast<const uint32_t *>(gp + (size_t)(w0 + GP + g) * hdr) : 0u;
#pragma unroll
                for (int g = 0; g < GP; g++) {
                    float4 a, bq;
                    ima_word(tb, w[g], pred, row, a, bq);
                    stage_put<NQ>(st, lane, 2 * g, a);
                    stage_put<NQ>(st, lane, 2 * g + 1, bq);
                }
                __syncwarp();
#pragma unroll
                for (int k = 0; k < NQ; k++) {                           // rows 2k / 2k + 1: whole aligned 256-byte segments
                    const int r = 2 * k + (lane >> 4);
                    const float4 v = st[r * NQ + (q ^ stage_swz<NQ>(r))];
                    const uintptr_t p = rowp[r];
                    if (p) stg_stream(reinterpret_cast<float4 *>((p & ~(uintptr_t)255) + (size_t)i * (16 * NQ)) + q, v);
                }
#pragma unroll
                for (int g = 0; g < GP; g++) w[g] = n[g];
            } else {
                const int sa = max(0, -w0), sb = min(GP, groups - w0);
                for (int sl = sa; sl < sb; sl++) {
                    float4 a, bq;
                    ima_word(tb, *reinterpret_cast<const uint32_t *>(gp + (size_t)(w0 + sl) * hdr), pred, row, a, bq);
                    stage_put<NQ>(st, lane, 2 * sl, a);
                    stage_put<NQ>(st, lane, 2 * sl + 1, bq);
                }
                __syncwarp();
                for (int k = 0; k < NQ; k++) {
                    const int r = 2 * k + (lane >> 4);
                    const float4 v = st[r * NQ + (q ^ stage_swz<NQ>(r))];
                    const uintptr_t p = rowp[r];
                    const int j = 2 * (GP * i - (int)((p & 255) >> 5)) + q;      // row r's quad in slot q
                    if (p && j >= 0 && j < 2 * groups) stg_stream(reinterpret_cast<float4 *>(p) + j, v);
                }
            }
            __syncwarp();
        }
    }
}

// ---- MS-ADPCM, warp-tiled.  Fast path per 4 samples in 32-bit integers while delta < 2^14 (then
// delta < 2^14 * 3^4 < 2^21 inside the quad, so adapt*delta and nib*delta cannot overflow) and
// |c1| + |c2| <= 65535; anything else (huge deltas, the fp64 continuation) goes through the
// out-of-line general step, which carries the same state the chain-per-lane kernel does.
//
// Input records: with the two header samples counted as samples 0 and 1 of the block, the nibbles of
// samples 4j..4j+3 of ALL channels are the 2C bytes at block offset 2C(3 + j) (record 0 starts inside
// the header: its second half holds the first two coded frames).  For C in {1, 2, 4, 8} one aligned
// 2C-byte load per quad fetches the record and the chain's four nibbles come out with constant shifts.
struct ms_state { int s1, s2, delta, big; double ds1, ds2, dd; };

__device__ __noinline__ float ms_general_step(ms_state *st, int nib, int c1, int c2) {
    const int ad = c_ms_adapt[nib & 15];
    if (!st->big) {
        const long long lin = ((long long)st->s1 * c1 + (long long)st->s2 * c2) >> 8;   // floor(/256), A:1321
        const long long pl = lin + (long long)nib * st->delta;
        const int p = pl < -32768 ? -32768 : (pl > 32767 ? 32767 : (int)pl);
        st->s2 = st->s1; st->s1 = p;
        long long nd = ((long long)ad * st->delta) >> 8;                                // A:1324
        if (nd < 16) nd = 16;
        if (nd >= (1ll << 31)) {
            st->big = 1; st->ds1 = (double)st->s1; st->ds2 = (double)st->s2; st->dd = (double)nd;
            st->delta = 0x7FFFFFFF;                                     // keeps the caller on this path
        } else st->delta = (int)nd;
        return s16_to_float(p);
    }
    // the reference's own double arithmetic (Lua numbers), A:1321-1324
    double p = floor(__dadd_rn(__dmul_rn(st->ds1, (double)c1), __dmul_rn(st->ds2, (double)c2)) / 256.0);
    p = __dadd_rn(p, __dmul_rn((double)nib, st->dd));
    p = p < -32768.0 ? -32768.0 : (p > 32767.0 ? 32767.0 : p);          // NaN passes, A:228
    st->ds2 = st->ds1; st->ds1 = p;
    const double nd = floor(__dmul_rn((double)ad, st->dd) / 256.0);
    st->dd = (16.0 > nd) ? 16.0 : nd;                                   // math.max(nd, 16)
    return (float)(p / (p < 0 ? 32768.0 : 32767.0));
}

__device__ __forceinline__ int sx4(uint32_t w, int left) { return (int)(w << left) >> 28; }   // signed nibble

// One input record (the nibbles of 4 consecutive samples of all channels).  TC = channel count when it
// is 1, 2, 4 or 8 and the records are 2C-aligned: a single vector load, issued one quad ahead of its use;
// TC = 0 reads byte by byte for any C.  cs = bit offset of the chain's nibble inside a frame (high
// nibble first: 8*(c/2) + (c odd ? 0 : 4)).
template <int TC> struct ms_rec { uint32_t w[TC == 8 ? 4 : (TC == 4 ? 2 : 1)]; };

template <int TC>
__device__ __forceinline__ ms_rec<TC> ms_load_rec(const uint8_t *blk, int j) {
    ms_rec<TC> r;
    if (TC == 8) {
        const uint4 v = *reinterpret_cast<const uint4 *>(blk + 16 * (3 + j));
        r.w[0] = v.x; r.w[1] = v.y; r.w[2] = v.z; r.w[3] = v.w;
    } else if (TC == 4) {
        const uint2 v = *reinterpret_cast<const uint2 *>(blk + 8 * (3 + j));
        r.w[0] = v.x; r.w[1] = v.y;
    } else if (TC == 2) {
        r.w[0] = *reinterpret_cast<const uint32_t *>(blk + 4 * (3 + j));
    } else if (TC == 1) {
        r.w[0] = *reinterpret_cast<const uint16_t *>(blk + 2 * (3 + j));
    } else {
        r.w[0] = 0;
    }
    return r;
}

template <int TC>
__device__ __forceinline__ void ms_record_nibs(const ms_rec<TC> &r, const uint8_t *blk, int C, int c, int cs, int j,
                                               int (&nb)[4]) {
    if (TC == 8) {
        const uint32_t mul = 1u << (28 - cs);                           // shift as a multiply: IMAD runs on the FMA pipe
        nb[0] = (int)(r.w[0] * mul) >> 28; nb[1] = (int)(r.w[1] * mul) >> 28;
        nb[2] = (int)(r.w[2] * mul) >> 28; nb[3] = (int)(r.w[3] * mul) >> 28;
    } else if (TC == 4) {
        const uint32_t a = r.w[0] >> cs, b = r.w[TC == 4 ? 1 : 0] >> cs;
        nb[0] = sx4(a, 28); nb[1] = sx4(a, 12); nb[2] = sx4(b, 28); nb[3] = sx4(b, 12);
    } else if (TC == 2) {
        const uint32_t a = r.w[0] >> cs;
        nb[0] = sx4(a, 28); nb[1] = sx4(a, 20); nb[2] = sx4(a, 12); nb[3] = sx4(a, 4);
    } else if (TC == 1) {
        const uint32_t a = r.w[0];
        nb[0] = sx4(a, 24); nb[1] = sx4(a, 28); nb[2] = sx4(a, 16); nb[3] = sx4(a, 20);
    } else {
        const uint8_t *rec = blk + 2 * (size_t)C * (size_t)(3 + j);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int n = i * C + c;
            const int byte = rec[n >> 1];
            nb[i] = sx4((uint32_t)byte, (n & 1) ? 28 : 24);
        }
    }
}

// requires spb % 4 == 0 and 16-byte aligned output rows; chain layout as ms_adpcm_kernel
template <int TC, int NQ>
__global__ void __launch_bounds__(128, NQ == 16 ? 6 : 8)               // NQ = 16: shared memory allows 6 CTAs anyway
ms_adpcm_tiled_kernel(const uint8_t *__restrict__ data, int blockAlign, int C, int literal_mono,
                      size_t nblocks, size_t spb, const __grid_constant__ ms_coefs coefs,
                      float *__restrict__ out, size_t stride, int *status) {
    // adaptation table indexed by the signed nibble, 3 words apart: the odd stride keeps the 16 entries in
    // 16 different banks and makes the address an IMAD (FMA pipe) instead of a LEA (ALU pipe)
    __shared__ int adapt_s[48];
    __shared__ float4 stage_all[4][32 * NQ];
    __shared__ size_t rowb_all[4][32];
    if (threadIdx.x < 16) adapt_s[threadIdx.x * 3] = c_ms_adapt[(threadIdx.x - 8) & 15];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4 *st = stage_all[warp];
    size_t *rowb = rowb_all[warp];
    const size_t nchains = nblocks * (size_t)C;
    const size_t ntiles = (nchains + 31) / 32;
    const int ncoef = coefs.n;
    const int nquads = (int)(spb / 4);
    const int *adapt_mid = adapt_s + 24;
    for (size_t tile = (size_t)blockIdx.x * 4 + warp; tile < ntiles; tile += (size_t)gridDim.x * 4) {
        const size_t id = min(tile * 32 + lane, nchains - 1);
        const size_t b = id / (size_t)C;
        const int c = (int)(id % (size_t)C);
        rowb[lane] = tile * 32 + lane < nchains ? (size_t)c * stride + b * spb : ROW_NONE;
        const uint8_t *blk = data + b * (size_t)blockAlign;
        const uint8_t *hp = literal_mono ? data : blk;                  // A:1331: block 1's header
        int pi = hp[c];
        if (pi >= ncoef) { atomicOr(status, AUKIT_DEVERR_MS_PREDICTOR); pi = 0; }
        const int c1 = coefs.c1[pi], c2 = coefs.c2[pi];
        auto rd16 = [&](size_t off) { return (int)(int16_t)((uint32_t)hp[off] | ((uint32_t)hp[off + 1] << 8)); };
        int delta = rd16((size_t)C + 2 * (size_t)c);
        int s1 = rd16(3 * (size_t)C + 2 * (size_t)c);
        int s2 = rd16(5 * (size_t)C + 2 * (size_t)c);
        const int cs = 8 * (c >> 1) + ((c & 1) ? 0 : 4);
        const bool narrow = (abs(c1) + abs(c2)) <= 65535;
        ms_state gs;
        gs.big = 0;
        auto fast = [&](int nib) -> float {
            int p = ((s1 * c1 + s2 * c2) >> 8) + nib * delta;           // A:1321-1322
            p = min(max(p, -32768), 32767);
            s2 = s1; s1 = p;
            delta = max((adapt_mid[nib * 3] * delta) >> 8, 16);         // A:1324
            return s16_to_float(p);
        };
        auto general = [&](int nib) -> float {
            gs.s1 = s1; gs.s2 = s2; gs.delta = delta;
            const float v = ms_general_step(&gs, nib, c1, c2);
            s1 = gs.s1; s2 = gs.s2; delta = gs.delta;
            return v;
        };
        // records are fetched four quads ahead of their use (a ring of 4, statically indexed inside the unrolled
        // period): ncu showed the chain stalled on the record loads with only two in flight
        ms_rec<TC> ring[4];
#pragma unroll
        for (int i = 0; i < 4; i++) ring[i] = ms_load_rec<TC>(blk, i < nquads ? i : 0);
        // samples 4j..4j+3 of the block; `first`: the two header samples lead (A:1312-1315)
        auto quad = [&](int j, bool first, const ms_rec<TC> &rec) -> float4 {
            int nb[4];
            ms_record_nibs<TC>(rec, blk, C, c, cs, j, nb);
            const bool quick = narrow && delta < (1 << 14) && delta > -(1 << 16);
            float4 v;
            if (first) {
                v.x = s16_to_float(s2); v.y = s16_to_float(s1);
                if (quick) { v.z = fast(nb[2]); v.w = fast(nb[3]); }
                else { v.z = general(nb[2]); v.w = general(nb[3]); }
            } else if (quick) {
                v.x = fast(nb[0]); v.y = fast(nb[1]); v.z = fast(nb[2]); v.w = fast(nb[3]);
            } else {
                v.x = general(nb[0]); v.y = general(nb[1]); v.z = general(nb[2]); v.w = general(nb[3]);
            }
            return v;
        };
        size_t col = 0;
        const int nper = nquads / NQ;
        for (int per = 0; per < nper; per++, col += 4 * NQ) {
            if (TC > 0) {
                // the records of one period span NQ * 2C bytes; every eighth record load opened a new 128-byte line
                // and waited for DRAM (ncu: a third of all stall samples) -- ask for the next period's lines now
                const uint8_t *nx = blk + (size_t)(2 * TC) * (size_t)(3 + NQ * (per + 1));
#pragma unroll
                for (int o = 0; o < NQ * 2 * TC; o += 128)
                    if (nx + o < blk + blockAlign) asm volatile("prefetch.global.L1 [%0];" ::"l"(nx + o));
            }
#pragma unroll
            for (int q = 0; q < NQ; q++) {
                const int j = NQ * per + q;
                const ms_rec<TC> rec = ring[q & 3];
                if (j + 4 < nquads) ring[q & 3] = ms_load_rec<TC>(blk, j + 4);
                stage_put<NQ>(st, lane, q, quad(j, q == 0 && per == 0, rec));
            }
            __syncwarp();
            stage_flush<NQ>(st, out, rowb, col, lane, NQ);
            __syncwarp();
        }
        const int rem = nquads - nper * NQ;
        if (rem) {
            for (int q = 0; q < rem; q++)
                stage_put<NQ>(st, lane, q, quad(NQ * nper + q, nper == 0 && q == 0, ms_load_rec<TC>(blk, NQ * nper + q)));
            __syncwarp();
            stage_flush<NQ>(st, out, rowb, col, lane, rem);
            __syncwarp();
        }
    }
}


// ---- MS-ADPCM, 8 channels (config 4): staged input, straight-line periods, bulk-copy output.
// ncu on the kernel above (profiles/r2_ms_adpcm_ncu.txt): a warp issued in 12 % of its cycles.  The chains waited on
// their record loads although these were issued four quads ahead, every quad was its own basic block (the
// `quick` test branches to the out-of-line general step), so no load could move across a quad, and the flush's
// LDS.128 -> STG.128 pairs stalled on each other's registers; the ALU pipe -- half rate on sm_100
// (profiles/r2_microbench_int.txt) -- carried 11 of the 25.6 instructions per sample.  Here
//   * a warp's records travel global -> shared with cp.async one period (1 KB per warp) ahead of their use; the
//     chain reads its record with one LDS.128 (the 8 lanes of a block share it: a broadcast);
//   * a period (16 quads = 64 samples) is SPECULATED in the 32-bit fast arithmetic as one straight-line block:
//     the `delta < 2^14` test of each quad only ORs into a flag, and a lane whose flag is set at the end restores
//     the state it had at the start of the period and redoes it through the checked / general code (never for
//     encoder-made data; wrapping int arithmetic cannot trap);
//   * the predictor pair is carried offset-binary, u = s + 32768 in [0, 65535]: the offset's contribution to
//     s1*c1 + s2*c2 is a per-chain constant folded into the first IMAD, both clamps of A:1322 are ONE
//     VIMNMX.RELU (min(x, 65535) then max(., 0)), and u + 0x4B000000 is already the float 2^23 + u, from which
//     the output p / (p < 0 and 32768 or 32767) follows with three FFMAs and no int->float convert; the nibble
//     is isolated with a multiply (FMA pipe) + one arithmetic shift;
//   * output segments are aligned in absolute address (below) and leave through the transposed, XOR-swizzled staging
//     tile of the kernels above (a per-lane cp.async.bulk shared -> global was tried: UBLKCP takes uniform operands,
//     so ptxas serialises the 32 lanes in a loop and the wait for its reads cost 27 % of the samples).
// Same integers as the kernels above at every step (tests hold the staged, tiled and chain-per-lane kernels to the same bits).
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
constexpr int MS8_NQ = 16;                        // quads per iteration = float4 slots of a staging row = one 256-byte output segment
constexpr int MS8_RING = 4 * 16 * MS8_NQ;         // record ring of one block: 4 stages of 16 records ...
constexpr int MS8_BLK_PITCH = MS8_RING + 16 * MS8_NQ + 16;   // ... + a mirror of stage slot 0 (a lane's 16 records never wrap) + bank skew

// requires C == 8, spb % 4 == 0, 16-byte aligned input / output rows, blockAlign % 16 == 0.
//
// Output segments are aligned in ABSOLUTE address: a row (one block of one channel, spb floats) starts wherever
// c * stride + b * spb puts it -- 16 bytes past a 32-byte sector for every other block of config 4 -- and 256-byte
// flushes that start there straddle sectors and 128-byte lines (measured: 0.52 of the copy rate against 0.75 for the
// same kernel when spb * 4 is a multiple of 256).  So a lane's iteration i produces the quads of ITS row that fall
// into the i-th 256-byte aligned segment: local quads j = 16 i - Q0 + q, q = 0..15, with Q0 = (row address % 256) / 16.
// The lanes of a warp are therefore up to 15 quads apart in their blocks; they read their records out of a 4-stage
// ring per block (stage s = records 16 s .. 16 s + 15, filled by cp.async one iteration ahead).  Iteration 0 (short by
// Q0 quads, holds the header samples) and the last one or two (partial) run a plain per-quad loop.
__global__ void __launch_bounds__(128, 4)
ms_adpcm_staged8_kernel(const uint8_t *__restrict__ data, int blockAlign, size_t nblocks, size_t spb,
                        const __grid_constant__ ms_coefs coefs, float *__restrict__ out, size_t stride, int *status) {
    constexpr int NQ = MS8_NQ, C = 8;
    extern __shared__ __align__(16) uint8_t ms8_smem[];
    float4 *stage_all = reinterpret_cast<float4 *>(ms8_smem);            // [4 warps][32 rows][NQ], XOR-swizzled (stage_put)
    uint8_t *rec_all = ms8_smem + 4 * 32 * NQ * 16;                      // [4 warps][4 blocks][MS8_BLK_PITCH]
    uintptr_t *rowp_all = reinterpret_cast<uintptr_t *>(rec_all + 4 * 4 * MS8_BLK_PITCH);   // [4 warps][32]: row address, 0 = no chain
    int *adapt_s = reinterpret_cast<int *>(rowp_all + 4 * 32);           // [48]
    if (threadIdx.x < 16) adapt_s[threadIdx.x * 3] = c_ms_adapt[(threadIdx.x - 8) & 15];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4 *st = stage_all + warp * 32 * NQ;
    uintptr_t *rowp = rowp_all + warp * 32;
    uint8_t *rec_w = rec_all + warp * 4 * MS8_BLK_PITCH;
    const size_t ntiles = (nblocks + 3) / 4;
    const int ncoef = coefs.n;
    const int nquads = (int)(spb / 4);
    const int nstage = (nquads + NQ - 1) / NQ;                           // record stages of a block
    const int M = nquads / NQ;                                           // iterations 1 .. M - 1 are full for every lane
    // shared-window address of adapt[0], opaque so that it is held in a register instead of being rebuilt per quad
    uint32_t adapt_mid = (uint32_t)__cvta_generic_to_shared(adapt_s + 24);
    asm volatile("mov.u32 %0, %0;" : "+r"(adapt_mid));
    const int c = lane & 7, bl = lane >> 3;                              // channel, block within the tile
    const int cs = 8 * (c >> 1) + ((c & 1) ? 0 : 4);                     // bit offset of the chain's nibble in a frame word
    uint32_t mul = 1u << (28 - cs);
    asm volatile("mov.u32 %0, %0;" : "+r"(mul));                         // opaque: keeps w * mul an IMAD, not a shift on the ALU pipe
    const uint8_t *rb = rec_w + bl * MS8_BLK_PITCH;                      // my block's record ring
    for (size_t tile = (size_t)blockIdx.x * 4 + warp; tile < ntiles; tile += (size_t)gridDim.x * 4) {
        const bool live = tile * 4 + (size_t)bl < nblocks;               // surplus lanes shadow the last block, write nothing
        const size_t b = min(tile * 4 + (size_t)bl, nblocks - 1);
        float *orow = out + (size_t)c * stride + b * spb;
        const int Q0 = (int)(((uintptr_t)orow & 255) >> 4);
        rowp[lane] = live ? (uintptr_t)orow : 0;
        const uint8_t *blk = data + b * (size_t)blockAlign;
        // staging: chunk k = lane + 32 h covers record k % 16 of a stage of block k / 16 of the tile
        const uint8_t *src[2];
        uint8_t *dst[2];
        int rec_of[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int k = lane + 32 * h;
            const size_t kb = min(tile * 4 + (size_t)(k >> 4), nblocks - 1);
            rec_of[h] = k & 15;
            src[h] = data + kb * (size_t)blockAlign + 16 * (3 + rec_of[h]);
            dst[h] = rec_w + (k >> 4) * MS8_BLK_PITCH + 16 * rec_of[h];
        }
        auto issue = [&](int s) {
            if (s < nstage) {
#pragma unroll
                for (int h = 0; h < 2; h++)
                    if (NQ * s + rec_of[h] < nquads) {
                        cp_async16(dst[h] + (s & 3) * (16 * NQ), src[h] + (size_t)s * (16 * NQ));
                        if ((s & 3) == 0) cp_async16(dst[h] + MS8_RING, src[h] + (size_t)s * (16 * NQ));   // the mirror
                    }
            }
            cp_async_commit();
        };
        issue(0);
        int pi = blk[c];
        if (pi >= ncoef) { atomicOr(status, AUKIT_DEVERR_MS_PREDICTOR); pi = 0; }
        const int c1 = coefs.c1[pi], c2 = coefs.c2[pi];
        auto rd16 = [&](size_t off) { return (int)(int16_t)((uint32_t)blk[off] | ((uint32_t)blk[off + 1] << 8)); };
        int delta = rd16((size_t)C + 2 * (size_t)c);
        int u1 = rd16(3 * (size_t)C + 2 * (size_t)c) + 32768;
        int u2 = rd16(5 * (size_t)C + 2 * (size_t)c) + 32768;
        const bool narrow = (abs(c1) + abs(c2)) <= 16384;                // |u1 c1 + u2 c2 + kc| < 2^31
        const int kc = (int)((1u << 23) - 32768u * (unsigned)(c1 + c2)); // (s1 c1 + s2 c2) + 2^23 = u1 c1 + u2 c2 + kc
        ms_state gs;
        gs.big = 0;
        auto fast = [&](int nib) -> float {
            const int t = u1 * c1 + (u2 * c2 + kc);
            const int u = __vimin_s32_relu((t >> 8) + nib * delta, 65535);      // A:1321-1322, both clamps
            u2 = u1; u1 = u;
            int ad;
            asm("ld.shared.s32 %0, [%1];" : "=r"(ad) : "r"(adapt_mid + (uint32_t)(nib * 12)));
            delta = max((ad * delta) >> 8, 16);                                 // A:1324
            const float x = __int_as_float(u + 0x4B000000);                     // 2^23 + u
            const float lo = __fmaf_rn(x, 1.0f / 32768.0f, -257.0f);            // (u - 32768) / 32768, exact
            return __fmaf_rn(__saturatef(lo), (1.0f / (32767.0f * 32768.0f)) * 32768.0f, lo);   // s16_to_float's FMA
        };
        auto general = [&](int nib) -> float {
            gs.s1 = u1 - 32768; gs.s2 = u2 - 32768; gs.delta = delta;
            const float v = ms_general_step(&gs, nib, c1, c2);
            u1 = gs.s1 + 32768; u2 = gs.s2 + 32768; delta = gs.delta;
            return v;
        };
        auto nibs = [&](const uint8_t *rp, int (&n)[4]) {
            const uint4 r = *reinterpret_cast<const uint4 *>(rp);
            n[0] = (int)(r.x * mul) >> 28; n[1] = (int)(r.y * mul) >> 28;
            n[2] = (int)(r.z * mul) >> 28; n[3] = (int)(r.w * mul) >> 28;
        };
        // the fast arithmetic is exact for a quad that starts with -65535 < delta < 2^14 (delta <= 3^4 * 2^14 inside it)
        auto quick = [&]() { return (unsigned)(delta + 65535) < (unsigned)(16384 + 65535); };
        // local quads [j0 + qa, j0 + qb) into slots [qa, qb), each with its own validity test
        auto checked = [&](int j0, int qa, int qb) {
            for (int q = qa; q < qb; q++) {
                const int j = j0 + q;
                int n[4];
                nibs(rb + ((16 * j) & (MS8_RING - 1)), n);
                const bool qk = narrow && quick();
                float4 v;
                if (j == 0) { v.x = s16_to_float(u2 - 32768); v.y = s16_to_float(u1 - 32768); }   // the header samples lead (A:1312-1315)
                else if (qk) { v.x = fast(n[0]); v.y = fast(n[1]); }
                else { v.x = general(n[0]); v.y = general(n[1]); }
                if (qk) { v.z = fast(n[2]); v.w = fast(n[3]); }
                else { v.z = general(n[2]); v.w = general(n[3]); }
                stage_put<NQ>(st, lane, q, v);
            }
        };
        const int niter = (nquads + 15 + NQ - 1) / NQ;                   // segments a row can touch (Q0 <= 15)
        for (int i = 0; i < niter; i++) {
            issue(i + 1);
            cp_async_wait<1>();
            __syncwarp();                                                // every lane's chunks of stages <= i have landed
            const int j0 = NQ * i - Q0;                                  // local quad of slot 0
            if (i >= 1 && i < M && narrow) {
                // straight line: 16 quads, no branch; `bad` collects the per-quad validity tests
                const int su1 = u1, su2 = u2, sdelta = delta;
                const uint8_t *rp = rb + ((16 * j0) & (MS8_RING - 1));   // 16 records from here never wrap (mirror)
                bool bad = false;
#pragma unroll
                for (int q = 0; q < NQ; q++) {
                    int n[4];
                    nibs(rp + 16 * q, n);
                    bad |= (unsigned)delta >= 16384u;                    // delta >= 16 after the first step of a block
                    float4 v;
                    v.x = fast(n[0]); v.y = fast(n[1]); v.z = fast(n[2]); v.w = fast(n[3]);
                    stage_put<NQ>(st, lane, q, v);
                }
                if (bad) {
                    u1 = su1; u2 = su2; delta = sdelta;
                    checked(j0, 0, NQ);
                }
                __syncwarp();
                // transposed flush: lanes 0..15 / 16..31 write the 16 pieces of rows 2k / 2k + 1 -- whole aligned 256-byte segments
                const int q = lane & 15;
#pragma unroll
                for (int k = 0; k < NQ; k++) {
                    const int r = 2 * k + (lane >> 4);
                    const float4 v = st[r * NQ + (q ^ stage_swz<NQ>(r))];
                    const uintptr_t p = rowp[r];
                    if (p) stg_stream(reinterpret_cast<float4 *>((p & ~(uintptr_t)255) + (size_t)i * (16 * NQ)) + q, v);
                }
            } else {
                const int qa = max(0, -j0), qb = min(NQ, nquads - j0);
                if (qb > qa) checked(j0, qa, qb);
                __syncwarp();
                const int q = lane & 15;
                for (int k = 0; k < NQ; k++) {
                    const int r = 2 * k + (lane >> 4);
                    const float4 v = st[r * NQ + (q ^ stage_swz<NQ>(r))];
                    const uintptr_t p = rowp[r];
                    const int j = NQ * i - (int)((p & 255) >> 4) + q;    // row r's local quad in slot q
                    if (p && j >= 0 && j < nquads) stg_stream(reinterpret_cast<float4 *>(p) + j, v);
                }
            }
            __syncwarp();                                                // staging tile and this iteration's records are free again
        }
        cp_async_wait<0>();
    }
}

constexpr size_t MS8_SMEM = 4 * 32 * MS8_NQ * 16 + 4 * 4 * MS8_BLK_PITCH + 4 * 32 * sizeof(uintptr_t) + 48 * sizeof(int);

}  // namespace

extern "C" size_t aukit_ima_adpcm_wav_frames(size_t nbytes, int blockAlign, int channels, int dialect) {
    if (blockAlign < 1 || channels < 1) return 0;
    const size_t bA = (size_t)blockAlign, C = (size_t)channels;
    const size_t nblocks = (nbytes + bA - 1) / bA;
    if (dialect == AUKIT_DIALECT_LITERAL && channels == 2)              // for i = 8, blockAlign-1, 8
        return nblocks * (blockAlign > 8 ? ((bA - 9) / 8 + 1) * 8 : 0);
    if (dialect == AUKIT_DIALECT_LITERAL && channels == 1) {
        const size_t full = nbytes / bA, rem = nbytes % bA;
        return full * (bA > 4 ? (bA - 4) * 2 : 0) + (rem > 4 ? (rem - 4) * 2 : 0);
    }
    const size_t hdr = 4 * C;
    return (nbytes / bA) * (bA > hdr ? (bA - hdr) / hdr : 0) * 8;
}

extern "C" size_t aukit_msadpcm_frames(size_t nbytes, int blockAlign, int channels) {
    if (blockAlign < 1 || channels < 1) return 0;
    const size_t bA = (size_t)blockAlign, C = (size_t)channels;
    const size_t nblocks = (nbytes + bA - 1) / bA;
    const size_t body = bA > 7 * C ? bA - 7 * C : 0;
    return nblocks * (2 + body * 2 / C);
}

extern "C" int aukit_cuda_dev_ima_adpcm_wav(aukit_ctx *ctx, const void *d_in, size_t nbytes, int blockAlign,
                                            int channels, int dialect, float *d_out, size_t out_stride) {
    if (!ctx) return aukit_fail("aukit_cuda: null context");
    if (blockAlign < 1) return aukit_fail("'for' step must be positive");
    if (nbytes == 0) return aukit_fail("attempt to index a nil value");                      // A:1548, no blocks
    const size_t bA = (size_t)blockAlign;
    int mode = IMA_GENERAL, groups = 0;
    size_t nblocks = nbytes / bA, spb = 0;
    if (dialect == AUKIT_DIALECT_LITERAL) {
        if (channels == 2) {
            nblocks = (nbytes + bA - 1) / bA;
            groups = blockAlign > 8 ? (blockAlign - 9) / 8 + 1 : 0;
            // the last block must hold its 7-byte header and every group the loop touches
            const size_t last = (nblocks - 1) * bA;
            if (last + 7 > nbytes) return aukit_fail("data string too short");
            if (groups && last + 8 + 8 * (size_t)groups > nbytes)
                return aukit_fail("bad argument #1 to 'band' (number expected, got nil)");
        } else if (channels == 1) {
            mode = IMA_LITERAL_MONO;
            nblocks = (nbytes + bA - 1) / bA;
            const size_t last = (nblocks - 1) * bA;
            if (last + 3 > nbytes) return aukit_fail("data string too short");
        } else {
            return aukit_fail("bad argument #6 (table too short)");                           // A:1199
        }
    } else {
        if (channels < 1) return aukit_fail("aukit_cuda: channels < 1");
        if (nbytes % bA) return aukit_fail("aukit_cuda: IMA ADPCM data is not a whole number of blocks");
        if (bA < 4 * (size_t)channels) return aukit_fail("aukit_cuda: blockAlign smaller than the block header");
        groups = (int)((bA - 4 * (size_t)channels) / (4 * (size_t)channels));
    }
    if (mode == IMA_LITERAL_MONO) spb = bA > 4 ? (bA - 4) * 2 : 0;
    else spb = (size_t)groups * 8;
    const size_t frames = aukit_ima_adpcm_wav_frames(nbytes, blockAlign, channels, dialect);
    if (channels > 1 && out_stride < frames) return aukit_fail("aukit_cuda: out_stride < frames");
    const int word_aligned = ((uintptr_t)d_in % 4 == 0) && (blockAlign % 4 == 0);
    const int threads = 128;
    const unsigned grid = aukit_grid(nblocks * (size_t)channels, threads, (size_t)ctx->num_sms * 64);
    const int out_aligned = ((uintptr_t)d_out % 16 == 0) && (out_stride % 4 == 0);
    // AUKIT_DISABLE_TILED_ADPCM=1 (diagnostic, tests): the chain-per-lane kernel for everything
    if (mode != IMA_LITERAL_MONO && word_aligned && groups > 0 && (uintptr_t)d_out % 32 == 0 && out_stride % 8 == 0 &&
        !getenv("AUKIT_DISABLE_TILED_ADPCM")) {
        const unsigned g5 = aukit_grid(nblocks * (size_t)channels, threads, (size_t)ctx->num_sms * 5);
        ima_wav_aligned_kernel<<<g5, threads, 0, ctx->stream>>>(static_cast<const uint8_t *>(d_in), blockAlign, channels, nblocks, spb,
                                                              groups, d_out, out_stride, ctx->d_status);
        ctx->launches++;
        return aukit_cuda_check(cudaGetLastError(), "ima_wav_aligned_kernel launch");
    }
    ima_wav_kernel<<<grid, threads, 0, ctx->stream>>>(static_cast<const uint8_t *>(d_in), nbytes, blockAlign, channels,
                                                      mode, nblocks, spb, groups, d_out, out_stride, ctx->d_status,
                                                      word_aligned, out_aligned);
    ctx->launches++;
    return aukit_cuda_check(cudaGetLastError(), "ima_wav_kernel launch");
}

extern "C" int aukit_cuda_dev_msadpcm(aukit_ctx *ctx, const void *d_in, size_t nbytes, int blockAlign,
                                      int channels, const int *coef1, const int *coef2, int ncoef, int dialect,
                                      float *d_out, size_t out_stride) {
    if (!ctx) return aukit_fail("aukit_cuda: null context");
    if (blockAlign < 1) return aukit_fail("'for' step must be positive");
    if (nbytes == 0) return 0;
    if (dialect == AUKIT_DIALECT_LITERAL && channels != 1 && channels != 2)
        return aukit_fail("Unsupported number of channels: %d", channels);                   // A:1349
    if (channels < 1) return aukit_fail("aukit_cuda: channels < 1");
    const size_t bA = (size_t)blockAlign, C = (size_t)channels;
    if (nbytes % bA) {
        // a trailing partial block makes the reference read nil bytes (A:1318 / A:1336)
        if (nbytes % bA < 7 * C && !(dialect == AUKIT_DIALECT_LITERAL && channels == 1))
            return aukit_fail("data string too short");
        return aukit_fail("bad argument #1 to 'rshift' (number expected, got nil)");
    }
    if (bA < 7 * C) return aukit_fail("data string too short");
    ms_coefs h;
    static const int d1[7] = {256, 512, 0, 192, 240, 460, 392}, d2[7] = {0, -256, 0, 64, 0, -208, -232};  // A:1304
    if (!coef1 || !coef2 || ncoef <= 0) { coef1 = d1; coef2 = d2; ncoef = 7; }
    if (ncoef > 256) return aukit_fail("aukit_cuda: more than 256 coefficient pairs");
    for (int i = 0; i < ncoef; i++) { h.c1[i] = coef1[i]; h.c2[i] = coef2[i]; }
    h.n = ncoef;
    const size_t nblocks = nbytes / bA;
    const size_t spb = 2 + (bA - 7 * C) * 2 / C;
    if (channels > 1 && out_stride < nblocks * spb) return aukit_fail("aukit_cuda: out_stride < frames");
    const int threads = 128;
    const unsigned grid = aukit_grid(nblocks * C, threads, (size_t)ctx->num_sms * 64);
    const int vec_ok = ((uintptr_t)d_out % 16 == 0) && (out_stride % 4 == 0);
    // diagnostic switches (tests hold the three implementations to the same bits): AUKIT_DISABLE_TILED_ADPCM=1 -> the
    // chain-per-lane kernel for everything, AUKIT_DISABLE_STAGED_ADPCM=1 -> the register-staged tiled kernel for 8 channels too
    if (vec_ok && spb % 4 == 0 && !getenv("AUKIT_DISABLE_TILED_ADPCM")) {
        const bool rec_aligned = ((uintptr_t)d_in % 16 == 0) && (bA % (2 * C) == 0);
        const int tc = (rec_aligned && (channels == 1 || channels == 2 || channels == 4 || channels == 8)) ? channels : 0;
        if (tc == 8 && bA % 16 == 0 && dialect != AUKIT_DIALECT_LITERAL && !getenv("AUKIT_DISABLE_STAGED_ADPCM")) {
            // per device, and a process may hold contexts on several: set on every launch (microseconds against a multi-ms kernel)
            if (aukit_cuda_check(cudaFuncSetAttribute(ms_adpcm_staged8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MS8_SMEM),
                                 "cudaFuncSetAttribute")) return -1;
            const unsigned g4 = aukit_grid(nblocks * C, threads, (size_t)ctx->num_sms * 4);
            ms_adpcm_staged8_kernel<<<g4, threads, MS8_SMEM, ctx->stream>>>(static_cast<const uint8_t *>(d_in), blockAlign, nblocks, spb, h,
                                                                          d_out, out_stride, ctx->d_status);
            ctx->launches++;
            return aukit_cuda_check(cudaGetLastError(), "ms_adpcm_staged8_kernel launch");
        }
#define AUKIT_MS_TILED(TC)                                                                                               \
    ms_adpcm_tiled_kernel<TC, 16><<<grid, threads, 0, ctx->stream>>>(static_cast<const uint8_t *>(d_in), blockAlign, channels, \
                                                                 dialect == AUKIT_DIALECT_LITERAL && channels == 1, nblocks, \
                                                                 spb, h, d_out, out_stride, ctx->d_status)
        switch (tc) {
            case 8: AUKIT_MS_TILED(8); break;
            case 4: AUKIT_MS_TILED(4); break;
            case 2: AUKIT_MS_TILED(2); break;
            case 1: AUKIT_MS_TILED(1); break;
            default: AUKIT_MS_TILED(0); break;
        }
#undef AUKIT_MS_TILED
        ctx->launches++;
        return aukit_cuda_check(cudaGetLastError(), "ms_adpcm_tiled_kernel launch");
    }
    ms_adpcm_kernel<<<grid, threads, 0, ctx->stream>>>(static_cast<const uint8_t *>(d_in), blockAlign, channels,
                                                       dialect == AUKIT_DIALECT_LITERAL && channels == 1, nblocks, spb,
                                                       h, d_out, out_stride,
                                                       ctx->d_status, vec_ok);
    ctx->launches++;
    return aukit_cuda_check(cudaGetLastError(), "ms_adpcm_kernel launch");
}

// used by capi.cu for aukit_cuda_adpcm (headerless nibble strings)
int aukit_launch_adpcm_stream(aukit_ctx *ctx, const uint8_t *d_in, size_t len, int channels, int topFirst,
                              int interleaved, const int *d_pred, const int *d_idx, float *d_out, size_t stride) {
    const int threads = 32;
    adpcm_stream_kernel<<<(channels + threads - 1) / threads, threads, 0, ctx->stream>>>(
        d_in, len, channels, topFirst, interleaved, d_pred, d_idx, d_out, stride);
    ctx->launches++;
    return aukit_cuda_check(cudaGetLastError(), "adpcm_stream_kernel launch");
}
