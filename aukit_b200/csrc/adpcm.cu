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

struct ms_coefs { int c1[256], c2[256]; int n; };

// One thread per (block, channel) chain.  Nibble stream after the 7*C-byte header is
// high-nibble-first; nibble m belongs to channel m % C (A:1317-1347 for C = 1, 2).
__global__ void __launch_bounds__(128)
ms_adpcm_kernel(const uint8_t *__restrict__ data, int blockAlign, int C, int literal_mono,
                size_t nblocks, size_t spb, const ms_coefs *__restrict__ coefs,
                float *__restrict__ out, size_t stride, int *status, int vec_ok) {
    __shared__ int adapt[16];
    if (threadIdx.x < 16) adapt[threadIdx.x] = c_ms_adapt[threadIdx.x];
    __syncthreads();
    const size_t nchains = nblocks * (size_t)C;
    const int ncoef = coefs->n;
    for (size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x; id < nchains;
         id += (size_t)gridDim.x * blockDim.x) {
        const size_t b = id / (size_t)C;
        const int c = (int)(id % (size_t)C);
        const size_t start = b * (size_t)blockAlign;
        const uint8_t *hp = data + (literal_mono ? 0 : start);         // A:1331: block 1's header
        int pi = hp[c];
        if (pi >= ncoef) { atomicOr(status, AUKIT_DEVERR_MS_PREDICTOR); pi = 0; }
        const int c1 = coefs->c1[pi], c2 = coefs->c2[pi];
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
    ima_wav_kernel<<<grid, threads, 0, ctx->stream>>>(static_cast<const uint8_t *>(d_in), nbytes, blockAlign, channels,
                                                      mode, nblocks, spb, groups, d_out, out_stride, ctx->d_status,
                                                      word_aligned, ((uintptr_t)d_out % 16 == 0) && (out_stride % 4 == 0));
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
    void *d_coefs = nullptr;
    if (aukit_upload_bytes(ctx, &h, sizeof h, &d_coefs)) return -1;
    const size_t nblocks = nbytes / bA;
    const size_t spb = 2 + (bA - 7 * C) * 2 / C;
    if (channels > 1 && out_stride < nblocks * spb) { aukit_dev_free(ctx, d_coefs); return aukit_fail("aukit_cuda: out_stride < frames"); }
    const int threads = 128;
    const unsigned grid = aukit_grid(nblocks * C, threads, (size_t)ctx->num_sms * 64);
    ms_adpcm_kernel<<<grid, threads, 0, ctx->stream>>>(static_cast<const uint8_t *>(d_in), blockAlign, channels,
                                                       dialect == AUKIT_DIALECT_LITERAL && channels == 1, nblocks, spb,
                                                       static_cast<const ms_coefs *>(d_coefs), d_out, out_stride,
                                                       ctx->d_status, ((uintptr_t)d_out % 16 == 0) && (out_stride % 4 == 0));
    ctx->launches++;
    int rc = aukit_cuda_check(cudaGetLastError(), "ms_adpcm_kernel launch");
    aukit_dev_free(ctx, d_coefs);
    return rc;
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
