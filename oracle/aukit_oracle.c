/*
 * aukit_oracle.c -- literal CPU restatement of AUKit's preload path.
 * TEST INFRASTRUCTURE ONLY (see aukit_oracle.h).  "A:n" = /root/reference/aukit.lua line n.
 *
 * Build with -ffp-contract=off: the reference evaluates every expression as separately
 * rounded IEEE-754 double operations (Lua 5.2 numbers), so no FMA contraction is allowed.
 */
#include "aukit_oracle.h"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static __thread char g_err[256];

const char *auko_last_error(void) { return g_err; }

static int fail(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return -1;
}

void auko_free(void *p) { free(p); }

/* clamp, A:228-232: three-way compare; NaN passes through unchanged. */
static double clampd(double n, double lo, double hi) {
    if (n < lo) return lo;
    else if (n > hi) return hi;
    else return n;
}

/* ---------------------------------------------------------------- tables, A:156-176 */
static const int ima_index_table[16] = {-1, -1, -1, -1, 2, 4, 6, 8, -1, -1, -1, -1, 2, 4, 6, 8};
static const int ima_step_table[89] = {
    7,     8,     9,     10,    11,    12,    13,    14,    16,    17,    19,    21,    23,
    25,    28,    31,    34,    37,    41,    45,    50,    55,    60,    66,    73,    80,
    88,    97,    107,   118,   130,   143,   157,   173,   190,   209,   230,   253,   279,
    307,   337,   371,   408,   449,   494,   544,   598,   658,   724,   796,   876,   963,
    1060,  1166,  1282,  1411,  1552,  1707,  1878,  2066,  2272,  2499,  2749,  3024,  3327,
    3660,  4026,  4428,  4871,  5358,  5894,  6484,  7132,  7845,  8630,  9493,  10442, 11487,
    12635, 13899, 15289, 16818, 18500, 20350, 22385, 24623, 27086, 29794, 32767};
/* msadpcm_adaption_table indexed by SIGNED nibble -8..7 (A:173-176); stored at [nib+8]. */
static const int ms_adapt[16] = {768, 614, 512, 409, 307, 230, 230, 230,
                                 230, 230, 230, 230, 307, 409, 512, 614};
static const int ms_coef1_default[7] = {256, 512, 0, 192, 240, 460, 392};   /* A:1304 */
static const int ms_coef2_default[7] = {0, -256, 0, 64, 0, -208, -232};

/* ---------------------------------------------------------------- aukit.pcm, A:1049-1171 */
static double pcm_scale_signed(double s, double maxValue) {                  /* A:1133 */
    return s / (s < 0 ? maxValue : maxValue - 1);
}
static double pcm_scale_unsigned(double s, double maxValue) {                /* A:1152 */
    return (s - 128) / (s < 128 ? maxValue : maxValue - 1);  /* literal 128 for every depth */
}

/* string.unpack of one sample: i1..i4 / I1..I4 / f, either endianness (A:1067-1070). */
static double pcm_read(const uint8_t *p, int byteDepth, int dataType, int bigEndian,
                       double maxValue) {
    uint32_t u = 0;
    if (bigEndian) for (int k = 0; k < byteDepth; k++) u = (u << 8) | p[k];
    else for (int k = byteDepth - 1; k >= 0; k--) u = (u << 8) | p[k];
    if (dataType == AUKO_FLOAT) {
        float f;
        memcpy(&f, &u, 4);
        return (double)f;                                                    /* A:1112-1114 */
    }
    if (dataType == AUKO_SIGNED) {
        int64_t s = (int64_t)u;
        if (s >= ((int64_t)1 << (8 * byteDepth - 1))) s -= (int64_t)1 << (8 * byteDepth);
        return pcm_scale_signed((double)s, maxValue);
    }
    return pcm_scale_unsigned((double)u, maxValue);
}

int auko_pcm(const uint8_t *data, size_t nbytes, int bitDepth, int dataType, int channels,
             int interleaved, int bigEndian, double *out, size_t stride, size_t *len_out) {
    g_err[0] = 0;
    if (bitDepth != 8 && bitDepth != 16 && bitDepth != 24 && bitDepth != 32)
        return fail("bad argument #2 (invalid bit depth)");                  /* A:1058 */
    if (dataType != AUKO_SIGNED && dataType != AUKO_UNSIGNED && dataType != AUKO_FLOAT)
        return fail("bad argument #3 (invalid data type)");                  /* A:1059 */
    if (dataType == AUKO_FLOAT && bitDepth != 32)
        return fail("bad argument #2 (float audio must have 32-bit depth)"); /* A:1060 */
    if (channels < 1) return fail("number outside of range (expected %d to be at least 1)", channels);
    int byteDepth = bitDepth / 8;
    /* A:1064: (#data / byteDepth) % channels ~= 0, evaluated in doubles */
    double q = (double)nbytes / (double)byteDepth;
    if (fmod(q, (double)channels) != 0.0)
        return fail("bad argument #1 (uneven amount of data per channel)");
    size_t len = nbytes / (size_t)byteDepth / (size_t)channels;              /* A:1065 */
    double maxValue = ldexp(1.0, bitDepth - 1);                              /* A:1071 */
    if (len_out) *len_out = len;
    if (len > stride && channels > 1) return fail("oracle: stride too small");
    if (interleaved && channels > 1) {                                       /* A:1156-1161 */
        const uint8_t *p = data;
        for (size_t i = 0; i < len; i++)
            for (int j = 0; j < channels; j++, p += byteDepth)
                out[(size_t)j * stride + i] = pcm_read(p, byteDepth, dataType, bigEndian, maxValue);
    } else {                                                                 /* A:1162-1169 */
        const uint8_t *p = data;
        for (int j = 0; j < channels; j++)
            for (size_t i = 0; i < len; i++, p += byteDepth)
                out[(size_t)j * stride + i] = pcm_read(p, byteDepth, dataType, bigEndian, maxValue);
    }
    return 0;
}

/* ---------------------------------------------------------------- aukit.g711, A:1361-1384 */
int auko_g711(const uint8_t *data, size_t nbytes, int ulaw, int channels, double *out,
              size_t stride, size_t *lens) {
    g_err[0] = 0;
    if (channels < 1) return fail("oracle: channels < 1");
    uint32_t xr = ulaw ? 0xFF : 0x55;                                        /* A:1368 */
    for (size_t k = 0; k < nbytes; k++) {
        uint32_t b = data[k] ^ xr;                                           /* A:1374 */
        uint32_t m = b & 0x0F, e = (b >> 4) & 7;                             /* A:1375 */
        if (!ulaw && e == 0) m = m * 4 + 2;                                  /* A:1376 */
        else m = (m * 2 + 33) << e;                                          /* A:1377 */
        double md = (double)m;
        if (ulaw) md = md - 33;                                              /* A:1378 */
        int btest = (b & 0x80) != 0;
        double div = (btest == (ulaw != 0)) ? -8192.0 : 8192.0;              /* A:1379 */
        size_t c = k % (size_t)channels, i = k / (size_t)channels;
        if (i >= stride) return fail("oracle: stride too small");
        out[c * stride + i] = md / div;
    }
    if (lens)
        for (int c = 0; c < channels; c++)
            lens[c] = nbytes > (size_t)c ? (nbytes - (size_t)c + (size_t)channels - 1) / (size_t)channels : 0;
    return 0;
}

/* ---------------------------------------------------------------- IMA, A:1183-1274 */
double auko_ima_step(int nibble, int *pred, int *idx) {
    int step = ima_step_table[*idx];                                         /* A:1250 */
    int ni = *idx + ima_index_table[nibble];                                 /* A:1251 */
    *idx = ni < 0 ? 0 : (ni > 88 ? 88 : ni);
    uint32_t diff = (((uint32_t)(nibble % 8) * (uint32_t)step) >> 2) + ((uint32_t)step >> 3); /* A:1252 */
    int p = *pred;
    if (nibble >= 8) p -= (int)diff; else p += (int)diff;                    /* A:1253-1254 */
    if (p < -32768) p = -32768;
    if (p > 32767) p = 32767;
    *pred = p;
    return (double)p / (p < 0 ? 32768.0 : 32767.0);                          /* A:1255 */
}

/* read() closure of aukit.adpcm for string input (A:1216-1230) */
typedef struct { const uint8_t *d; size_t n, pos; int have_tmp, tmp, topFirst; } nibreader;
static int nib_read(nibreader *r, int *err) {
    if (r->have_tmp) { r->have_tmp = 0; return r->tmp; }
    if (r->pos >= r->n) { *err = 1; return 0; }
    int b = r->d[r->pos++];
    int first;
    if (r->topFirst) { r->tmp = b & 0x0F; first = b >> 4; }
    else { r->tmp = b >> 4; first = b & 0x0F; }
    r->have_tmp = 1;
    return first;
}

int auko_adpcm(const uint8_t *data, size_t nbytes, int channels, int topFirst, int interleaved,
               const int *predictor, const int *step_index, double *out, size_t stride,
               size_t *len_out) {
    g_err[0] = 0;
    if (channels < 1 || channels > 64) return fail("oracle: bad channel count");
    int pred[64], idx[64];
    for (int j = 0; j < channels; j++) {
        pred[j] = predictor ? predictor[j] : 0;
        idx[j] = step_index ? step_index[j] : 0;
        if (pred[j] < -32768 || pred[j] > 32767)
            return fail("number outside of range (expected %d to be within -32768 and 32767)", pred[j]);
        if (idx[j] < 0 || idx[j] > 88)
            return fail("number outside of range (expected %d to be within 0 and 88)", idx[j]);
    }
    size_t len = (nbytes * 2) / (size_t)channels;                            /* A:1231 */
    if (len_out) *len_out = len;
    if (len > stride && channels > 1) return fail("oracle: stride too small");
    nibreader r = {data, nbytes, 0, 0, 0, topFirst};
    int err = 0;
    if (interleaved) {                                                       /* A:1243-1258 */
        for (size_t i = 0; i < len; i++)
            for (int j = 0; j < channels; j++)
                out[(size_t)j * stride + i] = auko_ima_step(nib_read(&r, &err), &pred[j], &idx[j]);
    } else {                                                                 /* A:1259-1272 */
        for (int j = 0; j < channels; j++)
            for (size_t i = 0; i < len; i++)
                out[(size_t)j * stride + i] = auko_ima_step(nib_read(&r, &err), &pred[j], &idx[j]);
    }
    if (err) return fail("oracle: nibble read past end");
    return 0;
}

static int rd_i16(const uint8_t *p) { return (int16_t)(p[0] | (p[1] << 8)); }
static int rd_u16(const uint8_t *p) { return p[0] | (p[1] << 8); }
static uint32_t rd_u32(const uint8_t *p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

/* Number of 8-byte groups the literal stereo loop "for i = 8, blockAlign-1, 8" runs. */
static size_t ima_stereo_groups(int blockAlign) {
    return blockAlign > 8 ? (size_t)(blockAlign - 1 - 8) / 8 + 1 : 0;
}

size_t auko_wav_ima_len(size_t nbytes, int blockAlign, int channels, int dialect) {
    if (blockAlign < 1 || channels < 1) return 0;
    size_t nblocks = (nbytes + (size_t)blockAlign - 1) / (size_t)blockAlign;
    if (dialect == AUKO_DIALECT_LITERAL && channels == 2)
        return nblocks * ima_stereo_groups(blockAlign) * 8;
    if (dialect == AUKO_DIALECT_LITERAL && channels == 1) {
        size_t total = 0;
        for (size_t off = 0; off < nbytes; off += (size_t)blockAlign) {
            size_t end = off + (size_t)blockAlign < nbytes ? off + (size_t)blockAlign : nbytes;
            total += end > off + 4 ? (end - off - 4) * 2 : 0;
        }
        return total;
    }
    /* general: full blocks only; groups of 4*C bytes after a 4*C header */
    size_t full = nbytes / (size_t)blockAlign;
    size_t hdr = 4 * (size_t)channels;
    size_t groups = (size_t)blockAlign > hdr ? ((size_t)blockAlign - hdr) / hdr : 0;
    return full * groups * 8;
}

int auko_wav_ima(const uint8_t *data, size_t nbytes, int blockAlign, int channels, int dialect,
                 double *out, size_t stride, size_t *len_out) {
    g_err[0] = 0;
    if (blockAlign < 1) return fail("'for' step must be positive");
    if (nbytes == 0) return fail("attempt to index a nil value (no blocks)");   /* A:1548 */
    size_t pos = 0; /* samples written per channel so far (concat, A:1548) */
    if (dialect == AUKO_DIALECT_LITERAL && channels == 2) {                  /* A:1512-1541 */
        size_t groups = ima_stereo_groups(blockAlign);
        for (size_t off = 0; off < nbytes; off += (size_t)blockAlign) {
            if (off + 7 > nbytes) return fail("data string too short");
            int pred[2] = {rd_i16(data + off), rd_i16(data + off + 4)};      /* <hBxhB */
            int idx[2] = {data[off + 2], data[off + 6]};
            for (int c = 0; c < 2; c++)
                if (idx[c] > 88)
                    return fail("number outside of range (expected %d to be within 0 and 88)", idx[c]);
            /* nibble table -> interleaved L,R decode; equivalent per-channel order:
             * group g: bytes off+8+8g .. +3 = L (low nibble first), +4..+7 = R */
            for (size_t g = 0; g < groups; g++) {
                size_t b0 = off + 8 + 8 * g;
                if (b0 + 8 > nbytes) return fail("bad argument #1 to 'band' (number expected, got nil)");
                if (pos + 8 > stride) return fail("oracle: stride too small");
                /* order of state updates in the reference is L1,R1,L2,R2,...; channels are
                 * independent so per-channel order is what matters */
                for (int k = 0; k < 8; k++) {
                    int nl = (data[b0 + (size_t)(k >> 1)] >> ((k & 1) * 4)) & 0x0F;
                    int nr = (data[b0 + 4 + (size_t)(k >> 1)] >> ((k & 1) * 4)) & 0x0F;
                    out[0 * stride + pos + (size_t)k] = auko_ima_step(nl, &pred[0], &idx[0]);
                    out[1 * stride + pos + (size_t)k] = auko_ima_step(nr, &pred[1], &idx[1]);
                }
                pos += 8;
            }
        }
    } else if (dialect == AUKO_DIALECT_LITERAL) {                            /* A:1542-1546 */
        if (channels != 1) return fail("bad argument #6 (table too short)"); /* A:1199 */
        for (size_t off = 0; off < nbytes; off += (size_t)blockAlign) {
            if (off + 3 > nbytes) return fail("data string too short");
            int pred = rd_i16(data + off);                                   /* <hB */
            int idx = data[off + 2] & 0x0F;                                  /* A:1544 (bug) */
            size_t end = off + (size_t)blockAlign < nbytes ? off + (size_t)blockAlign : nbytes;
            for (size_t b = off + 4; b < end; b++) {                         /* topFirst=false */
                if (pos + 2 > stride) return fail("oracle: stride too small");
                out[pos++] = auko_ima_step(data[b] & 0x0F, &pred, &idx);
                out[pos++] = auko_ima_step(data[b] >> 4, &pred, &idx);
            }
        }
    } else {                                                                 /* A:2798-2815 layout */
        if (channels < 1 || channels > 64) return fail("oracle: bad channel count");
        size_t hdr = 4 * (size_t)channels;
        size_t groups = (size_t)blockAlign > hdr ? ((size_t)blockAlign - hdr) / hdr : 0;
        if (nbytes % (size_t)blockAlign) return fail("oracle: general dialect needs whole blocks");
        for (size_t off = 0; off < nbytes; off += (size_t)blockAlign) {
            int pred[64], idx[64];
            for (int c = 0; c < channels; c++) {
                pred[c] = rd_i16(data + off + 4 * (size_t)c);                /* A:2799 */
                idx[c] = data[off + 4 * (size_t)c + 2];
                if (idx[c] > 88) return fail("attempt to perform arithmetic on a nil value (step index %d)", idx[c]);
            }
            for (size_t g = 0; g < groups; g++) {
                if (pos + 8 > stride) return fail("oracle: stride too small");
                for (int c = 0; c < channels; c++) {
                    uint32_t num = rd_u32(data + off + hdr + g * hdr + 4 * (size_t)c);  /* A:2804 */
                    for (int k = 0; k < 8; k++)
                        out[(size_t)c * stride + pos + (size_t)k] =
                            auko_ima_step((int)((num >> (4 * k)) & 0xF), &pred[c], &idx[c]);
                }
                pos += 8;
            }
        }
    }
    if (len_out) *len_out = pos;
    return 0;
}

/* ---------------------------------------------------------------- MS-ADPCM, A:1283-1353 */
typedef struct { double s1, s2, delta; int c1, c2; } ms_state;

/* math.max(a, b) of Lua 5.2 for two args: starts at a, takes b if b > a (NaN a stays). */
static double lua_max2(double a, double b) { return b > a ? b : a; }

static double ms_emit(double p) { return p / (p < 0 ? 32768.0 : 32767.0); }

static double ms_step(ms_state *st, int nib /* signed -8..7 */) {
    /* A:1321-1324, all in doubles like the reference */
    double predictor = clampd(floor((st->s1 * st->c1 + st->s2 * st->c2) / 256.0) + nib * st->delta,
                              -32768, 32767);
    st->s2 = st->s1;
    st->s1 = predictor;
    double nd = floor(ms_adapt[nib + 8] * st->delta / 256.0);
    st->delta = lua_max2(nd, 16);                                            /* A:1324 */
    return ms_emit(predictor);
}

static int sgn4(int n) { return n >= 8 ? n - 16 : n; }                       /* A:1319-1320 */

size_t auko_msadpcm_len(size_t nbytes, int blockAlign, int channels) {
    if (blockAlign < 1 || channels < 1) return 0;
    size_t nblocks = (nbytes + (size_t)blockAlign - 1) / (size_t)blockAlign;
    size_t hdr = 7 * (size_t)channels;
    size_t body = (size_t)blockAlign > hdr ? (size_t)blockAlign - hdr : 0;
    return nblocks * (2 + body * 2 / (size_t)channels);
}

int auko_msadpcm(const uint8_t *data, size_t nbytes, int blockAlign, int channels,
                 const int *coef1, const int *coef2, int ncoef, int dialect, double *out,
                 size_t stride, size_t *len_out) {
    g_err[0] = 0;
    if (blockAlign < 1) return fail("'for' step must be positive");
    if (!coef1 || !coef2 || ncoef <= 0) { coef1 = ms_coef1_default; coef2 = ms_coef2_default; ncoef = 7; }
    size_t pos = 0;
    if (dialect == AUKO_DIALECT_LITERAL && channels != 1 && channels != 2)
        return nbytes ? fail("Unsupported number of channels: %d", channels) : (len_out ? (*len_out = 0, 0) : 0);
    if (channels < 1 || channels > 64) return fail("oracle: bad channel count");
    size_t hdr = 7 * (size_t)channels;
    for (size_t off = 0; off < nbytes; off += (size_t)blockAlign) {
        ms_state st[64];
        int C = channels;
        if (dialect == AUKO_DIALECT_LITERAL && C == 1) {
            /* A:1331: str_unpack("<!1Bhhh", data) -- no position: ALWAYS block 1's header */
            if (nbytes < 7) return fail("data string too short");
            int pi = data[0];
            if (pi >= ncoef) return fail("attempt to perform arithmetic on a nil value (coefficient %d)", pi);
            st[0].c1 = coef1[pi]; st[0].c2 = coef2[pi];
            st[0].delta = rd_i16(data + 1); st[0].s1 = rd_i16(data + 3); st[0].s2 = rd_i16(data + 5);
        } else {
            /* stereo literal A:1310 "<BBhhhhhh" == general layout with C = 2 */
            if (off + hdr > nbytes) return fail("data string too short");
            for (int c = 0; c < C; c++) {
                int pi = data[off + (size_t)c];
                if (pi >= ncoef) return fail("attempt to perform arithmetic on a nil value (coefficient %d)", pi);
                st[c].c1 = coef1[pi]; st[c].c2 = coef2[pi];
                st[c].delta = rd_i16(data + off + (size_t)C + 2 * (size_t)c);
                st[c].s1 = rd_i16(data + off + 3 * (size_t)C + 2 * (size_t)c);
                st[c].s2 = rd_i16(data + off + 5 * (size_t)C + 2 * (size_t)c);
            }
        }
        if (pos + 2 > stride) return fail("oracle: stride too small");
        for (int c = 0; c < C; c++) {                                        /* A:1312-1315 */
            out[(size_t)c * stride + pos] = ms_emit(st[c].s2);
            out[(size_t)c * stride + pos + 1] = ms_emit(st[c].s1);
        }
        pos += 2;
        /* nibble stream, high nibble first, channel = nibble index mod C (A:1317-1347) */
        size_t nn = 0;
        for (size_t i = hdr; i < (size_t)blockAlign; i++) {
            if (off + i >= nbytes) return fail("bad argument #1 to 'rshift' (number expected, got nil)");
            int b = data[off + i];
            int nibs[2] = {sgn4(b >> 4), sgn4(b & 0x0F)};
            for (int h = 0; h < 2; h++, nn++) {
                int c = (int)(nn % (size_t)C);
                size_t k = nn / (size_t)C;
                if (pos + k >= stride) return fail("oracle: stride too small");
                out[(size_t)c * stride + pos + k] = ms_step(&st[c], nibs[h]);
            }
        }
        pos += nn / (size_t)C;
    }
    if (len_out) *len_out = pos;
    return 0;
}

/* ---------------------------------------------------------------- resample, A:253-266, A:653-673 */
/* 1-based table access with nil outside [1, n]: returns 0 and sets *nil. */
static double tget(const double *d, size_t n, double idx, int *nil) {
    if (idx < 1 || idx > (double)n) { *nil = 1; return 0; }
    *nil = 0;
    return d[(size_t)idx - 1];
}

static int interp_eval(int mode, const double *d, size_t n, double x, double *res) {
    double ffx = floor(x);
    int nil;
    if (mode == AUKO_INTERP_NONE) {                                          /* A:254-256 */
        *res = tget(d, n, ffx, &nil);
        return nil ? -1 : 0;
    }
    if (mode == AUKO_INTERP_LINEAR) {                                        /* A:257-260 */
        double a = tget(d, n, ffx, &nil);
        if (nil) return -1;
        double b = tget(d, n, ffx + 1, &nil);
        if (nil) b = a;
        *res = a + (b - a) * (x - ffx);
        return 0;
    }
    if (mode == AUKO_INTERP_SINC) {                                          /* A:267-281, sincWindowSize = 10 (A:129) */
        double fx = x - ffx, sum = 0;
        for (int k = -10; k <= 10; k++) {
            double dv = tget(d, n, ffx + k, &nil);
            if (!nil) {
                double px = 3.14159265358979323846 * (fx - k);
                if (px == 0) sum = sum + dv;
                else sum = sum + dv * sin(px) / px;
            }
        }
        *res = sum;
        return 0;
    }
    /* cubic, A:261-266 */
    int n0, n1, n2, n3;
    double p0 = tget(d, n, ffx - 1, &n0), p1 = tget(d, n, ffx, &n1);
    double p2 = tget(d, n, ffx + 1, &n2), p3 = tget(d, n, ffx + 2, &n3);
    double fx = x - ffx;
    if (n1) return -1;
    if (n0) p0 = p1;
    if (n2) p2 = p1;
    if (n3) p3 = p2; /* p3 or p2 or p1, with p2 already substituted */
    *res = (-0.5 * p0 + 1.5 * p1 - 1.5 * p2 + 0.5 * p3) * pow(fx, 3) +
           (p0 - 2.5 * p1 + 2 * p2 - 0.5 * p3) * pow(fx, 2) + (-0.5 * p0 + 0.5 * p2) * fx + p1;
    return 0;
}

size_t auko_resample_len(size_t n_in, double srcRate, double dstRate) {
    double ratio = dstRate / srcRate;                                        /* A:658 */
    double newlen = (double)n_in * ratio;                                    /* A:659 */
    if (!(newlen >= 1)) return 0;
    return (size_t)floor(newlen);                                            /* for i = 1, newlen */
}

double auko_resample_pos(uint64_t i, double srcRate, double dstRate) {
    double ratio = dstRate / srcRate;
    return ((double)i - 1) / ratio + 1;                                      /* A:666 */
}

int auko_resample(const double *in, size_t in_stride, int channels, size_t n_in, double srcRate,
                  double dstRate, int interp, double *out, size_t out_stride, size_t *n_out) {
    g_err[0] = 0;
    if (interp < 0 || interp > 3) return fail("bad argument #2 (invalid interpolation type)");
    double ratio = dstRate / srcRate;
    size_t newlen = auko_resample_len(n_in, srcRate, dstRate);
    if (n_out) *n_out = newlen;
    if (newlen > out_stride && channels > 1) return fail("oracle: stride too small");
    for (int y = 0; y < channels; y++) {
        const double *c = in + (size_t)y * in_stride;
        double *line = out + (size_t)y * out_stride;
        for (size_t i = 1; i <= newlen; i++) {
            double x = ((double)i - 1) / ratio + 1;                          /* A:666 */
            if (x - floor(x / 1) * 1 == 0) {                                 /* x % 1 == 0, A:667 */
                int nil;
                line[i - 1] = tget(c, n_in, x, &nil);
                if (nil) return fail("oracle: exact-hit index %g out of range (nil hole)", x);
            } else {
                double v;
                if (interp_eval(interp, c, n_in, x, &v)) return fail("attempt to perform arithmetic on a nil value");
                line[i - 1] = clampd(v, -1, 1);                              /* A:668 */
            }
        }
    }
    return 0;
}

/* ---------------------------------------------------------------- mono, A:677-689 */
int auko_mono(const double *in, size_t in_stride, int channels, size_t n, double *out) {
    g_err[0] = 0;
    for (size_t i = 0; i < n; i++) {
        double s = 0;
        for (int c = 0; c < channels; c++) s = s + in[(size_t)c * in_stride + i];
        out[i] = s / channels;
    }
    return 0;
}

/* ---------------------------------------------------------------- effects */
int auko_amplify(double *d, size_t stride, int channels, size_t n, double multiplier) {
    g_err[0] = 0;
    if (multiplier == 1) return 0;                                           /* A:3359 */
    for (int c = 0; c < channels; c++)
        for (size_t i = 0; i < n; i++)
            d[(size_t)c * stride + i] = clampd(d[(size_t)c * stride + i] * multiplier, -1, 1);
    return 0;
}

int auko_normalize(double *d, size_t stride, int channels, size_t n, double peak,
                   int independent) {
    g_err[0] = 0;
    double mult = 0;
    if (!independent) {                                                      /* A:3437-3445 */
        double mx = 0;
        for (int c = 0; c < channels; c++)
            for (size_t i = 0; i < n; i++) mx = lua_max2(mx, fabs(d[(size_t)c * stride + i]));
        mult = peak / mx;
    }
    for (int c = 0; c < channels; c++) {
        double *ch = d + (size_t)c * stride;
        if (independent) {                                                   /* A:3448-3452 */
            double mx = 0;
            for (size_t i = 0; i < n; i++) mx = lua_max2(mx, fabs(ch[i]));
            mult = peak / mx;
        }
        for (size_t i = 0; i < n; i++) ch[i] = clampd(ch[i] * mult, -1, 1);  /* A:3455 */
    }
    return 0;
}

double auko_encode_pcm(double d, int bitDepth, int dataType) {               /* A:869-874 */
    double maxValue = ldexp(1.0, bitDepth - 1);
    double add = dataType == AUKO_UNSIGNED ? maxValue : 0;
    if (dataType == AUKO_FLOAT) return d;
    return d * (d < 0 ? maxValue : maxValue - 1) + add;
}

/* ---------------------------------------------------------------- more in-place effects */
/* table read with a float key: a non-integer key has no entry */
static double tgetx(const double *d, size_t n, double idx, int *nil) {
    if (idx != floor(idx)) { *nil = 1; return 0; }
    return tget(d, n, idx, nil);
}

int auko_invert(double *d, size_t stride, int channels, size_t n) {          /* A:3412-3419 */
    g_err[0] = 0;
    for (int c = 0; c < channels; c++)
        for (size_t i = 0; i < n; i++) d[(size_t)c * stride + i] = -d[(size_t)c * stride + i];
    return 0;
}

int auko_fade(double *d, size_t stride, int channels, size_t n, double sampleRate, double startTime,
              double startAmplitude, double endTime, double endAmplitude) {  /* A:3392-3410 */
    g_err[0] = 0;
    if (startAmplitude == 1 && endAmplitude == 1) return 0;
    for (int c = 0; c < channels; c++) {
        double *ch = d + (size_t)c * stride;
        double start = startTime * sampleRate;
        double m = (endAmplitude - startAmplitude) / ((endTime - startTime) * sampleRate);
        for (double i = start; i <= endTime * sampleRate; i = i + 1) {
            int nil;
            double v = tgetx(ch, n, i, &nil);                                 /* non-integer or out-of-range key: nil */
            if (nil) return fail("attempt to perform arithmetic on a nil value (field '?')");
            ch[(size_t)i - 1] = clampd(v * (m * (i - start) + startAmplitude), -1, 1);
        }
    }
    return 0;
}

int auko_delay(double *d, size_t stride, int channels, size_t n, double sampleRate, double delay,
               double multiplier) {                                          /* A:3500-3513 */
    g_err[0] = 0;
    double samples = floor(delay * sampleRate);
    double *original = (double *)malloc(sizeof(double) * (n ? n : 1));
    for (int c = 0; c < channels; c++) {
        double *o = d + (size_t)c * stride;
        for (size_t i = 0; i < n; i++) original[i] = o[i];
        for (double i = samples + 1; i <= (double)n; i = i + 1) {
            int nil1, nil2;
            double a = tgetx(o, n, i, &nil1), b = tgetx(original, n, i - samples, &nil2);
            if (nil1 || nil2) { free(original); return fail("attempt to perform arithmetic on a nil value (field '?')"); }
            o[(size_t)i - 1] = clampd(a + b * multiplier, -1, 1);
        }
    }
    free(original);
    return 0;
}

int auko_center(double *d, size_t stride, int channels, size_t n, double sampleRate) {   /* A:3465-3478 */
    g_err[0] = 0;
    if (!(sampleRate > 0)) return fail("'for' step must be positive");
    for (int c = 0; c < channels; c++) {
        double *ch = d + (size_t)c * stride;
        for (double i = 0; i <= (double)n - 1; i = i + sampleRate) {
            double avg = 0;
            double l = (double)n - i < sampleRate ? (double)n - i : sampleRate;
            for (double j = 1; j <= l; j = j + 1) {
                int nil;
                double v = tgetx(ch, n, i + j, &nil);
                if (nil) return fail("attempt to perform arithmetic on a nil value (field '?')");
                avg = avg + v;
            }
            avg = avg / l;
            for (double j = 1; j <= l; j = j + 1) ch[(size_t)(i + j) - 1] = clampd(ch[(size_t)(i + j) - 1] - avg, -1, 1);
        }
    }
    return 0;
}

/* Audio:pcm(bitDepth, dataType, interleaved), A:901-911 -> encodePCM(info, 1), A:868-894.
 * out holds channels*n numbers: interleaved data[(n-1)*nc+c] (A:883), else data[(c-1)*len+n] (A:894). */
int auko_audio_pcm(const double *d, size_t stride, int channels, size_t n, int bitDepth, int dataType,
                   int interleaved, double *out) {
    g_err[0] = 0;
    if (bitDepth != 8 && bitDepth != 16 && bitDepth != 24 && bitDepth != 32)
        return fail("bad argument #2 (invalid bit depth)");                                  /* A:907 */
    if (dataType != AUKO_SIGNED && dataType != AUKO_UNSIGNED && dataType != AUKO_FLOAT)
        return fail("bad argument #3 (invalid data type)");                                  /* A:908 */
    if (dataType == AUKO_FLOAT && bitDepth != 32)
        return fail("bad argument #2 (float audio must have 32-bit depth)");                 /* A:909 */
    for (int c = 0; c < channels; c++)
        for (size_t i = 0; i < n; i++) {
            double v = auko_encode_pcm(d[(size_t)c * stride + i], bitDepth, dataType);
            if (interleaved) out[i * (size_t)channels + (size_t)c] = v;
            else out[(size_t)c * n + i] = v;
        }
    return 0;
}

int auko_lowpass(double *d, size_t stride, int channels, size_t n, double frequency,
                 double sampleRate) {                                        /* A:3586-3598 */
    g_err[0] = 0;
    double a = 1 - exp(-(frequency / sampleRate) * 2 * 3.14159265358979323846);
    for (int c = 0; c < channels; c++) {
        double *ch = d + (size_t)c * stride;
        for (size_t i = 1; i < n; i++) {
            double l = ch[i - 1];
            ch[i] = l + a * (ch[i] - l);
        }
    }
    return 0;
}

int auko_highpass(double *d, size_t stride, int channels, size_t n, double frequency,
                  double sampleRate) {                                       /* A:3605-3618 */
    g_err[0] = 0;
    double a = 1 / (2 * 3.14159265358979323846 * (frequency / sampleRate) + 1);
    for (int c = 0; c < channels; c++) {
        double *ch = d + (size_t)c * stride;
        if (n == 0) continue;
        double lx = ch[0];
        for (size_t i = 1; i < n; i++) {
            double llx = ch[i];
            ch[i] = a * (ch[i - 1] + llx - lx);
            lx = llx;
        }
    }
    return 0;
}

/* ---------------------------------------------------------------- aukit.wav, A:1456-1574 */
static const uint8_t guid_tail[12] = {0x00, 0x00, 0x10, 0x00, 0x80, 0x00,
                                      0x00, 0xaa, 0x00, 0x38, 0x9b, 0x71};   /* A:133-139 */
static const uint8_t guid_dfpwm[16] = {0x3a, 0xc1, 0xfa, 0x38, 0x81, 0x1d, 0x43, 0x61,
                                       0xa4, 0x0d, 0xce, 0x53, 0xca, 0x60, 0x7c, 0xd1}; /* A:125 */

int auko_wav_parse(const uint8_t *data, size_t nbytes, auko_wav_info *info) {
    g_err[0] = 0;
    memset(info, 0, sizeof *info);
    info->format = AUKO_WAV_NONE;
    /* every data chunk is decoded where it is met, with the fmt state seen so far (A:1505-1555); `coefficients`
     * and `dataType` are locals of the whole walk (A:1458), so they survive a later fmt chunk that does not set them */
    static __thread auko_wav_info cur;
    memset(&cur, 0, sizeof cur);
    cur.format = AUKO_WAV_NONE;
    if (nbytes < 4) return fail("data string too short");
    if (memcmp(data, "RIFF", 4)) return fail("bad argument #1 (not a WAV file)");   /* A:1460 */
    if (nbytes < 12) return fail("data string too short");
    if (memcmp(data + 8, "WAVE", 4)) return fail("bad argument #1 (not a WAV file)"); /* A:1463 */
    size_t pos = 12; /* 0-based */
    while (pos < nbytes) {                                                   /* pos <= #data */
        if (pos + 8 > nbytes) return fail("data string too short");
        const uint8_t *id = data + pos;
        size_t size = rd_u32(data + pos + 4);
        pos += 8;
        if (!memcmp(id, "fmt ", 4)) {
            size_t clen = pos + size <= nbytes ? size : (pos < nbytes ? nbytes - pos : 0);
            const uint8_t *ch = data + pos;
            pos += size;
            if (clen < 16) return fail("data string too short");
            int format = rd_u16(ch);                                         /* A:1473 */
            cur.channels = rd_u16(ch + 2);
            cur.sampleRate = (int)rd_u32(ch + 4);
            cur.blockAlign = rd_u16(ch + 12);
            cur.bitDepth = rd_u16(ch + 14);
            cur.have_fmt = 1;
            if (format == 1) cur.format = cur.bitDepth == 8 ? AUKO_WAV_PCM_UNSIGNED : AUKO_WAV_PCM_SIGNED;
            else if (format == 2) {
                cur.format = AUKO_WAV_MSADPCM;
                if (clen < 22) return fail("data string too short");
                int numcoeff = rd_u16(ch + 20);                              /* A:1478 */
                if (numcoeff > 0) {
                    if (numcoeff > 256) return fail("oracle: too many coefficients");
                    for (int i = 1; i <= numcoeff; i++) {                    /* A:1481-1483 */
                        size_t o = (size_t)i * 4 + 18;
                        if (o + 4 > clen) return fail("data string too short");
                        cur.coef1[i - 1] = rd_i16(ch + o);
                        cur.coef2[i - 1] = rd_i16(ch + o + 2);
                    }
                    cur.ncoef = numcoeff;
                }
            } else if (format == 3) cur.format = AUKO_WAV_FLOAT;
            else if (format == 6) cur.format = AUKO_WAV_ALAW;
            else if (format == 7) cur.format = AUKO_WAV_ULAW;
            else if (format == 0x11) cur.format = AUKO_WAV_ADPCM;
            else if (format == 0xFFFE) {                                     /* A:1493-1503 */
                if (clen < 20) return fail("data string too short");
                cur.bitDepth = rd_u16(ch + 18);
                uint8_t uuid[16] = {0};
                size_t have = clen > 24 ? (clen - 24 < 16 ? clen - 24 : 16) : 0;
                memcpy(uuid, ch + 24, have);
                if (have == 16 && !memcmp(uuid, guid_dfpwm, 16)) cur.format = AUKO_WAV_DFPWM;
                else if (have == 16 && !memcmp(uuid + 4, guid_tail, 12) && uuid[1] == 0 && uuid[2] == 0 && uuid[3] == 0) {
                    switch (uuid[0]) {
                    case 0x01: cur.format = cur.bitDepth == 8 ? AUKO_WAV_PCM_UNSIGNED : AUKO_WAV_PCM_SIGNED; break;
                    case 0x02: cur.format = AUKO_WAV_MSADPCM; break;
                    case 0x03: cur.format = AUKO_WAV_FLOAT; break;
                    case 0x06: cur.format = AUKO_WAV_ALAW; break;
                    case 0x07: cur.format = AUKO_WAV_ULAW; break;
                    case 0x11: cur.format = AUKO_WAV_ADPCM; break;
                    default: return fail("unsupported WAV file");
                    }
                } else return fail("unsupported WAV file");
            } else return fail("unsupported WAV file");                      /* A:1504 */
        } else if (!memcmp(id, "data", 4)) {
            if (pos + size > nbytes) return fail("invalid WAV file");        /* A:1507 */
            info->data_off = pos;
            info->data_size = size;
            info->have_data = 1;
            info->format = cur.format; info->channels = cur.channels; info->sampleRate = cur.sampleRate;
            info->blockAlign = cur.blockAlign; info->bitDepth = cur.bitDepth; info->have_fmt = cur.have_fmt;
            info->ncoef = cur.ncoef;
            memcpy(info->coef1, cur.coef1, sizeof info->coef1);
            memcpy(info->coef2, cur.coef2, sizeof info->coef2);
            pos += size;
        } else if (!memcmp(id, "LIST", 4)) {                                 /* A:1559-1568 */
            if (pos + 4 > nbytes) return fail("data string too short");
            if (!memcmp(data + pos, "INFO", 4)) {
                size_t e = pos + size;
                pos += 4;
                while (pos < e) {
                    /* "!2<c4s4Xh": id, u32 len, bytes, then pad to an even absolute offset */
                    if (pos + 8 > nbytes) return fail("data string too short");
                    size_t len = rd_u32(data + pos + 4);
                    if (pos + 8 + len > nbytes) return fail("data string too short");
                    if (info->ntags < 64) {
                        memcpy(info->tags[info->ntags].id, data + pos, 4);
                        info->tags[info->ntags].id[4] = 0;
                        info->tags[info->ntags].off = pos + 8;
                        info->tags[info->ntags].len = len;
                        info->ntags++;
                    }
                    pos += 8 + len;
                    if (pos & 1) {
                        if (pos + 1 > nbytes) return fail("data string too short");
                        pos += 1;
                    }
                }
            } else pos += size;
        } else pos += size;                                                  /* incl. "fact" */
    }
    if (!info->have_data) return fail("invalid WAV file");                   /* A:1573 */
    return 0;
}

/* ---------------------------------------------------------------- CPU-baseline chain */
double *auko_chain_s16(const uint8_t *data, size_t nbytes, int channels, double srcRate,
                       double dstRate, int interp, double peak, size_t *n_out) {
    size_t len = nbytes / 2 / (size_t)channels;
    double *dec = malloc(sizeof(double) * len * (size_t)channels + 8);
    if (!dec) return NULL;
    if (auko_pcm(data, nbytes, 16, AUKO_SIGNED, channels, 1, 0, dec, len, NULL)) { free(dec); return NULL; }
    size_t nl = auko_resample_len(len, srcRate, dstRate);
    double *rs = malloc(sizeof(double) * nl * (size_t)channels + 8);
    if (!rs) { free(dec); return NULL; }
    if (auko_resample(dec, len, channels, len, srcRate, dstRate, interp, rs, nl, NULL)) { free(dec); free(rs); return NULL; }
    free(dec);
    double *mono = malloc(sizeof(double) * nl + 8);
    if (!mono) { free(rs); return NULL; }
    auko_mono(rs, nl, channels, nl, mono);
    free(rs);
    auko_normalize(mono, nl, 1, nl, peak, 0);
    if (n_out) *n_out = nl;
    return mono;
}

/* ---------------------------------------------------------------- aukit.au, A:1634-1647 */
static uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

/* str_sub(s, i, j) with 1-based i >= 0 and a signed j (nil = -1): the 0-based [*off, *off + *len) it selects */
static void lua_sub(size_t n, long long i, long long j, size_t *off, size_t *len) {
    if (i < 1) i = 1;
    if (j < 0) j = (long long)n + j + 1;
    if (j > (long long)n) j = (long long)n;
    if (i > j) { *off = 0; *len = 0; return; }
    *off = (size_t)(i - 1);
    *len = (size_t)(j - i + 1);
}

int auko_au_parse(const uint8_t *data, size_t nbytes, auko_container_info *info) {
    g_err[0] = 0;
    memset(info, 0, sizeof *info);
    if (nbytes < 24) return fail("bad argument #2 to 'unpack' (data string too short)");                 /* str_unpack(">c4IIIII") */
    if (memcmp(data, ".snd", 4)) return fail("invalid AU file");            /* A:1637 */
    const uint32_t offset = be32(data + 4), size = be32(data + 8), encoding = be32(data + 12);
    info->sampleRate = (double)be32(data + 16);
    info->channels = (int)be32(data + 20);
    /* str_sub(data, offset, size ~= 0xFFFFFFFF and offset + size - 1 or nil): offset is used as a 1-based index */
    lua_sub(nbytes, (long long)offset, size != 0xFFFFFFFFu ? (long long)offset + (long long)size - 1 : -1,
            &info->data_off, &info->data_len);
    info->bigEndian = 1;
    switch (encoding) {                                                      /* A:1638-1645 */
    case 1: info->codec = 1; info->ulaw = 1; break;
    case 2: info->bitDepth = 8; info->dataType = AUKO_SIGNED; break;
    case 3: info->bitDepth = 16; info->dataType = AUKO_SIGNED; break;
    case 4: info->bitDepth = 24; info->dataType = AUKO_SIGNED; break;
    case 5: info->bitDepth = 32; info->dataType = AUKO_SIGNED; break;
    case 6: info->bitDepth = 32; info->dataType = AUKO_FLOAT; break;
    case 27: info->codec = 1; info->ulaw = 0; break;
    default: return fail("unsupported encoding type %u", (unsigned)encoding); /* A:1646 */
    }
    return 0;
}

/* ---------------------------------------------------------------- aukit.aiff, A:1580-1631 */
int auko_aiff_parse(const uint8_t *data, size_t nbytes, auko_container_info *info) {
    g_err[0] = 0;
    memset(info, 0, sizeof *info);
    if (nbytes < 4) return fail("bad argument #2 to 'unpack' (data string too short)");
    if (memcmp(data, "FORM", 4)) return fail("bad argument #1 (not an AIFF file)");   /* A:1585 */
    size_t pos = 8;                                                          /* 0-based; A:1586 skips the size */
    if (pos + 4 > nbytes) return fail("bad argument #2 to 'unpack' (data string too short)");
    int isAIFC = 0;
    if (!memcmp(data + pos, "AIFC", 4)) isAIFC = 1;
    else if (memcmp(data + pos, "AIFF", 4)) return fail("bad argument #1 (not an AIFF file)");
    pos += 4;
    int have_comm = 0, have_comp = 0;
    char comp[5] = {0};
    double length = 0;
    while (pos < nbytes) {                                                   /* pos <= #data, 1-based */
        if (pos + 8 > nbytes) return fail("bad argument #2 to 'unpack' (data string too short)");
        const uint8_t *id = data + pos;
        const uint32_t size = be32(data + pos + 4);
        pos += 8;
        if (!memcmp(id, "COMM", 4)) {                                        /* ">hIhHI7x", A:1597 */
            if (pos + 18 > nbytes) return fail("bad argument #2 to 'unpack' (data string too short)");
            info->channels = (int)(int16_t)((data[pos] << 8) | data[pos + 1]);
            const double frames = (double)be32(data + pos + 2);
            info->bitDepth = (int)(int16_t)((data[pos + 6] << 8) | data[pos + 7]);
            int e = (data[pos + 8] << 8) | data[pos + 9];
            double m = 0;                                                    /* I7 as a Lua number (double) */
            { unsigned long long mi = 0; for (int k = 0; k < 7; k++) mi = (mi << 8) | data[pos + 10 + k]; m = (double)mi; }
            pos += 18;
            if (isAIFC) {                                                    /* ">c4s1", A:1599-1600 */
                if (pos + 5 > nbytes) return fail("bad argument #2 to 'unpack' (data string too short)");
                memcpy(comp, data + pos, 4);
                have_comp = 1;
                const size_t sl = data[pos + 4];
                if (pos + 5 + sl > nbytes) return fail("bad argument #2 to 'unpack' (data string too short)");
                pos += 5 + sl;
                if (sl % 2 == 0) pos += 1;
            }
            length = frames * (double)info->channels * floor((double)info->bitDepth / 8.0);     /* A:1602 */
            const int sgn = (e & 0x8000) != 0;
            int ee = ((e & 0x7FFF) - 0x3FFE) % 0x800;                        /* Lua %: floored */
            if (ee < 0) ee += 0x800;
            info->sampleRate = ldexp(m * (sgn ? -1.0 : 1.0) / 72057594037927936.0, ee);          /* A:1605 */
            have_comm = 1;
        } else if (!memcmp(id, "SSND", 4)) {                                 /* A:1606 */
            if (pos + 8 > nbytes) return fail("bad argument #2 to 'unpack' (data string too short)");
            const uint32_t offset = be32(data + pos);
            pos += 8;
            if (!have_comm) return fail("attempt to perform arithmetic on a nil value (local 'length')");
            /* str_sub(data, pos + offset, pos + offset + length - 1), pos 1-based */
            const double i1 = (double)(pos + 1) + (double)offset, j1 = i1 + length - 1.0;
            lua_sub(nbytes, (long long)i1, (long long)j1, &info->data_off, &info->data_len);
            info->bigEndian = 1;
            info->dataType = AUKO_SIGNED;
            if (!have_comp || !memcmp(comp, "NONE", 4)) { /* pcm big-endian, A:1612 */ }
            else if (!memcmp(comp, "sowt", 4)) info->bigEndian = 0;
            else if (!memcmp(comp, "fl32", 4) || !memcmp(comp, "FL32", 4)) { info->bitDepth = 32; info->dataType = AUKO_FLOAT; }
            else if (!memcmp(comp, "alaw", 4) || !memcmp(comp, "ALAW", 4)) { info->codec = 1; info->ulaw = 0; }
            else if (!memcmp(comp, "ulaw", 4) || !memcmp(comp, "ULAW", 4)) { info->codec = 1; info->ulaw = 1; }
            else return fail("Unsupported compression scheme %.4s", comp);   /* A:1616 */
            return 0;
        } else {
            const char *key = !memcmp(id, "NAME", 4) ? "title" : !memcmp(id, "AUTH", 4) ? "artist"
                            : !memcmp(id, "(c) ", 4) ? "copyright" : !memcmp(id, "ANNO", 4) ? "comment" : NULL;
            if (key && info->nmeta < 16) {                                   /* A:1619-1630: str_sub(data, pos, pos+size-1) */
                size_t off, len;
                lua_sub(nbytes, (long long)pos + 1, (long long)pos + (long long)size, &off, &len);
                strcpy(info->meta[info->nmeta].key, key);
                info->meta[info->nmeta].off = off;
                info->meta[info->nmeta].len = len;
                info->nmeta++;
            }
            pos += size;
        }
    }
    return fail("invalid AIFF file");                                        /* A:1632 */
}

/* ---------------------------------------------------------------- host check of a device formula (test helper)
 * The CUDA kernels convert a non-negative 24-bit sample as fmaf(lo, RN(1 / (2^23 - 1)), lo) with lo = s * 2^-23 (and an
 * 8-bit one alike with 127 / 128).  Returns how many s in [0, 2^(bits-1)) give a float that differs from
 * (float)((double)s / (2^(bits-1) - 1)), the reference's value (A:1133) narrowed: must be 0. */
long auko_selftest_fma_scale(int bits) {
    const long n = 1L << (bits - 1);
    const float inv = 1.0f / (float)n;
    const float c = (float)(1.0 / (double)(n - 1));
    long bad = 0;
    for (long s = 0; s < n; s++) {
        const float ref = (float)((double)s / (double)(n - 1));
        const float lo = (float)s * inv;
        const float r = fmaf(lo, c, lo);
        if (memcmp(&r, &ref, sizeof r)) bad++;
    }
    return bad;
}
