/*
 * aukit_oracle.h -- CPU restatement of AUKit's preload path (TEST INFRASTRUCTURE ONLY).
 *
 * This is the parity oracle: a deliberately literal, single-threaded, double-precision C
 * restatement of the hot-path functions of /root/reference/aukit.lua (cited as A:line).
 * It is NOT part of the product. Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it. The product path (aukit_b200 ->
 * libaukit_cuda.so) never links, imports or calls anything in this directory.
 *
 * Pinning status: PINNED.  The reference ships no tests, fixtures or golden vectors (SURVEY.md
 * finding 2) and no Lua interpreter exists in this image, so the reference itself is executed
 * -- unmodified -- inside oracle/luavm (a small Lua 5.2 interpreter written for this purpose) to
 * produce tests/golden/reference_vectors.npz (156 seeded calls over every function below);
 * tests/test_golden_reference.py requires this oracle to reproduce every vector bit-exactly in
 * float64, errors included.  The hand-derived anchors of SURVEY.md Appendix B are checked too
 * (tests/test_oracle_anchors.py).  See tests/golden/README.md for what that does and does not prove.
 *
 * All sample outputs are doubles, planar: out[c * stride + i].
 */
#ifndef AUKIT_ORACLE_H
#define AUKIT_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { AUKO_SIGNED = 0, AUKO_UNSIGNED = 1, AUKO_FLOAT = 2 };
enum { AUKO_INTERP_NONE = 0, AUKO_INTERP_LINEAR = 1, AUKO_INTERP_CUBIC = 2, AUKO_INTERP_SINC = 3 };
/* ADPCM dialects: LITERAL = what aukit.wav/aukit.msadpcm actually do for channels 1/2
 * (bugs included); GENERAL = the standard N-channel layouts (authority for IMA:
 * aukit.stream.adpcm A:2798-2815), used where the reference itself errors (C > 2). */
enum { AUKO_DIALECT_LITERAL = 0, AUKO_DIALECT_GENERAL = 1 };

/* WAV formats after aukit.wav's fmt-chunk decode (A:1473-1504). */
enum {
    AUKO_WAV_PCM_SIGNED = 0, AUKO_WAV_PCM_UNSIGNED = 1, AUKO_WAV_FLOAT = 2, AUKO_WAV_ALAW = 3,
    AUKO_WAV_ULAW = 4, AUKO_WAV_ADPCM = 5, AUKO_WAV_MSADPCM = 6, AUKO_WAV_DFPWM = 7,
    AUKO_WAV_NONE = 8 /* data chunk seen before any fmt chunk: aukit.pcm defaults apply */
};

/* Last error message of the calling thread ("" if none).  Messages follow the reference's
 * error strings where it has one. */
const char *auko_last_error(void);

/* aukit.pcm (A:1049-1171). nbytes must be a multiple of (bitDepth/8)*channels, else error
 * "bad argument #1 (uneven amount of data per channel)".  out: [channels][len]. */
int auko_pcm(const uint8_t *data, size_t nbytes, int bitDepth, int dataType, int channels,
             int interleaved, int bigEndian, double *out, size_t stride, size_t *len);

/* aukit.g711 (A:1361-1384). Ragged input allowed: channel c gets ceil((nbytes-c)/channels)
 * samples; lens[c] receives each channel's length (may be NULL). */
int auko_g711(const uint8_t *data, size_t nbytes, int ulaw, int channels, double *out,
              size_t stride, size_t *lens);

/* One IMA step (A:1249-1255): updates *pred, *idx; returns the emitted sample. */
double auko_ima_step(int nibble, int *pred, int *idx);

/* aukit.adpcm (A:1183-1274) on a raw nibble string, default headerless form.
 * predictor/step_index may be NULL (zeros). */
int auko_adpcm(const uint8_t *data, size_t nbytes, int channels, int topFirst, int interleaved,
               const int *predictor, const int *step_index, double *out, size_t stride,
               size_t *len);

/* aukit.wav's IMA "adpcm" data-chunk path (A:1509-1548) incl. block framing + concat. */
int auko_wav_ima(const uint8_t *data, size_t nbytes, int blockAlign, int channels, int dialect,
                 double *out, size_t stride, size_t *len);
size_t auko_wav_ima_len(size_t nbytes, int blockAlign, int channels, int dialect);

/* aukit.msadpcm (A:1283-1353). coef1/coef2 NULL => defaults (A:1304); ncoef entries else. */
int auko_msadpcm(const uint8_t *data, size_t nbytes, int blockAlign, int channels,
                 const int *coef1, const int *coef2, int ncoef, int dialect, double *out,
                 size_t stride, size_t *len);
size_t auko_msadpcm_len(size_t nbytes, int blockAlign, int channels);

/* Audio:resample (A:653-673) with interpolate.none/linear/cubic (A:253-266).
 * n_out = floor(n_in * (dstRate/srcRate)) evaluated in double. */
size_t auko_resample_len(size_t n_in, double srcRate, double dstRate);
int auko_resample(const double *in, size_t in_stride, int channels, size_t n_in, double srcRate,
                  double dstRate, int interp, double *out, size_t out_stride, size_t *n_out);
/* The position the reference computes for 1-based output index i (A:666). */
double auko_resample_pos(uint64_t i, double srcRate, double dstRate);

/* Audio:mono (A:677-689). */
int auko_mono(const double *in, size_t in_stride, int channels, size_t n, double *out);

/* effects.amplify (A:3356-3369), in place. */
int auko_amplify(double *d, size_t stride, int channels, size_t n, double multiplier);

/* effects.normalize (A:3431-3459), in place. */
int auko_normalize(double *d, size_t stride, int channels, size_t n, double peak,
                   int independent);

/* encodePCM's per-sample formula (A:874): d*(d<0 and max or max-1)+add, un-rounded. */
double auko_encode_pcm(double d, int bitDepth, int dataType);
/* effects.invert A:3412, effects.fade A:3392, effects.delay A:3500, effects.center A:3465 (in place) */
int auko_invert(double *d, size_t stride, int channels, size_t n);
int auko_fade(double *d, size_t stride, int channels, size_t n, double sampleRate, double startTime,
              double startAmplitude, double endTime, double endAmplitude);
int auko_delay(double *d, size_t stride, int channels, size_t n, double sampleRate, double delay, double multiplier);
int auko_center(double *d, size_t stride, int channels, size_t n, double sampleRate);
/* Audio:pcm (A:901-911): all samples through encodePCM into a flat array (interleaved or channel-major). */
int auko_audio_pcm(const double *d, size_t stride, int channels, size_t n, int bitDepth, int dataType,
                   int interleaved, double *out);

/* effects.lowpass (A:3586-3598), in place. */
int auko_highpass(double *d, size_t stride, int channels, size_t n, double frequency, double sampleRate);
int auko_lowpass(double *d, size_t stride, int channels, size_t n, double frequency,
                 double sampleRate);

/* aukit.wav container walk (A:1456-1574): header parse only (decode is dispatched by the
 * caller onto the functions above).  The LAST data chunk wins (A:1505-1555). */
typedef struct {
    int format;            /* AUKO_WAV_* */
    int channels, sampleRate, blockAlign, bitDepth;
    int have_fmt;
    int ncoef;             /* msadpcm: 0 => defaults */
    int coef1[256], coef2[256];
    size_t data_off, data_size; /* payload of the last data chunk */
    int have_data;
    int ntags;             /* LIST/INFO tags (A:1559-1568), in file order */
    struct { char id[5]; size_t off, len; } tags[64];
} auko_wav_info;
int auko_wav_parse(const uint8_t *data, size_t nbytes, auko_wav_info *info);

/* aukit.au (A:1634-1647) and aukit.aiff (A:1580-1631): header parse only; the payload then goes through
 * auko_pcm (bigEndian as flagged) or auko_g711, exactly as the reference dispatches. */
typedef struct {
    int codec;             /* 0 = aukit.pcm, 1 = aukit.g711 */
    int bitDepth, dataType, bigEndian, ulaw;
    int channels;
    double sampleRate;
    size_t data_off, data_len; /* payload bytes (clipped to the file like str_sub) */
    int nmeta;             /* aiff: NAME/AUTH/"(c) "/ANNO -> title/artist/copyright/comment, file order */
    struct { char key[12]; size_t off, len; } meta[16];
} auko_container_info;
int auko_au_parse(const uint8_t *data, size_t nbytes, auko_container_info *info);
int auko_aiff_parse(const uint8_t *data, size_t nbytes, auko_container_info *info);

/* Whole auplay-style chain on one buffer, used for the CPU baseline timing:
 * s16le interleaved PCM -> resample -> mono -> normalize. Returns malloc'd mono doubles. */
double *auko_chain_s16(const uint8_t *data, size_t nbytes, int channels, double srcRate,
                       double dstRate, int interp, double peak, size_t *n_out);
void auko_free(void *p);
/* test helper: mismatches of the kernels' FMA sample scaling against the double division, over every non-negative sample */
long auko_selftest_fma_scale(int bits);

#ifdef __cplusplus
}
#endif
#endif
