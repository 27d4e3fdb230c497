/*
 * aukit_cuda.h -- C ABI of libaukit_cuda.so: AUKit's preload path on B200 (sm_100a).
 *
 * The reference (MCJack123/AUKit) is one interpreted Lua file with no FFI of its own; the
 * drop-in boundary is therefore the Lua module table of aukit.lua and its Audio object.
 * Each entry point below replaces one reference function ("A:n" = aukit.lua line n) and is
 * what the thin Lua module (luaopen_aukit_cuda, csrc/lua_binding.c) and the Python host
 * mirror (aukit_b200/) bind.  Plain pointers and sizes only.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; aukit_cuda_last_error()
 *     then returns a thread-local message that uses the reference's own error strings
 *     where it has one (A:1058-1064, A:656, A:1460-1507, A:1349 ...).
 *   - audio lives on the device as planar float32: sample (c, i) = data[c * stride + i],
 *     stride % 4 == 0 and data 256-byte aligned (one cudaMalloc per Audio).
 *   - all kernels are enqueued on the context's stream; host-visible results (download,
 *     deferred decode errors) synchronise that stream.
 *   - there is NO CPU fallback: without a CUDA device every call fails.
 */
#ifndef AUKIT_CUDA_H
#define AUKIT_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AUKIT_CUDA_ABI_VERSION 1

typedef struct aukit_ctx aukit_ctx;     /* one per (host thread, device) */
typedef struct aukit_audio aukit_audio; /* device-resident aukit.Audio (A:116-123) */

enum { AUKIT_SIGNED = 0, AUKIT_UNSIGNED = 1, AUKIT_FLOAT = 2 };              /* dataType */
enum { AUKIT_INTERP_NONE = 0, AUKIT_INTERP_LINEAR = 1, AUKIT_INTERP_CUBIC = 2,
       AUKIT_INTERP_SINC = 3 /* Audio:resample only (A:267-281, window +-10); not in the fused chain */ };
/* ADPCM dialect: LITERAL reproduces aukit.wav / aukit.msadpcm for channels 1 and 2 exactly
 * (A:1544 index mask, A:1331 first-header reuse); GENERAL is the standard N-channel block
 * layout (IMA: aukit.stream.adpcm A:2798-2815), the only one defined for channels > 2. */
enum { AUKIT_DIALECT_LITERAL = 0, AUKIT_DIALECT_GENERAL = 1 };
enum {                                                                       /* aukit.wav dataType */
    AUKIT_WAV_PCM_SIGNED = 0, AUKIT_WAV_PCM_UNSIGNED = 1, AUKIT_WAV_FLOAT = 2, AUKIT_WAV_ALAW = 3,
    AUKIT_WAV_ULAW = 4, AUKIT_WAV_ADPCM = 5, AUKIT_WAV_MSADPCM = 6, AUKIT_WAV_DFPWM = 7,
    AUKIT_WAV_NONE = 8
};

/* ------------------------------------------------------------------ context */
int aukit_cuda_abi_version(void);
const char *aukit_cuda_last_error(void);
/* device < 0 => current device.  Fails (no fallback) when no CUDA device is usable. */
int aukit_cuda_init(int device, aukit_ctx **ctx);
void aukit_cuda_shutdown(aukit_ctx *ctx);
/* Use an external cudaStream_t (e.g. torch's current stream); NULL => the ctx's own.  The legacy
 * default stream is cudaStreamLegacy ((void *)1), not NULL. */
int aukit_cuda_set_stream(aukit_ctx *ctx, void *cuda_stream);
void *aukit_cuda_get_stream(aukit_ctx *ctx);
/* A process that holds contexts on SEVERAL devices makes one current (cudaSetDevice) before issuing calls on it; the
 * aukit_cuda_group_* calls do this themselves. */
int aukit_cuda_make_current(aukit_ctx *ctx);
/* Waits for the stream and reports deferred device-side decode errors (bad IMA step index,
 * A:1213; bad MS-ADPCM predictor index, A:1311). */
int aukit_cuda_synchronize(aukit_ctx *ctx);
/* Kernel launches issued by this context since init (bench.py's gpu_launches). */
uint64_t aukit_cuda_launch_count(aukit_ctx *ctx);

/* ------------------------------------------------------------------ Audio handle (A:116-123) */
/* aukit.new-like: allocates zero-filled [channels][frames] (A:1784-1799 with duration*rate frames). */
int aukit_cuda_audio_new(aukit_ctx *ctx, int channels, size_t frames, double sampleRate,
                         aukit_audio **out);
/* Wrap caller-owned device memory (not freed by audio_free). */
int aukit_cuda_audio_wrap(aukit_ctx *ctx, float *d_data, int channels, size_t frames, size_t stride,
                          double sampleRate, aukit_audio **out);
void aukit_cuda_audio_free(aukit_ctx *ctx, aukit_audio *a);
int aukit_cuda_audio_channels(const aukit_audio *a);                         /* Audio:channels A:644 */
size_t aukit_cuda_audio_frames(const aukit_audio *a);                        /* #data[1] */
size_t aukit_cuda_audio_stride(const aukit_audio *a);
double aukit_cuda_audio_sample_rate(const aukit_audio *a);
float *aukit_cuda_audio_data(const aukit_audio *a);                          /* device pointer */
/* audio.sampleRate = x: a plain writable field in the reference (effects.speed assigns it, A:3383); every later
 * device call (resample, fade, delay, center, low/highpass, Audio:len) reads this value. */
int aukit_cuda_audio_set_sample_rate(aukit_audio *a, double sampleRate);
/* Per-channel length; differs from frames only for ragged G.711 input (A:1379). */
size_t aukit_cuda_audio_channel_frames(const aukit_audio *a, int channel);
/* Copy channel `channel` frames [first, first+count) to host floats (synchronises). */
int aukit_cuda_audio_download(aukit_ctx *ctx, const aukit_audio *a, int channel, size_t first,
                              size_t count, float *h_out);
/* Copy host floats into channel `channel` (audio.data[c][i] = v writes). */
int aukit_cuda_audio_upload(aukit_ctx *ctx, aukit_audio *a, int channel, size_t first, size_t count,
                            const float *h_in);

/* ------------------------------------------------------------------ loaders (host bytes in) */
/* aukit.pcm(data, bitDepth, dataType, channels, sampleRate, interleaved, bigEndian) A:1049 */
int aukit_cuda_pcm(aukit_ctx *ctx, const void *h_data, size_t nbytes, int bitDepth, int dataType,
                   int channels, double sampleRate, int interleaved, int bigEndian,
                   aukit_audio **out);
/* aukit.g711(data, ulaw, channels, sampleRate) A:1361 */
int aukit_cuda_g711(aukit_ctx *ctx, const void *h_data, size_t nbytes, int ulaw, int channels,
                    double sampleRate, aukit_audio **out);
/* aukit.adpcm(data, channels, sampleRate, topFirst, interleaved, predictor, step_index) A:1183
 * for string input (one serial chain per channel; predictor/step_index NULL => zeros). */
int aukit_cuda_adpcm(aukit_ctx *ctx, const void *h_data, size_t nbytes, int channels,
                     double sampleRate, int topFirst, int interleaved, const int *predictor,
                     const int *step_index, aukit_audio **out);
/* aukit.wav's IMA ADPCM data path: block framing + aukit.adpcm + concat, A:1509-1548 */
int aukit_cuda_ima_adpcm_wav(aukit_ctx *ctx, const void *h_data, size_t nbytes, int blockAlign,
                             int channels, double sampleRate, int dialect, aukit_audio **out);
/* aukit.msadpcm(data, blockAlign, channels, sampleRate, coefficients) A:1283 */
int aukit_cuda_msadpcm(aukit_ctx *ctx, const void *h_data, size_t nbytes, int blockAlign,
                       int channels, double sampleRate, const int *coef1, const int *coef2,
                       int ncoef, int dialect, aukit_audio **out);

/* aukit.wav(data, head) A:1456: container walk on the host, decode on the device. */
typedef struct {
    int format;                 /* AUKIT_WAV_* */
    int channels, sampleRate, blockAlign, bitDepth, have_fmt;
    int ncoef;                  /* msadpcm coefficient pairs from the fmt chunk; 0 => defaults */
    int coef1[256], coef2[256];
    size_t data_off, data_size; /* payload of the LAST data chunk (A:1505-1555) */
    int ntags;                  /* LIST/INFO entries (A:1559-1568) */
    struct { char id[5]; size_t off, len; } tags[64];
} aukit_wav_info;
int aukit_cuda_wav_parse(const void *h_data, size_t nbytes, aukit_wav_info *info); /* host only */
int aukit_cuda_wav(aukit_ctx *ctx, const void *h_data, size_t nbytes, int head_only, int dialect,
                   aukit_wav_info *info_out, aukit_audio **out);

/* aukit.au(data) A:1634-1647 and aukit.aiff(data, head) A:1580-1631: header walk on the host, payload
 * through aukit_cuda_pcm (bigEndian as flagged) or aukit_cuda_g711, as the reference dispatches.  The
 * reference's own index arithmetic is kept (see csrc/containers.cu). */
enum { AUKIT_CODEC_PCM = 0, AUKIT_CODEC_G711 = 1 };
typedef struct {
    int codec;                  /* AUKIT_CODEC_* */
    int bitDepth, dataType, bigEndian, ulaw;
    int channels;
    double sampleRate;
    size_t data_off, data_len;  /* payload bytes inside the file */
    int nmeta;                  /* aiff NAME/AUTH/"(c) "/ANNO -> title/artist/copyright/comment, file order */
    struct { char key[12]; size_t off, len; } meta[16];
} aukit_container_info;
int aukit_cuda_au_parse(const void *h_data, size_t nbytes, aukit_container_info *info);   /* host only */
int aukit_cuda_aiff_parse(const void *h_data, size_t nbytes, aukit_container_info *info); /* host only */
int aukit_cuda_au(aukit_ctx *ctx, const void *h_data, size_t nbytes, aukit_container_info *info_out,
                  aukit_audio **out);
int aukit_cuda_aiff(aukit_ctx *ctx, const void *h_data, size_t nbytes, int head_only,
                    aukit_container_info *info_out, aukit_audio **out);

/* ------------------------------------------------------------------ transforms (new Audio) */
/* Audio:resample(sampleRate, interpolation) A:653 */
int aukit_cuda_resample(aukit_ctx *ctx, const aukit_audio *in, double sampleRate, int interpolation,
                        aukit_audio **out);
/* Audio:mono() A:677 */
int aukit_cuda_mono(aukit_ctx *ctx, const aukit_audio *in, aukit_audio **out);
/* Audio:concat(...) A:696 for same-rate inputs (the block join of A:1548). */
int aukit_cuda_concat(aukit_ctx *ctx, const aukit_audio *const *parts, int nparts, aukit_audio **out);

/* ------------------------------------------------------------------ effects (in place) */
/* aukit.effects.amplify(audio, multiplier) A:3356 */
int aukit_cuda_amplify(aukit_ctx *ctx, aukit_audio *a, double multiplier);
/* aukit.effects.normalize(audio, peakAmplitude, independent) A:3431 (single device) */
int aukit_cuda_normalize(aukit_ctx *ctx, aukit_audio *a, double peakAmplitude, int independent);
/* The two halves of normalize, for time-sharded buffers: d_max is DEVICE memory holding 1
 * float (global) or `channels` floats (independent).  absmax max-combines into whatever
 * d_max already holds (zero it first); between the two calls the caller all-reduces d_max
 * with MAX over the ranks (the path's only collective). */
int aukit_cuda_absmax(aukit_ctx *ctx, const aukit_audio *a, int independent, float *d_max);
int aukit_cuda_scale_clamp(aukit_ctx *ctx, aukit_audio *a, double peakAmplitude, int independent,
                           const float *d_max);

/* aukit.effects.lowpass(audio, frequency) A:3586-3598: in-place one-pole IIR per channel,
 * a = 1 - exp(-(frequency/sampleRate) * 2 pi), d[i] = d[i-1] + a*(d[i] - d[i-1]) for i >= 2 (auplay.lua:30).
 * Single pass, chained tiles (decoupled look-back), fp64 state. */
int aukit_cuda_lowpass(aukit_ctx *ctx, aukit_audio *a, double frequency);
/* aukit.effects.highpass(audio, frequency) A:3605-3618: a = 1 / (2 pi frequency/sampleRate + 1),
 * d[i] = a * (d[i-1] + x[i] - x[i-1]) for i >= 2; the same chained-tile scan with ratio a. */
int aukit_cuda_highpass(aukit_ctx *ctx, aukit_audio *a, double frequency);

/* The remaining elementwise in-place effects: effects.invert A:3412, effects.fade A:3392, effects.delay A:3500,
 * effects.center A:3465 (argument meaning as in the reference; times in seconds). */
int aukit_cuda_invert(aukit_ctx *ctx, aukit_audio *a);
int aukit_cuda_fade(aukit_ctx *ctx, aukit_audio *a, double startTime, double startAmplitude,
                    double endTime, double endAmplitude);
int aukit_cuda_delay(aukit_ctx *ctx, aukit_audio *a, double delay, double multiplier);
int aukit_cuda_center(aukit_ctx *ctx, aukit_audio *a);

/* Audio:pcm(bitDepth, dataType, interleaved) A:901-911 -> encodePCM A:868-894: every sample as
 * d * (d < 0 ? 2^(b-1) : 2^(b-1)-1) + (unsigned ? 2^(b-1) : 0), UN-ROUNDED like the reference's Lua
 * numbers, written as doubles to h_out[frames*channels] (interleaved: [n*C + c], else [c*frames + n]). */
int aukit_cuda_audio_pcm(aukit_ctx *ctx, const aukit_audio *a, int bitDepth, int dataType,
                         int interleaved, double *h_out);
/* The same values as packed little-endian integers of bitDepth/8 bytes (Audio:wav's sample packing,
 * A:981-985; float = the f32 bits).  rounding: 0 truncate (Cobalt / C cast), 1 floor, 2 nearest-even --
 * the reference leaves this to the host's string.pack.  Out-of-range values saturate. */
int aukit_cuda_audio_pcm_bytes(aukit_ctx *ctx, const aukit_audio *a, int bitDepth, int dataType,
                               int interleaved, int rounding, void *h_out);

/* Audio:stream(chunkSize, bitDepth, dataType) A:921-937: one step of the chunk iterator.  Frames [first, first+count)
 * of every channel through encodePCM (A:868-894, `multiple` branch), un-rounded doubles, planar
 * h_out[c * count + k]; *got = frames written (clipped at the end of the audio; 0 => the iterator is done, A:878). */
int aukit_cuda_audio_stream_chunk(aukit_ctx *ctx, const aukit_audio *a, int bitDepth, int dataType,
                                  size_t first, size_t count, double *h_out, size_t *got);

/* ------------------------------------------------------------------ device-pointer level */
/* Same kernels on caller-owned DEVICE buffers (inputs already resident in HBM; what
 * bench.py's `value` times).  Output layout: d_out[c * out_stride + i]. */
int aukit_cuda_dev_pcm(aukit_ctx *ctx, const void *d_in, size_t nbytes, int bitDepth, int dataType,
                       int channels, int interleaved, int bigEndian, float *d_out,
                       size_t out_stride);
int aukit_cuda_dev_g711(aukit_ctx *ctx, const void *d_in, size_t nbytes, int ulaw, int channels,
                        float *d_out, size_t out_stride);
int aukit_cuda_dev_ima_adpcm_wav(aukit_ctx *ctx, const void *d_in, size_t nbytes, int blockAlign,
                                 int channels, int dialect, float *d_out, size_t out_stride);
int aukit_cuda_dev_msadpcm(aukit_ctx *ctx, const void *d_in, size_t nbytes, int blockAlign,
                           int channels, const int *coef1, const int *coef2, int ncoef,
                           int dialect, float *d_out, size_t out_stride);
/* Time-shardable resample: produces global output frames [out_first, out_first+n_out)
 * (0-based) of a signal whose full length is n_in_total frames.  d_in holds input frames
 * [in_first, in_first + in_avail) of every channel.  Positions are computed from the GLOBAL
 * output index in fp64 exactly as A:666 does; use aukit_resample_window() to size the halo. */
int aukit_cuda_dev_resample(aukit_ctx *ctx, const float *d_in, size_t in_stride, int channels,
                            uint64_t n_in_total, uint64_t in_first, size_t in_avail,
                            double srcRate, double dstRate, int interpolation,
                            uint64_t out_first, size_t n_out, float *d_out, size_t out_stride);
int aukit_cuda_dev_mono(aukit_ctx *ctx, const float *d_in, size_t in_stride, int channels, size_t n,
                        float *d_out);
int aukit_cuda_dev_amplify(aukit_ctx *ctx, float *d, size_t stride, int channels, size_t n,
                           double multiplier);
int aukit_cuda_dev_lowpass(aukit_ctx *ctx, float *d, size_t stride, int channels, size_t n,
                           double frequency, double sampleRate);
int aukit_cuda_dev_invert(aukit_ctx *ctx, float *d, size_t stride, int channels, size_t n);
int aukit_cuda_dev_fade(aukit_ctx *ctx, float *d, size_t stride, int channels, size_t n, double sampleRate,
                        double startTime, double startAmplitude, double endTime, double endAmplitude);
int aukit_cuda_dev_delay(aukit_ctx *ctx, float *d, size_t stride, int channels, size_t n,
                         double sampleRate, double delay, double multiplier);
int aukit_cuda_dev_center(aukit_ctx *ctx, float *d, size_t stride, int channels, size_t n,
                          double sampleRate);
int aukit_cuda_dev_encode_pcm(aukit_ctx *ctx, const float *d, size_t stride, int channels, size_t n,
                              int bitDepth, int dataType, int interleaved, double *d_out);
int aukit_cuda_dev_encode_pcm_bytes(aukit_ctx *ctx, const float *d, size_t stride, int channels,
                                    size_t n, int bitDepth, int dataType, int interleaved,
                                    int rounding, void *d_out);
int aukit_cuda_dev_highpass(aukit_ctx *ctx, float *d, size_t stride, int channels, size_t n,
                            double frequency, double sampleRate);
int aukit_cuda_dev_absmax(aukit_ctx *ctx, const float *d, size_t stride, int channels, size_t n,
                          int independent, float *d_max);
int aukit_cuda_dev_scale_clamp(aukit_ctx *ctx, float *d, size_t stride, int channels, size_t n,
                               double peakAmplitude, int independent, const float *d_max);

/* Fused auplay chain (auplay.lua:12-27) for integer/float PCM input resident on the device:
 * unpack -> resample(dstRate, interpolation) -> [mono] -> normalize(peak), without
 * materialising intermediates.  Two launches around the max barrier:
 *   pass 1 (peak)  reads the packed bytes, max-combines |result| into d_max[0];
 *   (caller all-reduces d_max with MAX when time-sharded)
 *   pass 2 (apply) re-reads the packed bytes and writes clamp(result * peak/max).
 * d_in holds interleaved frames [in_first, in_first + in_avail) of the full signal. */
typedef struct {
    int bitDepth, dataType, channels, bigEndian;  /* interleaved input only */
    double srcRate, dstRate;
    int interpolation;
    int mono;                   /* 1 => Audio:mono() after resample */
    uint64_t n_in_total;        /* frames of the whole (unsharded) signal */
    uint64_t in_first;          /* global index of the first frame present in d_in */
    size_t in_avail;            /* frames present in d_in */
    uint64_t out_first;         /* first global output frame this call produces */
    size_t n_out;               /* number of output frames this call produces */
} aukit_pipeline_desc;
int aukit_cuda_dev_pipeline_peak(aukit_ctx *ctx, const aukit_pipeline_desc *p, const void *d_in,
                                 float *d_max);
int aukit_cuda_dev_pipeline_apply(aukit_ctx *ctx, const aukit_pipeline_desc *p, const void *d_in,
                                  double peakAmplitude, const float *d_max, float *d_out,
                                  size_t out_stride);
/* A batch of clips through aukit.pcm (A:1049) -> Audio:resample(dstRate, interpolation) (A:653) -> effects.amplify(multiplier)
 * (A:3356) in ONE pass per rate class (BASELINE config 3; replaces a loop of those three calls over a playlist).
 * All clips share the sample format; every clip has its own length and rate.  d_in holds the packed interleaved
 * frames of all clips (16-byte aligned base; clip k starts at byte in_offset), d_out receives planar float32 rows:
 * sample (c, i) of clip k = d_out[out_offset + c * out_stride + i].  aukit_batch_plan fills n_out
 * (= floor(frames * dstRate / srcRate), A:658-664) and, for clips whose out_stride is 0, lays the outputs out back to back
 * (rows padded to 32 floats); it returns the number of floats d_out must hold. */
typedef struct {
    uint64_t in_offset;     /* bytes */
    uint64_t frames;        /* input frames */
    double srcRate;
    uint64_t out_offset;    /* floats */
    uint64_t out_stride;    /* floats between channel rows; 0 => let aukit_batch_plan choose */
    uint64_t n_out;         /* output frames (written by aukit_batch_plan) */
} aukit_clip;
uint64_t aukit_batch_plan(aukit_clip *clips, size_t nclips, int out_channels, double dstRate);
int aukit_cuda_dev_batch_resample_amplify(aukit_ctx *ctx, const aukit_clip *clips, size_t nclips, int bitDepth,
                                          int dataType, int channels, int bigEndian, double dstRate,
                                          int interpolation, double multiplier, const void *d_in, float *d_out);

/* The same from HOST clips (one pointer / byte count / rate per clip, e.g. a table of Lua strings): uploads them into one
 * device buffer, runs the batch, returns one Audio per clip in out[0..nclips).  The Audios share one device allocation
 * that is released with the last of them. */
int aukit_cuda_batch_resample_amplify(aukit_ctx *ctx, const void *const *h_clips, const size_t *clip_bytes,
                                      const double *srcRates, size_t nclips, int bitDepth, int dataType,
                                      int channels, int bigEndian, double dstRate, int interpolation,
                                      double multiplier, aukit_audio **out);

/* Host string in, device-resident Audio out: aukit.pcm -> Audio:resample -> [Audio:mono] -> effects.normalize
 * (A:1049, A:653, A:677, A:3431; auplay.lua:12-27) in the two fused passes.  p describes the whole call (in_first = 0,
 * in_avail = n_in_total, out_first = 0, n_out = aukit_resample_out_len(...)). */
int aukit_cuda_preload_audio(aukit_ctx *ctx, const aukit_pipeline_desc *p, const void *h_in, size_t nbytes,
                             double peakAmplitude, aukit_audio **out);
/* Host-buffer convenience = the end-to-end call (H2D + peak + apply + D2H), single device. */
int aukit_cuda_pipeline_host(aukit_ctx *ctx, const aukit_pipeline_desc *p, const void *h_in,
                             size_t nbytes, double peakAmplitude, float *h_out);

/* Pipelined host-buffer preloader: the same end-to-end call, double buffered.  PCIe is full duplex but
 * one clip cannot use it (the normalisation peak needs the whole upload before the first output can
 * leave), so the preloader overlaps clip i's download with clip i+1's upload: H2D on one stream, the two
 * passes on a second, D2H on a third, `slots` device buffers in rotation.  Replaces a loop of
 * aukit.wav / Audio:resample / Audio:mono / effects.normalize calls over a playlist (A:1373, A:653,
 * A:678, A:3434).  Host buffers should be pinned (aukit_cuda_host_alloc) for the copies to be async.
 *   submit  = begin + finish;  begin/finish are split so that a multi-GPU caller can all-reduce the
 *   slot's peak (aukit_cuda_preloader_peak_ptr, on aukit_cuda_preloader_stream) between the passes.
 *   drain   = wait for every submitted clip and surface device-side errors. */
typedef struct aukit_preloader aukit_preloader;
/* max_out_samples: output channels x frames of the largest clip (rows are padded to 32 frames). */
int aukit_cuda_preloader_create(aukit_ctx *ctx, size_t max_in_bytes, size_t max_out_samples, int slots,
                                aukit_preloader **out);
void aukit_cuda_preloader_destroy(aukit_preloader *pl);
int aukit_cuda_preloader_begin(aukit_preloader *pl, const aukit_pipeline_desc *p, const void *h_in,
                               size_t nbytes, int *slot);
float *aukit_cuda_preloader_peak_ptr(aukit_preloader *pl, int slot);
void *aukit_cuda_preloader_stream(aukit_preloader *pl);
int aukit_cuda_preloader_finish(aukit_preloader *pl, int slot, double peakAmplitude, float *h_out);
int aukit_cuda_preloader_submit(aukit_preloader *pl, const aukit_pipeline_desc *p, const void *h_in,
                                size_t nbytes, double peakAmplitude, float *h_out);
int aukit_cuda_preloader_drain(aukit_preloader *pl);
/* pinned host memory for the calls above */
int aukit_cuda_host_alloc(size_t nbytes, void **out);
void aukit_cuda_host_free(void *p);

/* ------------------------------------------------------------------ multi-GPU: the one collective
 * effects.normalize (A:3431-3459) on a buffer that is time-sharded over several GPUs needs the GLOBAL max of A:3438-3443:
 * a MAX of one float per rank (one per channel when `independent`).  aukit_comm does that exchange inside the library,
 * over peer-mapped device memory (NVLink / NVSwitch): one tiny kernel per rank between the two passes, no NCCL launch, no
 * host round trip.  Results are bit-identical to a single GPU (MAX is exact and order-free).
 *   one process per GPU (torchrun / MPI): comm_create on every rank, exchange the comm_handle blobs (comm_handle_bytes
 *     each) by any host means (torch.distributed all_gather, MPI, a file), comm_connect with all of them in rank order;
 *   one process driving all GPUs (the Lua module, a single host thread): comm_create per device, comm_connect_local.
 * Every rank must make the same sequence of exchange calls. */
typedef struct aukit_comm aukit_comm;
int aukit_cuda_comm_create(aukit_ctx *ctx, int world, int rank, aukit_comm **out);
size_t aukit_cuda_comm_handle_bytes(void);
int aukit_cuda_comm_handle(aukit_comm *c, void *handle_out);
int aukit_cuda_comm_connect(aukit_comm *c, const void *handles_in_rank_order);
int aukit_cuda_comm_connect_local(aukit_comm *const *comms, int world);
void aukit_cuda_comm_destroy(aukit_comm *c);
/* MAX-combine nvals (<= 16) non-negative device floats over the ranks, in place, in stream order. */
int aukit_cuda_comm_allreduce_max(aukit_comm *c, float *d_vals, int nvals);
float *aukit_cuda_comm_values(aukit_comm *c);      /* the communicator's own 16-float device scratch */
/* effects.normalize(audio, peakAmplitude, independent) where `a` is THIS rank's time shard of the Audio. */
int aukit_cuda_comm_normalize(aukit_comm *c, aukit_audio *a, double peakAmplitude, int independent);
/* The fused chain (aukit_cuda_dev_pipeline_peak -> exchange -> aukit_cuda_dev_pipeline_apply) on this rank's shard. */
int aukit_cuda_comm_pipeline(aukit_comm *c, const aukit_pipeline_desc *p, const void *d_in,
                             double peakAmplitude, float *d_out, size_t out_stride);

/* One host thread driving several GPUs (what the Lua module does: the reference is a single coroutine): a
 * group = one context + one communicator per listed device, connected by plain peer pointers.  devices == NULL: devices
 * 0..ndev-1; ndev <= 0: every visible device.  A device may be listed twice (two contexts / streams on it). */
typedef struct aukit_group aukit_group;
int aukit_cuda_group_create(const int *devices, int ndev, aukit_group **out);
void aukit_cuda_group_destroy(aukit_group *g);
int aukit_cuda_group_size(const aukit_group *g);
aukit_ctx *aukit_cuda_group_ctx(aukit_group *g, int i);
aukit_comm *aukit_cuda_group_comm(aukit_group *g, int i);
/* effects.normalize (A:3431) on an Audio held as one time shard per group member (shards[i] lives on context i). */
int aukit_cuda_group_normalize(aukit_group *g, aukit_audio *const *shards, double peakAmplitude, int independent);
/* auplay's chain on ONE host buffer, time-sharded over the group (tile-aligned output ranges, halo windows, local peak
 * passes, the MAX exchange, local apply passes); h_out[c * n_out + i], n_out = floor(n_in_total * dstRate / srcRate).
 * `whole` describes the unsharded call (in_first = 0, in_avail = n_in_total, out_first = 0).  Same bits as one GPU. */
int aukit_cuda_group_preload(aukit_group *g, const aukit_pipeline_desc *whole, const void *h_in, size_t nbytes,
                             double peakAmplitude, float *h_out);
/* The same, gathered device-to-device into ONE Audio owned by `owner` (any context; the Lua module passes its own). */
int aukit_cuda_group_preload_audio(aukit_group *g, aukit_ctx *owner, const aukit_pipeline_desc *whole, const void *h_in,
                                   size_t nbytes, double peakAmplitude, aukit_audio **out);

/* ------------------------------------------------------------------ pure host helpers */
/* floor(n_in * (dstRate/srcRate)) in double, the reference's loop bound (A:658-664). */
uint64_t aukit_resample_out_len(uint64_t n_in, double srcRate, double dstRate);
/* (i - 1)/ratio + 1 for the 1-based output index i, as A:666 computes it. */
double aukit_resample_position(uint64_t i, double srcRate, double dstRate);
/* Input frames (0-based, clipped to [0, n_in_total)) needed to produce global output frames
 * [out_first, out_first + n_out): the shard plus its interpolation halo (-1/+2 cubic,
 * 0/+1 linear, 0/0 none). */
int aukit_resample_window(uint64_t n_in_total, double srcRate, double dstRate, int interpolation,
                          uint64_t out_first, uint64_t n_out, uint64_t *in_first,
                          uint64_t *in_count);
/* ADPCM data shards by contiguous block range (every block header is the full decoder state, A:1310 / A:1513): rank's
 * blocks are [*first, *first + *count); decode them with the dev_* ADPCM calls on that byte range -- no halo, no collective. */
int aukit_block_shard(uint64_t nblocks, int world, int rank, uint64_t *first, uint64_t *count);
size_t aukit_ima_adpcm_wav_frames(size_t nbytes, int blockAlign, int channels, int dialect);
size_t aukit_msadpcm_frames(size_t nbytes, int blockAlign, int channels);

#ifdef __cplusplus
}
#endif
#endif
