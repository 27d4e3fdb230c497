// capi.cu -- context, Audio handles, host-buffer loaders and the aukit.wav container walk.
//
// Host-side half of the C ABI in include/aukit_cuda.h.  Everything numeric happens in the
// kernels of pcm.cu / adpcm.cu / resample.cu / effects.cu / pipeline*.cu; this file owns
// device memory (stream-ordered allocations from the device's default pool), the H2D / D2H
// copies of the end-to-end calls, and the integer-only parsing of RIFF headers (A:1456-1574).
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <sched.h>
#include <ctype.h>
#include "common.cuh"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int aukit_launch_adpcm_stream(aukit_ctx *ctx, const uint8_t *d_in, size_t len, int channels, int topFirst,
                              int interleaved, const int *d_pred, const int *d_idx, float *d_out, size_t stride);

static thread_local char g_err[512];

int aukit_fail(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return -1;
}

int aukit_cuda_check(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return 0;
    return aukit_fail("CUDA error in %s: %s", what, cudaGetErrorString(e));
}

extern "C" const char *aukit_cuda_last_error(void) { return g_err; }
extern "C" int aukit_cuda_abi_version(void) { return AUKIT_CUDA_ABI_VERSION; }

// ------------------------------------------------------------------ context
extern "C" int aukit_cuda_init(int device, aukit_ctx **out) {
    g_err[0] = 0;
    if (!out) return aukit_fail("aukit_cuda: null argument");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return aukit_fail("aukit_cuda: no CUDA device available (%s); there is no CPU fallback",
                          e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0) AUKIT_CUDA_TRY(cudaGetDevice(&device));
    if (device >= count) return aukit_fail("aukit_cuda: device %d out of range (%d devices)", device, count);
    AUKIT_CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    AUKIT_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return aukit_fail("aukit_cuda: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                          prop.major, prop.minor);
    aukit_ctx *ctx = static_cast<aukit_ctx *>(calloc(1, sizeof(aukit_ctx)));
    if (!ctx) return aukit_fail("aukit_cuda: out of host memory");
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    AUKIT_CUDA_TRY(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
    ctx->stream = ctx->own_stream;
    AUKIT_CUDA_TRY(cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking));
    AUKIT_CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    AUKIT_CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
    AUKIT_CUDA_TRY(cudaMalloc(&ctx->d_status, sizeof(int)));
    AUKIT_CUDA_TRY(cudaMemset(ctx->d_status, 0, sizeof(int)));
    AUKIT_CUDA_TRY(cudaMallocHost(&ctx->h_status, sizeof(int)));
    AUKIT_CUDA_TRY(cudaMalloc(&ctx->d_scratch, 1024 * sizeof(float)));
    AUKIT_CUDA_TRY(cudaMalloc(&ctx->d_hint, sizeof(int)));
    AUKIT_CUDA_TRY(cudaMemset(ctx->d_hint, 0, sizeof(int)));
    // keep freed blocks in the pool: the end-to-end calls allocate per call
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    *out = ctx;
    return 0;
}

extern "C" void aukit_cuda_shutdown(aukit_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    cudaFree(ctx->d_status);
    cudaFree(ctx->d_scratch);
    cudaFree(ctx->d_hint);
    cudaFreeHost(ctx->h_status);
    cudaStreamSynchronize(ctx->side_stream);
    cudaEventDestroy(ctx->ev_fork);
    cudaEventDestroy(ctx->ev_join);
    cudaStreamDestroy(ctx->side_stream);
    cudaStreamDestroy(ctx->own_stream);
    free(ctx);
}

extern "C" int aukit_cuda_set_stream(aukit_ctx *ctx, void *cuda_stream) {
    if (!ctx) return aukit_fail("aukit_cuda: null context");
    ctx->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
    return 0;
}

extern "C" int aukit_cuda_make_current(aukit_ctx *ctx) {
    if (!ctx) return aukit_fail("aukit_cuda: null context");
    return aukit_cuda_check(cudaSetDevice(ctx->device), "cudaSetDevice");
}

extern "C" void *aukit_cuda_get_stream(aukit_ctx *ctx) { return ctx ? ctx->stream : nullptr; }
extern "C" uint64_t aukit_cuda_launch_count(aukit_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int aukit_cuda_synchronize(aukit_ctx *ctx) {
    if (!ctx) return aukit_fail("aukit_cuda: null context");
    AUKIT_CUDA_TRY(cudaMemcpyAsync(ctx->h_status, ctx->d_status, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    AUKIT_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    const int st = *ctx->h_status;
    if (st) {
        cudaMemsetAsync(ctx->d_status, 0, sizeof(int), ctx->stream);
        if (st & AUKIT_DEVERR_COMM_TIMEOUT)
            return aukit_fail("aukit_cuda: a rank never arrived at the normalize MAX exchange (10 s timeout); the output of this call is invalid");
        if (st & AUKIT_DEVERR_IMA_INDEX)
            return aukit_fail("number outside of range (expected step index to be within 0 and 88)");   // A:1213
        return aukit_fail("attempt to perform arithmetic on a nil value (MS-ADPCM predictor index has no coefficients)");
    }
    return 0;
}

int aukit_dev_alloc(aukit_ctx *ctx, size_t nbytes, void **d_out) {
    return aukit_cuda_check(cudaMallocAsync(d_out, nbytes ? nbytes : 16, ctx->stream), "cudaMallocAsync");
}

void aukit_dev_free(aukit_ctx *ctx, void *d) {
    if (d) cudaFreeAsync(d, ctx->stream);
}

// Lua strings / Python bytes are pageable: stage through a pinned buffer in slices so the
// copy runs at PCIe rate and overlaps the next slice's memcpy.
int aukit_upload_into(aukit_ctx *ctx, const void *h, size_t nbytes, void *d) {
    const size_t slice = (size_t)16 << 20;
    if (nbytes == 0) return 0;
    cudaPointerAttributes at{};
    const bool pinned = cudaPointerGetAttributes(&at, h) == cudaSuccess && at.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (pinned || nbytes <= 4096) {
        if (aukit_cuda_check(cudaMemcpyAsync(d, h, nbytes, cudaMemcpyHostToDevice, ctx->stream), "H2D")) return -1;
        if (!pinned) cudaStreamSynchronize(ctx->stream);   // small pageable copies are staged by the driver
        return 0;
    }
    if (!ctx->h_stage) {
        if (aukit_cuda_check(cudaMallocHost(&ctx->h_stage, 2 * slice), "cudaMallocHost")) return -1;
        ctx->h_stage_bytes = 2 * slice;
    }
    cudaEvent_t ev[2];
    cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming);
    int k = 0;
    for (size_t off = 0; off < nbytes; off += slice, k ^= 1) {
        const size_t n = nbytes - off < slice ? nbytes - off : slice;
        char *st = static_cast<char *>(ctx->h_stage) + (size_t)k * slice;
        if (off >= 2 * slice) cudaEventSynchronize(ev[k]);
        memcpy(st, static_cast<const char *>(h) + off, n);
        cudaMemcpyAsync(static_cast<char *>(d) + off, st, n, cudaMemcpyHostToDevice, ctx->stream);
        cudaEventRecord(ev[k], ctx->stream);
    }
    cudaEventSynchronize(ev[0]);
    cudaEventSynchronize(ev[1]);
    cudaEventDestroy(ev[0]);
    cudaEventDestroy(ev[1]);
    return aukit_cuda_check(cudaGetLastError(), "staged H2D") ? -1 : 0;
}

int aukit_upload_bytes(aukit_ctx *ctx, const void *h, size_t nbytes, void **d_out) {
    void *d = nullptr;
    if (aukit_dev_alloc(ctx, nbytes, &d)) return -1;
    if (aukit_upload_into(ctx, h, nbytes, d)) { aukit_dev_free(ctx, d); return -1; }
    *d_out = d;
    return 0;
}

// ------------------------------------------------------------------ Audio handles
int aukit_audio_alloc(aukit_ctx *ctx, int channels, size_t frames, double rate, aukit_audio **out) {
    aukit_audio *a = static_cast<aukit_audio *>(calloc(1, sizeof(aukit_audio)));
    if (!a) return aukit_fail("aukit_cuda: out of host memory");
    a->channels = channels;
    a->frames = frames;
    a->stride = aukit_round_stride(frames ? frames : 1);
    a->sampleRate = rate;
    a->owned = true;
    void *d = nullptr;
    if (aukit_dev_alloc(ctx, a->stride * (size_t)(channels > 0 ? channels : 1) * sizeof(float), &d)) { free(a); return -1; }
    a->data = static_cast<float *>(d);
    *out = a;
    return 0;
}

extern "C" int aukit_cuda_audio_new(aukit_ctx *ctx, int channels, size_t frames, double sampleRate, aukit_audio **out) {
    if (!ctx || !out) return aukit_fail("aukit_cuda: null argument");
    if (channels < 1) return aukit_fail("number outside of range (expected %d to be at least 1)", channels);
    if (aukit_audio_alloc(ctx, channels, frames, sampleRate, out)) return -1;
    return aukit_cuda_check(cudaMemsetAsync((*out)->data, 0, (*out)->stride * (size_t)channels * sizeof(float), ctx->stream),
                            "cudaMemsetAsync");
}

extern "C" int aukit_cuda_audio_wrap(aukit_ctx *ctx, float *d_data, int channels, size_t frames, size_t stride,
                                     double sampleRate, aukit_audio **out) {
    if (!ctx || !out || !d_data) return aukit_fail("aukit_cuda: null argument");
    if (channels < 1) return aukit_fail("aukit_cuda: channels < 1");
    if (channels > 1 && stride < frames) return aukit_fail("aukit_cuda: stride < frames");
    aukit_audio *a = static_cast<aukit_audio *>(calloc(1, sizeof(aukit_audio)));
    if (!a) return aukit_fail("aukit_cuda: out of host memory");
    a->data = d_data; a->channels = channels; a->frames = frames; a->stride = stride;
    a->sampleRate = sampleRate; a->owned = false;
    *out = a;
    return 0;
}

extern "C" void aukit_cuda_audio_free(aukit_ctx *ctx, aukit_audio *a) {
    if (!a) return;
    if (a->block) {                                  // a slice of a shared allocation (batch calls): last one out frees it
        if (--a->block->refs == 0) {
            if (ctx) aukit_dev_free(ctx, a->block->d);
            free(a->block);
        }
    } else if (a->owned && ctx) aukit_dev_free(ctx, a->data);
    free(a->ch_frames);
    free(a);
}

extern "C" int aukit_cuda_audio_channels(const aukit_audio *a) { return a ? a->channels : 0; }
extern "C" size_t aukit_cuda_audio_frames(const aukit_audio *a) { return a ? a->frames : 0; }
extern "C" size_t aukit_cuda_audio_stride(const aukit_audio *a) { return a ? a->stride : 0; }
extern "C" double aukit_cuda_audio_sample_rate(const aukit_audio *a) { return a ? a->sampleRate : 0; }
extern "C" float *aukit_cuda_audio_data(const aukit_audio *a) { return a ? a->data : nullptr; }
extern "C" int aukit_cuda_audio_set_sample_rate(aukit_audio *a, double sampleRate) {
    if (!a) return aukit_fail("aukit_cuda: null argument");
    a->sampleRate = sampleRate;      // a plain field in the reference (effects.speed assigns it, A:3383)
    return 0;
}
extern "C" size_t aukit_cuda_audio_channel_frames(const aukit_audio *a, int channel) {
    if (!a || channel < 0 || channel >= a->channels) return 0;
    return a->ch_frames ? a->ch_frames[channel] : a->frames;
}

extern "C" int aukit_cuda_audio_download(aukit_ctx *ctx, const aukit_audio *a, int channel, size_t first, size_t count,
                                         float *h_out) {
    if (!ctx || !a || (!h_out && count)) return aukit_fail("aukit_cuda: null argument");
    if (channel < 0 || channel >= a->channels) return aukit_fail("aukit_cuda: channel out of range");
    if (first + count > aukit_cuda_audio_channel_frames(a, channel)) return aukit_fail("aukit_cuda: frame range out of bounds");
    if (count)
        AUKIT_CUDA_TRY(cudaMemcpyAsync(h_out, a->data + (size_t)channel * a->stride + first, count * sizeof(float),
                                       cudaMemcpyDeviceToHost, ctx->stream));
    return aukit_cuda_synchronize(ctx);
}

extern "C" int aukit_cuda_audio_upload(aukit_ctx *ctx, aukit_audio *a, int channel, size_t first, size_t count,
                                       const float *h_in) {
    if (!ctx || !a || (!h_in && count)) return aukit_fail("aukit_cuda: null argument");
    if (channel < 0 || channel >= a->channels) return aukit_fail("aukit_cuda: channel out of range");
    if (first + count > a->frames) return aukit_fail("aukit_cuda: frame range out of bounds");
    if (count) {
        AUKIT_CUDA_TRY(cudaMemcpyAsync(a->data + (size_t)channel * a->stride + first, h_in, count * sizeof(float),
                                       cudaMemcpyHostToDevice, ctx->stream));
        AUKIT_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
    return 0;
}

// ------------------------------------------------------------------ loaders
extern "C" int aukit_cuda_pcm(aukit_ctx *ctx, const void *h_data, size_t nbytes, int bitDepth, int dataType,
                              int channels, double sampleRate, int interleaved, int bigEndian, aukit_audio **out) {
    if (!ctx || !out) return aukit_fail("aukit_cuda: null argument");
    // argument validation in the reference's order (A:1058-1064) before any device work
    if (bitDepth != 8 && bitDepth != 16 && bitDepth != 24 && bitDepth != 32)
        return aukit_fail("bad argument #2 (invalid bit depth)");
    if (dataType != AUKIT_SIGNED && dataType != AUKIT_UNSIGNED && dataType != AUKIT_FLOAT)
        return aukit_fail("bad argument #3 (invalid data type)");
    if (dataType == AUKIT_FLOAT && bitDepth != 32) return aukit_fail("bad argument #2 (float audio must have 32-bit depth)");
    if (channels < 1) return aukit_fail("number outside of range (expected %d to be at least 1)", channels);
    if (sampleRate < 1) return aukit_fail("number outside of range (expected %g to be at least 1)", sampleRate);
    const size_t B = (size_t)bitDepth / 8;
    if (nbytes % (B * (size_t)channels)) return aukit_fail("bad argument #1 (uneven amount of data per channel)");
    const size_t frames = nbytes / B / (size_t)channels;
    void *d_in = nullptr;
    if (aukit_upload_bytes(ctx, h_data, nbytes, &d_in)) return -1;
    aukit_audio *a = nullptr;
    if (aukit_audio_alloc(ctx, channels, frames, sampleRate, &a)) { aukit_dev_free(ctx, d_in); return -1; }
    int rc = aukit_cuda_dev_pcm(ctx, d_in, nbytes, bitDepth, dataType, channels, interleaved, bigEndian, a->data, a->stride);
    aukit_dev_free(ctx, d_in);
    if (rc) { aukit_cuda_audio_free(ctx, a); return -1; }
    *out = a;
    return 0;
}

extern "C" int aukit_cuda_g711(aukit_ctx *ctx, const void *h_data, size_t nbytes, int ulaw, int channels,
                               double sampleRate, aukit_audio **out) {
    if (!ctx || !out) return aukit_fail("aukit_cuda: null argument");
    if (channels < 1) return aukit_fail("aukit_cuda: channels < 1");
    const size_t C = (size_t)channels;
    const size_t frames = (nbytes + C - 1) / C;                         // #data[1]
    void *d_in = nullptr;
    if (aukit_upload_bytes(ctx, h_data, nbytes, &d_in)) return -1;
    aukit_audio *a = nullptr;
    if (aukit_audio_alloc(ctx, channels, frames, sampleRate, &a)) { aukit_dev_free(ctx, d_in); return -1; }
    if (nbytes % C) {                                                   // ragged: later channels are one short
        a->ch_frames = static_cast<size_t *>(malloc(sizeof(size_t) * C));
        for (size_t c = 0; c < C; c++) a->ch_frames[c] = nbytes > c ? (nbytes - c + C - 1) / C : 0;
    }
    int rc = aukit_cuda_dev_g711(ctx, d_in, nbytes, ulaw, channels, a->data, a->stride);
    aukit_dev_free(ctx, d_in);
    if (rc) { aukit_cuda_audio_free(ctx, a); return -1; }
    *out = a;
    return 0;
}

extern "C" int aukit_cuda_adpcm(aukit_ctx *ctx, const void *h_data, size_t nbytes, int channels, double sampleRate,
                                int topFirst, int interleaved, const int *predictor, const int *step_index,
                                aukit_audio **out) {
    if (!ctx || !out) return aukit_fail("aukit_cuda: null argument");
    if (channels < 1) return aukit_fail("number outside of range (expected %d to be at least 1)", channels);
    for (int j = 0; j < channels; j++) {
        if (predictor && (predictor[j] < -32768 || predictor[j] > 32767))
            return aukit_fail("number outside of range (expected %d to be within -32768 and 32767)", predictor[j]);
        if (step_index && (step_index[j] < 0 || step_index[j] > 88))
            return aukit_fail("number outside of range (expected %d to be within 0 and 88)", step_index[j]);
    }
    const size_t len = nbytes * 2 / (size_t)channels;                   // A:1231
    void *d_in = nullptr, *d_p = nullptr, *d_i = nullptr;
    auto drop = [&]() { aukit_dev_free(ctx, d_in); aukit_dev_free(ctx, d_p); aukit_dev_free(ctx, d_i); return -1; };
    if (aukit_upload_bytes(ctx, h_data, nbytes, &d_in)) return drop();
    if (predictor && aukit_upload_bytes(ctx, predictor, sizeof(int) * (size_t)channels, &d_p)) return drop();
    if (step_index && aukit_upload_bytes(ctx, step_index, sizeof(int) * (size_t)channels, &d_i)) return drop();
    aukit_audio *a = nullptr;
    if (aukit_audio_alloc(ctx, channels, len, sampleRate, &a)) return drop();
    int rc = len ? aukit_launch_adpcm_stream(ctx, static_cast<const uint8_t *>(d_in), len, channels, topFirst, interleaved,
                                             static_cast<const int *>(d_p), static_cast<const int *>(d_i), a->data, a->stride)
                 : 0;
    aukit_dev_free(ctx, d_in); aukit_dev_free(ctx, d_p); aukit_dev_free(ctx, d_i);
    if (rc) { aukit_cuda_audio_free(ctx, a); return -1; }
    *out = a;
    return 0;
}

extern "C" int aukit_cuda_ima_adpcm_wav(aukit_ctx *ctx, const void *h_data, size_t nbytes, int blockAlign, int channels,
                                        double sampleRate, int dialect, aukit_audio **out) {
    if (!ctx || !out) return aukit_fail("aukit_cuda: null argument");
    if (blockAlign < 1) return aukit_fail("'for' step must be positive");
    if (channels < 1) return aukit_fail("aukit_cuda: channels < 1");
    if (nbytes == 0) return aukit_fail("attempt to index a nil value");
    const size_t frames = aukit_ima_adpcm_wav_frames(nbytes, blockAlign, channels, dialect);
    void *d_in = nullptr;
    if (aukit_upload_bytes(ctx, h_data, nbytes, &d_in)) return -1;
    aukit_audio *a = nullptr;
    if (aukit_audio_alloc(ctx, channels, frames, sampleRate, &a)) { aukit_dev_free(ctx, d_in); return -1; }
    int rc = aukit_cuda_dev_ima_adpcm_wav(ctx, d_in, nbytes, blockAlign, channels, dialect, a->data, a->stride);
    aukit_dev_free(ctx, d_in);
    if (rc) { aukit_cuda_audio_free(ctx, a); return -1; }
    *out = a;
    return 0;
}

extern "C" int aukit_cuda_msadpcm(aukit_ctx *ctx, const void *h_data, size_t nbytes, int blockAlign, int channels,
                                  double sampleRate, const int *coef1, const int *coef2, int ncoef, int dialect,
                                  aukit_audio **out) {
    if (!ctx || !out) return aukit_fail("aukit_cuda: null argument");
    if (blockAlign < 1) return aukit_fail("'for' step must be positive");
    if (sampleRate < 1) return aukit_fail("number outside of range (expected %g to be at least 1)", sampleRate);
    // the reference allocates data = {{}, channels == 2 and {} or nil} (A:1305): an empty string
    // yields an empty Audio with 1 or 2 channels whatever `channels` says
    const int out_ch = channels == 2 ? 2 : 1;
    if (nbytes && dialect == AUKIT_DIALECT_LITERAL && channels != 1 && channels != 2)
        return aukit_fail("Unsupported number of channels: %d", channels);
    const int ch = nbytes ? channels : out_ch;
    const size_t frames = nbytes ? aukit_msadpcm_frames(nbytes, blockAlign, channels) : 0;
    void *d_in = nullptr;
    if (aukit_upload_bytes(ctx, h_data, nbytes, &d_in)) return -1;
    aukit_audio *a = nullptr;
    if (aukit_audio_alloc(ctx, ch, frames, sampleRate, &a)) { aukit_dev_free(ctx, d_in); return -1; }
    int rc = aukit_cuda_dev_msadpcm(ctx, d_in, nbytes, blockAlign, channels, coef1, coef2, ncoef, dialect, a->data, a->stride);
    aukit_dev_free(ctx, d_in);
    if (rc) { aukit_cuda_audio_free(ctx, a); return -1; }
    *out = a;
    return 0;
}

// ------------------------------------------------------------------ aukit.wav, A:1456-1574
static uint32_t rd_u16(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
static int rd_i16(const uint8_t *p) { return (int)(int16_t)rd_u16(p); }
static uint32_t rd_u32(const uint8_t *p) { return rd_u16(p) | (rd_u16(p + 2) << 16); }

extern "C" int aukit_cuda_wav_parse(const void *h_data, size_t nbytes, aukit_wav_info *info) {
    g_err[0] = 0;
    if (!h_data || !info) return aukit_fail("aukit_cuda: null argument");
    const uint8_t *data = static_cast<const uint8_t *>(h_data);
    memset(info, 0, sizeof *info);
    info->format = AUKIT_WAV_NONE;
    // The reference decodes every data chunk where it meets it, with the fmt state seen SO FAR (A:1505-1555); the
    // last data chunk's Audio is returned.  The fmt fields are therefore parsed into `cur` and snapshotted into
    // *info at each data chunk: a fmt chunk that follows the last data chunk changes nothing (it is still
    // validated, as the reference would raise on an unsupported one).
    struct fmt_state { int format, channels, sampleRate, blockAlign, bitDepth, have_fmt, ncoef; int coef1[256], coef2[256]; };
    static thread_local fmt_state cur_s;
    fmt_state *cur = &cur_s;
    memset(cur, 0, sizeof *cur);
    cur->format = AUKIT_WAV_NONE;
    static const uint8_t ksTail[12] = {0x00, 0x00, 0x10, 0x00, 0x80, 0x00, 0x00, 0xaa, 0x00, 0x38, 0x9b, 0x71};
    static const uint8_t dfpwm[16] = {0x3a, 0xc1, 0xfa, 0x38, 0x81, 0x1d, 0x43, 0x61,
                                      0xa4, 0x0d, 0xce, 0x53, 0xca, 0x60, 0x7c, 0xd1};
    if (nbytes < 4) return aukit_fail("data string too short");
    if (memcmp(data, "RIFF", 4)) return aukit_fail("bad argument #1 (not a WAV file)");
    if (nbytes < 12) return aukit_fail("data string too short");
    if (memcmp(data + 8, "WAVE", 4)) return aukit_fail("bad argument #1 (not a WAV file)");
    bool have_data = false;
    size_t pos = 12;
    while (pos < nbytes) {
        if (nbytes - pos < 8) return aukit_fail("data string too short");
        const uint8_t *id = data + pos;
        const size_t size = rd_u32(data + pos + 4);
        pos += 8;                                       // no pad byte after odd chunks (A:1471)
        if (!memcmp(id, "fmt ", 4)) {
            const size_t avail = pos < nbytes ? nbytes - pos : 0;
            const size_t clen = size < avail ? size : avail;
            const uint8_t *ch = data + pos;
            pos += size;
            if (clen < 16) return aukit_fail("data string too short");
            const uint32_t format = rd_u16(ch);
            cur->channels = (int)rd_u16(ch + 2);
            cur->sampleRate = (int)rd_u32(ch + 4);
            cur->blockAlign = (int)rd_u16(ch + 12);
            cur->bitDepth = (int)rd_u16(ch + 14);
            cur->have_fmt = 1;
            switch (format) {
            case 1: cur->format = cur->bitDepth == 8 ? AUKIT_WAV_PCM_UNSIGNED : AUKIT_WAV_PCM_SIGNED; break;
            case 2: {
                cur->format = AUKIT_WAV_MSADPCM;
                if (clen < 22) return aukit_fail("data string too short");
                const int numcoeff = (int)rd_u16(ch + 20);
                if (numcoeff > 256) return aukit_fail("aukit_cuda: more than 256 coefficient pairs");
                for (int i = 0; i < numcoeff; i++) {
                    const size_t o = 22 + 4 * (size_t)i;
                    if (o + 4 > clen) return aukit_fail("data string too short");
                    cur->coef1[i] = rd_i16(ch + o);
                    cur->coef2[i] = rd_i16(ch + o + 2);
                }
                if (numcoeff > 0) cur->ncoef = numcoeff;
                break;
            }
            case 3: cur->format = AUKIT_WAV_FLOAT; break;
            case 6: cur->format = AUKIT_WAV_ALAW; break;
            case 7: cur->format = AUKIT_WAV_ULAW; break;
            case 0x11: cur->format = AUKIT_WAV_ADPCM; break;
            case 0xFFFE: {
                if (clen < 20) return aukit_fail("data string too short");
                cur->bitDepth = (int)rd_u16(ch + 18);
                if (clen < 40) return aukit_fail("unsupported WAV file");
                const uint8_t *u = ch + 24;
                if (!memcmp(u, dfpwm, 16)) { cur->format = AUKIT_WAV_DFPWM; break; }
                if (memcmp(u + 4, ksTail, 12) || u[1] || u[2] || u[3]) return aukit_fail("unsupported WAV file");
                switch (u[0]) {
                case 0x01: cur->format = cur->bitDepth == 8 ? AUKIT_WAV_PCM_UNSIGNED : AUKIT_WAV_PCM_SIGNED; break;
                case 0x02: cur->format = AUKIT_WAV_MSADPCM; break;
                case 0x03: cur->format = AUKIT_WAV_FLOAT; break;
                case 0x06: cur->format = AUKIT_WAV_ALAW; break;
                case 0x07: cur->format = AUKIT_WAV_ULAW; break;
                case 0x11: cur->format = AUKIT_WAV_ADPCM; break;
                default: return aukit_fail("unsupported WAV file");
                }
                break;
            }
            default: return aukit_fail("unsupported WAV file");
            }
        } else if (!memcmp(id, "data", 4)) {
            if (size > nbytes - pos) return aukit_fail("invalid WAV file");
            info->data_off = pos;
            info->data_size = size;
            info->format = cur->format; info->channels = cur->channels; info->sampleRate = cur->sampleRate;
            info->blockAlign = cur->blockAlign; info->bitDepth = cur->bitDepth; info->have_fmt = cur->have_fmt;
            info->ncoef = cur->ncoef;
            memcpy(info->coef1, cur->coef1, sizeof info->coef1);
            memcpy(info->coef2, cur->coef2, sizeof info->coef2);
            have_data = true;
            pos += size;
        } else if (!memcmp(id, "LIST", 4)) {
            if (nbytes - pos < 4) return aukit_fail("data string too short");
            if (!memcmp(data + pos, "INFO", 4)) {
                const size_t e = pos + size;
                pos += 4;
                while (pos < e) {                                       // "!2<c4s4Xh"
                    if (pos > nbytes || nbytes - pos < 8) return aukit_fail("data string too short");
                    const size_t len = rd_u32(data + pos + 4);
                    if (len > nbytes - pos - 8) return aukit_fail("data string too short");
                    if (info->ntags < 64) {
                        memcpy(info->tags[info->ntags].id, data + pos, 4);
                        info->tags[info->ntags].id[4] = 0;
                        info->tags[info->ntags].off = pos + 8;
                        info->tags[info->ntags].len = len;
                        info->ntags++;
                    }
                    pos += 8 + len;
                    if (pos & 1) {
                        if (pos + 1 > nbytes) return aukit_fail("data string too short");
                        pos++;
                    }
                }
            } else pos += size;
        } else pos += size;                                             // "fact" and unknown chunks
    }
    if (!have_data) return aukit_fail("invalid WAV file");
    return 0;
}

extern "C" int aukit_cuda_wav(aukit_ctx *ctx, const void *h_data, size_t nbytes, int head_only, int dialect,
                              aukit_wav_info *info_out, aukit_audio **out) {
    if (!ctx || !out) return aukit_fail("aukit_cuda: null argument");
    aukit_wav_info local;
    aukit_wav_info *info = info_out ? info_out : &local;
    if (aukit_cuda_wav_parse(h_data, nbytes, info)) return -1;
    const uint8_t *payload = static_cast<const uint8_t *>(h_data) + info->data_off;
    const size_t n = info->data_size;
    // aukit.pcm defaults when a data chunk precedes any fmt chunk (channels/sampleRate nil)
    const int ch = info->have_fmt ? info->channels : 1;
    const double rate = info->have_fmt ? (double)info->sampleRate : 48000.0;
    if (head_only) {                                                    // aukit.new(0, channels, sampleRate), A:1508
        if (ch < 1) return aukit_fail("number outside of range (expected %d to be at least 1)", ch);
        if (rate < 1) return aukit_fail("number outside of range (expected %g to be at least 1)", rate);
        return aukit_cuda_audio_new(ctx, ch, 0, rate, out);
    }
    switch (info->format) {
    case AUKIT_WAV_ADPCM: return aukit_cuda_ima_adpcm_wav(ctx, payload, n, info->blockAlign, ch, rate, dialect, out);
    case AUKIT_WAV_MSADPCM:
        return aukit_cuda_msadpcm(ctx, payload, n, info->blockAlign, ch, rate, info->ncoef ? info->coef1 : nullptr,
                                  info->ncoef ? info->coef2 : nullptr, info->ncoef, dialect, out);
    case AUKIT_WAV_ALAW: return aukit_cuda_g711(ctx, payload, n, 0, ch, rate, out);
    case AUKIT_WAV_ULAW: return aukit_cuda_g711(ctx, payload, n, 1, ch, rate, out);
    case AUKIT_WAV_DFPWM: return aukit_fail("aukit_cuda: DFPWM WAV data is outside the accelerated path");
    case AUKIT_WAV_NONE: return aukit_cuda_pcm(ctx, payload, n, 8, AUKIT_SIGNED, 1, 48000.0, 1, 0, out);
    case AUKIT_WAV_PCM_UNSIGNED: return aukit_cuda_pcm(ctx, payload, n, info->bitDepth, AUKIT_UNSIGNED, ch, rate, 1, 0, out);
    case AUKIT_WAV_FLOAT: return aukit_cuda_pcm(ctx, payload, n, info->bitDepth, AUKIT_FLOAT, ch, rate, 1, 0, out);
    default: return aukit_cuda_pcm(ctx, payload, n, info->bitDepth, AUKIT_SIGNED, ch, rate, 1, 0, out);
    }
}

// ------------------------------------------------------------------ transforms
static int ragged_error(const aukit_audio *a) {
    if (a->ch_frames)
        for (int c = 1; c < a->channels; c++)
            if (a->ch_frames[c] != a->frames) return aukit_fail("attempt to perform arithmetic on a nil value");
    return 0;
}

extern "C" int aukit_cuda_resample(aukit_ctx *ctx, const aukit_audio *in, double sampleRate, int interpolation,
                                   aukit_audio **out) {
    if (!ctx || !in || !out) return aukit_fail("aukit_cuda: null argument");
    if (interpolation < 0 || interpolation > 3) return aukit_fail("bad argument #2 (invalid interpolation type)");
    if (ragged_error(in)) return -1;
    const uint64_t n_out = aukit_resample_out_len(in->frames, in->sampleRate, sampleRate);
    aukit_audio *a = nullptr;
    if (aukit_audio_alloc(ctx, in->channels, (size_t)n_out, sampleRate, &a)) return -1;
    if (n_out && aukit_cuda_dev_resample(ctx, in->data, in->stride, in->channels, in->frames, 0, in->frames, in->sampleRate,
                                         sampleRate, interpolation, 0, (size_t)n_out, a->data, a->stride)) {
        aukit_cuda_audio_free(ctx, a);
        return -1;
    }
    *out = a;
    return 0;
}

extern "C" int aukit_cuda_mono(aukit_ctx *ctx, const aukit_audio *in, aukit_audio **out) {
    if (!ctx || !in || !out) return aukit_fail("aukit_cuda: null argument");
    if (ragged_error(in)) return -1;
    aukit_audio *a = nullptr;
    if (aukit_audio_alloc(ctx, 1, in->frames, in->sampleRate, &a)) return -1;
    if (aukit_cuda_dev_mono(ctx, in->data, in->stride, in->channels, in->frames, a->data)) { aukit_cuda_audio_free(ctx, a); return -1; }
    *out = a;
    return 0;
}

extern "C" int aukit_cuda_concat(aukit_ctx *ctx, const aukit_audio *const *parts, int nparts, aukit_audio **out) {
    if (!ctx || !parts || nparts < 1 || !out) return aukit_fail("aukit_cuda: null argument");
    int cn = 0;
    size_t total = 0;
    for (int i = 0; i < nparts; i++) {
        if (!parts[i]) return aukit_fail("bad argument #%d (expected Audio, got nil)", i);
        if (parts[i]->sampleRate != parts[0]->sampleRate)
            return aukit_fail("aukit_cuda: concat of different sample rates needs a resample first (A:702)");
        cn = parts[i]->channels > cn ? parts[i]->channels : cn;
        total += parts[i]->frames;
    }
    aukit_audio *a = nullptr;
    if (aukit_cuda_audio_new(ctx, cn, total, parts[0]->sampleRate, &a)) return -1;   // missing channels stay 0 (A:713)
    size_t pos = 0;
    for (int i = 0; i < nparts; i++) {
        if (parts[i]->frames)
            AUKIT_CUDA_TRY(cudaMemcpy2DAsync(a->data + pos, a->stride * sizeof(float), parts[i]->data,
                                             parts[i]->stride * sizeof(float), parts[i]->frames * sizeof(float),
                                             (size_t)parts[i]->channels, cudaMemcpyDeviceToDevice, ctx->stream));
        pos += parts[i]->frames;
    }
    *out = a;
    return 0;
}

// ------------------------------------------------------------------ effects
extern "C" int aukit_cuda_amplify(aukit_ctx *ctx, aukit_audio *a, double multiplier) {
    if (!ctx || !a) return aukit_fail("aukit_cuda: null argument");
    if (multiplier == 1.0) return 0;
    if (!a->ch_frames) return aukit_cuda_dev_amplify(ctx, a->data, a->stride, a->channels, a->frames, multiplier);
    for (int c = 0; c < a->channels; c++)
        if (aukit_cuda_dev_amplify(ctx, a->data + (size_t)c * a->stride, a->stride, 1, a->ch_frames[c], multiplier)) return -1;
    return 0;
}

extern "C" int aukit_cuda_absmax(aukit_ctx *ctx, const aukit_audio *a, int independent, float *d_max) {
    if (!ctx || !a || !d_max) return aukit_fail("aukit_cuda: null argument");
    if (!a->ch_frames) return aukit_cuda_dev_absmax(ctx, a->data, a->stride, a->channels, a->frames, independent, d_max);
    for (int c = 0; c < a->channels; c++)
        if (aukit_cuda_dev_absmax(ctx, a->data + (size_t)c * a->stride, a->stride, 1, a->ch_frames[c], 0,
                                  d_max + (independent ? c : 0)))
            return -1;
    return 0;
}

extern "C" int aukit_cuda_scale_clamp(aukit_ctx *ctx, aukit_audio *a, double peakAmplitude, int independent,
                                      const float *d_max) {
    if (!ctx || !a || !d_max) return aukit_fail("aukit_cuda: null argument");
    if (!a->ch_frames)
        return aukit_cuda_dev_scale_clamp(ctx, a->data, a->stride, a->channels, a->frames, peakAmplitude, independent, d_max);
    for (int c = 0; c < a->channels; c++)
        if (aukit_cuda_dev_scale_clamp(ctx, a->data + (size_t)c * a->stride, a->stride, 1, a->ch_frames[c], peakAmplitude, 0,
                                       d_max + (independent ? c : 0)))
            return -1;
    return 0;
}

extern "C" int aukit_cuda_normalize(aukit_ctx *ctx, aukit_audio *a, double peakAmplitude, int independent) {
    if (!ctx || !a) return aukit_fail("aukit_cuda: null argument");
    if (a->channels > 1024) return aukit_fail("aukit_cuda: more than 1024 channels");
    const int nmax = independent ? a->channels : 1;
    AUKIT_CUDA_TRY(cudaMemsetAsync(ctx->d_scratch, 0, sizeof(float) * (size_t)nmax, ctx->stream));
    if (aukit_cuda_absmax(ctx, a, independent, ctx->d_scratch)) return -1;
    return aukit_cuda_scale_clamp(ctx, a, peakAmplitude, independent, ctx->d_scratch);
}

// ------------------------------------------------------------------ fused end-to-end call
extern "C" int aukit_cuda_pipeline_host(aukit_ctx *ctx, const aukit_pipeline_desc *p, const void *h_in, size_t nbytes,
                                        double peakAmplitude, float *h_out) {
    if (!ctx || !p || !h_out) return aukit_fail("aukit_cuda: null argument");
    const size_t need = p->in_avail * (size_t)p->channels * (size_t)(p->bitDepth / 8);
    if (nbytes < need) return aukit_fail("aukit_cuda: host buffer smaller than in_avail frames");
    void *d_in = nullptr, *d_out = nullptr;
    if (aukit_upload_bytes(ctx, h_in, need, &d_in)) return -1;
    const int out_ch = p->mono ? 1 : p->channels;
    const size_t stride = aukit_round_stride(p->n_out ? p->n_out : 1);
    if (aukit_dev_alloc(ctx, stride * (size_t)out_ch * sizeof(float), &d_out)) { aukit_dev_free(ctx, d_in); return -1; }
    int rc = aukit_cuda_check(cudaMemsetAsync(ctx->d_scratch, 0, sizeof(float), ctx->stream), "memset");
    if (!rc) rc = aukit_cuda_dev_pipeline_peak(ctx, p, d_in, ctx->d_scratch);
    if (!rc) rc = aukit_cuda_dev_pipeline_apply(ctx, p, d_in, peakAmplitude, ctx->d_scratch, static_cast<float *>(d_out), stride);
    if (!rc && p->n_out)
        rc = aukit_cuda_check(cudaMemcpy2DAsync(h_out, p->n_out * sizeof(float), d_out, stride * sizeof(float),
                                                p->n_out * sizeof(float), (size_t)out_ch, cudaMemcpyDeviceToHost, ctx->stream),
                              "D2H");
    aukit_dev_free(ctx, d_in);
    aukit_dev_free(ctx, d_out);
    if (!rc) rc = aukit_cuda_synchronize(ctx);
    return rc;
}

// auplay's chain (A:1049 -> A:653 -> A:677 -> A:3431) from a host string to a device-resident Audio: what the Lua module's
// cu.preload returns on a one-GPU box (aukit_cuda_group_preload_audio is the several-GPU twin).
extern "C" int aukit_cuda_preload_audio(aukit_ctx *ctx, const aukit_pipeline_desc *p, const void *h_in, size_t nbytes,
                                        double peakAmplitude, aukit_audio **out) {
    if (!ctx || !p || !out) return aukit_fail("aukit_cuda: null argument");
    const size_t need = p->in_avail * (size_t)p->channels * (size_t)(p->bitDepth / 8);
    if (nbytes < need) return aukit_fail("aukit_cuda: host buffer smaller than in_avail frames");
    void *d_in = nullptr;
    if (aukit_upload_bytes(ctx, h_in, need, &d_in)) return -1;
    aukit_audio *a = nullptr;
    if (aukit_audio_alloc(ctx, p->mono ? 1 : p->channels, p->n_out, p->dstRate, &a)) { aukit_dev_free(ctx, d_in); return -1; }
    int rc = aukit_cuda_check(cudaMemsetAsync(ctx->d_scratch, 0, sizeof(float), ctx->stream), "memset");
    if (!rc) rc = aukit_cuda_dev_pipeline_peak(ctx, p, d_in, ctx->d_scratch);
    if (!rc) rc = aukit_cuda_dev_pipeline_apply(ctx, p, d_in, peakAmplitude, ctx->d_scratch, a->data, a->stride);
    aukit_dev_free(ctx, d_in);
    if (!rc) rc = aukit_cuda_synchronize(ctx);
    if (rc) { aukit_cuda_audio_free(ctx, a); return rc; }
    *out = a;
    return 0;
}

// ------------------------------------------------------------------ ADPCM block-range shards
extern "C" int aukit_block_shard(uint64_t nblocks, int world, int rank, uint64_t *first, uint64_t *count) {
    if (!first || !count) return aukit_fail("aukit_cuda: null argument");
    if (world < 1 || rank < 0 || rank >= world) return aukit_fail("aukit_cuda: rank %d outside world %d", rank, world);
    // near-equal contiguous ranges; the __int128 product keeps nblocks * rank exact
    const uint64_t b0 = (uint64_t)((unsigned __int128)nblocks * (unsigned)rank / (unsigned)world);
    const uint64_t b1 = (uint64_t)((unsigned __int128)nblocks * (unsigned)(rank + 1) / (unsigned)world);
    *first = b0;
    *count = b1 - b0;
    return 0;
}

// ------------------------------------------------------------------ clip batches (BASELINE config 3)
extern "C" int aukit_cuda_batch_resample_amplify(aukit_ctx *ctx, const void *const *h_clips, const size_t *clip_bytes,
                                                 const double *srcRates, size_t nclips, int bitDepth, int dataType, int channels,
                                                 int bigEndian, double dstRate, int interpolation, double multiplier,
                                                 aukit_audio **out) {
    if (!ctx || !out || (nclips && (!h_clips || !clip_bytes || !srcRates))) return aukit_fail("aukit_cuda: null argument");
    // argument validation in aukit.pcm's order (A:1058-1064), before any device work
    if (bitDepth != 8 && bitDepth != 16 && bitDepth != 24 && bitDepth != 32) return aukit_fail("bad argument #2 (invalid bit depth)");
    if (dataType != AUKIT_SIGNED && dataType != AUKIT_UNSIGNED && dataType != AUKIT_FLOAT) return aukit_fail("bad argument #3 (invalid data type)");
    if (dataType == AUKIT_FLOAT && bitDepth != 32) return aukit_fail("bad argument #2 (float audio must have 32-bit depth)");
    if (channels < 1) return aukit_fail("number outside of range (expected %d to be at least 1)", channels);
    if (interpolation < 0 || interpolation > 2) return aukit_fail("bad argument #2 (invalid interpolation type)");
    if (nclips == 0) return 0;
    const size_t FB = (size_t)(bitDepth / 8) * (size_t)channels;
    aukit_clip *clips = static_cast<aukit_clip *>(calloc(nclips, sizeof(aukit_clip)));
    if (!clips) return aukit_fail("aukit_cuda: out of host memory");
    size_t in_total = 0;
    for (size_t k = 0; k < nclips; k++) {
        if (clip_bytes[k] % FB) { free(clips); return aukit_fail("bad argument #1 (uneven amount of data per channel)"); }
        if (srcRates[k] < 1) { free(clips); return aukit_fail("number outside of range (expected %g to be at least 1)", srcRates[k]); }
        clips[k].in_offset = in_total;
        clips[k].frames = clip_bytes[k] / FB;
        clips[k].srcRate = srcRates[k];
        in_total += (clip_bytes[k] + 15) & ~(size_t)15;               // 16-byte aligned clip starts
    }
    const uint64_t out_floats = aukit_batch_plan(clips, nclips, channels, dstRate);
    void *d_in = nullptr, *d_out = nullptr;
    aukit_block *blk = static_cast<aukit_block *>(calloc(1, sizeof(aukit_block)));
    int rc = blk ? 0 : aukit_fail("aukit_cuda: out of host memory");
    if (!rc) rc = aukit_dev_alloc(ctx, in_total + 16, &d_in);
    if (!rc) rc = aukit_dev_alloc(ctx, (size_t)out_floats * sizeof(float) + 16, &d_out);
    for (size_t k = 0; k < nclips && !rc; k++)
        rc = aukit_upload_into(ctx, h_clips[k], clip_bytes[k], static_cast<char *>(d_in) + clips[k].in_offset);
    if (!rc) rc = aukit_cuda_dev_batch_resample_amplify(ctx, clips, nclips, bitDepth, dataType, channels, bigEndian, dstRate, interpolation,
                                                        multiplier, d_in, static_cast<float *>(d_out));
    aukit_dev_free(ctx, d_in);
    size_t made = 0;
    for (; made < nclips && !rc; made++) {
        aukit_audio *a = static_cast<aukit_audio *>(calloc(1, sizeof(aukit_audio)));
        if (!a) { rc = aukit_fail("aukit_cuda: out of host memory"); break; }
        a->data = static_cast<float *>(d_out) + clips[made].out_offset;
        a->channels = channels; a->frames = (size_t)clips[made].n_out; a->stride = (size_t)clips[made].out_stride;
        a->sampleRate = dstRate; a->owned = false; a->block = blk;
        out[made] = a;
    }
    if (rc) {
        for (size_t k = 0; k < made; k++) { free(out[k]); out[k] = nullptr; }
        aukit_dev_free(ctx, d_out);
        free(blk);
    } else {
        blk->d = d_out;
        blk->refs = (long)nclips;
    }
    free(clips);
    return rc;
}

// ------------------------------------------------------------------ pipelined preloader
struct aukit_preloader {
    aukit_ctx *ctx;
    int slots, next;
    size_t max_in, max_out;                 // bytes / samples per slot
    cudaStream_t s_in, s_run, s_out;
    struct slot_t {
        void *d_in;
        float *d_out, *d_peak;
        cudaEvent_t in_done, in_free, run_done, out_done;
        aukit_pipeline_desc desc;
        size_t stride;
        bool busy;
    } *slot;
};

// Pinned memory is first touched (and therefore placed) by the allocating thread: run the allocation on
// the CPUs next to the current device (sysfs local_cpulist of its PCI function) so that on a multi-socket
// host every GPU's staging buffers sit on its own NUMA node.  Best effort; the affinity is restored.
static bool gpu_local_cpus(cpu_set_t *set) {
    int dev = 0;
    char bus[32] = "";
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetPCIBusId(bus, sizeof bus, dev) != cudaSuccess) { cudaGetLastError(); return false; }
    for (char *c = bus; *c; c++) *c = (char)tolower(*c);
    char path[128];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/local_cpulist", bus);
    FILE *f = fopen(path, "r");
    if (!f) return false;
    char line[4096] = "";
    const bool ok = fgets(line, sizeof line, f) != nullptr;
    fclose(f);
    if (!ok) return false;
    CPU_ZERO(set);
    int n = 0;
    for (char *tok = strtok(line, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
        int a = 0, b = 0;
        const int k = sscanf(tok, "%d-%d", &a, &b);
        if (k < 1) continue;
        if (k == 1) b = a;
        for (int c = a; c <= b && c < CPU_SETSIZE; c++) { CPU_SET(c, set); n++; }
    }
    return n > 0;
}

extern "C" int aukit_cuda_host_alloc(size_t nbytes, void **out) {
    if (!out) return aukit_fail("aukit_cuda: null argument");
    cpu_set_t old, near, both;
    bool moved = false;
    if (sched_getaffinity(0, sizeof old, &old) == 0 && gpu_local_cpus(&near)) {
        CPU_AND(&both, &old, &near);
        if (CPU_COUNT(&both) > 0 && !CPU_EQUAL(&both, &old)) moved = sched_setaffinity(0, sizeof both, &both) == 0;
    }
    int rc = aukit_cuda_check(cudaMallocHost(out, nbytes ? nbytes : 1), "cudaMallocHost");
    if (!rc && nbytes) memset(*out, 0, nbytes);                      // first touch here, on the local node
    if (moved) sched_setaffinity(0, sizeof old, &old);
    return rc;
}
extern "C" void aukit_cuda_host_free(void *p) { if (p) cudaFreeHost(p); }

extern "C" int aukit_cuda_preloader_create(aukit_ctx *ctx, size_t max_in_bytes, size_t max_out_samples, int slots,
                                           aukit_preloader **out) {
    if (!ctx || !out) return aukit_fail("aukit_cuda: null argument");
    if (slots < 1 || slots > 8) return aukit_fail("aukit_cuda: preloader slots must be 1..8");
    aukit_preloader *pl = new aukit_preloader();
    pl->ctx = ctx; pl->slots = slots; pl->next = 0;
    pl->max_in = (max_in_bytes + 255) & ~(size_t)255;
    pl->max_out = aukit_round_stride(max_out_samples ? max_out_samples : 1);
    pl->slot = new aukit_preloader::slot_t[slots]();
    int rc = 0;
    rc |= aukit_cuda_check(cudaStreamCreateWithFlags(&pl->s_in, cudaStreamNonBlocking), "stream");
    rc |= aukit_cuda_check(cudaStreamCreateWithFlags(&pl->s_run, cudaStreamNonBlocking), "stream");
    rc |= aukit_cuda_check(cudaStreamCreateWithFlags(&pl->s_out, cudaStreamNonBlocking), "stream");
    for (int i = 0; i < slots && !rc; i++) {
        auto &s = pl->slot[i];
        rc |= aukit_cuda_check(cudaMalloc(&s.d_in, pl->max_in + 256), "cudaMalloc");
        rc |= aukit_cuda_check(cudaMalloc(&s.d_out, pl->max_out * sizeof(float)), "cudaMalloc");
        rc |= aukit_cuda_check(cudaMalloc(&s.d_peak, 256), "cudaMalloc");
        for (cudaEvent_t *e : {&s.in_done, &s.in_free, &s.run_done, &s.out_done})
            rc |= aukit_cuda_check(cudaEventCreateWithFlags(e, cudaEventDisableTiming), "event");
    }
    if (rc) { aukit_cuda_preloader_destroy(pl); return -1; }
    *out = pl;
    return 0;
}

extern "C" void aukit_cuda_preloader_destroy(aukit_preloader *pl) {
    if (!pl) return;
    cudaStreamSynchronize(pl->s_in); cudaStreamSynchronize(pl->s_run); cudaStreamSynchronize(pl->s_out);
    for (int i = 0; i < pl->slots; i++) {
        auto &s = pl->slot[i];
        cudaFree(s.d_in); cudaFree(s.d_out); cudaFree(s.d_peak);
        for (cudaEvent_t e : {s.in_done, s.in_free, s.run_done, s.out_done}) if (e) cudaEventDestroy(e);
    }
    cudaStreamDestroy(pl->s_in); cudaStreamDestroy(pl->s_run); cudaStreamDestroy(pl->s_out);
    delete[] pl->slot;
    delete pl;
}

extern "C" float *aukit_cuda_preloader_peak_ptr(aukit_preloader *pl, int slot) {
    return (pl && slot >= 0 && slot < pl->slots) ? pl->slot[slot].d_peak : nullptr;
}
extern "C" void *aukit_cuda_preloader_stream(aukit_preloader *pl) { return pl ? (void *)pl->s_run : nullptr; }

// the passes are enqueued through the context's entry points: point it at the preloader's stream meanwhile
struct stream_swap {
    aukit_ctx *c; cudaStream_t old;
    stream_swap(aukit_ctx *ctx, cudaStream_t s) : c(ctx), old(ctx->stream) { ctx->stream = s; }
    ~stream_swap() { c->stream = old; }
};

extern "C" int aukit_cuda_preloader_begin(aukit_preloader *pl, const aukit_pipeline_desc *p, const void *h_in, size_t nbytes,
                                          int *slot_out) {
    if (!pl || !p || !slot_out) return aukit_fail("aukit_cuda: null argument");
    const size_t need = p->in_avail * (size_t)p->channels * (size_t)(p->bitDepth / 8);
    if (nbytes < need) return aukit_fail("aukit_cuda: host buffer smaller than in_avail frames");
    if (need > pl->max_in) return aukit_fail("aukit_cuda: clip larger than the preloader's input slots");
    const int out_ch = p->mono ? 1 : p->channels;
    const size_t stride = aukit_round_stride(p->n_out ? p->n_out : 1);
    if (stride * (size_t)out_ch > pl->max_out) return aukit_fail("aukit_cuda: clip larger than the preloader's output slots");
    const int k = pl->next;
    auto &s = pl->slot[k];
    if (s.busy) {                                   // begun but never finished
        return aukit_fail("aukit_cuda: preloader slot %d is still between begin and finish", k);
    }
    s.desc = *p; s.stride = stride;
    // a failure below leaves the slot free and the rotation where it was (the caller may retry)
    auto enqueue = [&]() -> int {
        // upload: the slot's input buffer is free once the apply pass that last read it is done
        AUKIT_CUDA_TRY(cudaStreamWaitEvent(pl->s_in, s.in_free, 0));
        if (need) AUKIT_CUDA_TRY(cudaMemcpyAsync(s.d_in, h_in, need, cudaMemcpyHostToDevice, pl->s_in));
        AUKIT_CUDA_TRY(cudaEventRecord(s.in_done, pl->s_in));
        // peak pass
        AUKIT_CUDA_TRY(cudaStreamWaitEvent(pl->s_run, s.in_done, 0));
        AUKIT_CUDA_TRY(cudaMemsetAsync(s.d_peak, 0, sizeof(float), pl->s_run));
        stream_swap sw(pl->ctx, pl->s_run);
        return aukit_cuda_dev_pipeline_peak(pl->ctx, &s.desc, s.d_in, s.d_peak);
    };
    if (enqueue()) return -1;
    s.busy = true;
    pl->next = (k + 1) % pl->slots;
    *slot_out = k;
    return 0;
}

extern "C" int aukit_cuda_preloader_finish(aukit_preloader *pl, int k, double peakAmplitude, float *h_out) {
    if (!pl || !h_out) return aukit_fail("aukit_cuda: null argument");
    if (k < 0 || k >= pl->slots || !pl->slot[k].busy) return aukit_fail("aukit_cuda: preloader slot %d was not begun", k);
    auto &s = pl->slot[k];
    const aukit_pipeline_desc *p = &s.desc;
    const int out_ch = p->mono ? 1 : p->channels;
    // the slot's output buffer is free once its previous download is done
    AUKIT_CUDA_TRY(cudaStreamWaitEvent(pl->s_run, s.out_done, 0));
    {
        stream_swap sw(pl->ctx, pl->s_run);
        if (aukit_cuda_dev_pipeline_apply(pl->ctx, p, s.d_in, peakAmplitude, s.d_peak, s.d_out, s.stride)) { s.busy = false; return -1; }
    }
    AUKIT_CUDA_TRY(cudaEventRecord(s.run_done, pl->s_run));
    AUKIT_CUDA_TRY(cudaEventRecord(s.in_free, pl->s_run));
    AUKIT_CUDA_TRY(cudaStreamWaitEvent(pl->s_out, s.run_done, 0));
    if (p->n_out)
        AUKIT_CUDA_TRY(cudaMemcpy2DAsync(h_out, p->n_out * sizeof(float), s.d_out, s.stride * sizeof(float),
                                         p->n_out * sizeof(float), (size_t)out_ch, cudaMemcpyDeviceToHost, pl->s_out));
    AUKIT_CUDA_TRY(cudaEventRecord(s.out_done, pl->s_out));
    s.busy = false;
    return 0;
}

extern "C" int aukit_cuda_preloader_submit(aukit_preloader *pl, const aukit_pipeline_desc *p, const void *h_in, size_t nbytes,
                                           double peakAmplitude, float *h_out) {
    int k = -1;
    if (aukit_cuda_preloader_begin(pl, p, h_in, nbytes, &k)) return -1;
    return aukit_cuda_preloader_finish(pl, k, peakAmplitude, h_out);
}

extern "C" int aukit_cuda_preloader_drain(aukit_preloader *pl) {
    if (!pl) return aukit_fail("aukit_cuda: null argument");
    AUKIT_CUDA_TRY(cudaStreamSynchronize(pl->s_in));
    AUKIT_CUDA_TRY(cudaStreamSynchronize(pl->s_run));
    AUKIT_CUDA_TRY(cudaStreamSynchronize(pl->s_out));
    stream_swap sw(pl->ctx, pl->s_run);
    return aukit_cuda_synchronize(pl->ctx);          // surfaces device-side status bits
}
