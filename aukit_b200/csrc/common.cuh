// common.cuh -- shared declarations for libaukit_cuda.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#include "../../include/aukit_cuda.h"

#define AUKIT_NUM_SMS 148  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

// Deferred device-side decode errors (checked at aukit_cuda_synchronize / download).
enum : int {
    AUKIT_DEVERR_IMA_INDEX = 1,   // step index > 88 in a block header (A:1213 expect.range)
    AUKIT_DEVERR_MS_PREDICTOR = 2,// predictor index >= #coefficients (A:1311 nil arithmetic)
    AUKIT_DEVERR_COMM_TIMEOUT = 4 // a rank never arrived at the normalize exchange (comm.cu)
};

struct aukit_ctx {
    int device;
    cudaStream_t stream;      // stream kernels are enqueued on
    cudaStream_t own_stream;  // created at init
    uint64_t launches;
    int *d_status;            // device word: OR of AUKIT_DEVERR_*
    int *h_status;            // pinned mirror
    void *h_stage;            // pinned staging buffer for H2D of pageable host strings
    size_t h_stage_bytes;
    float *d_scratch;         // small device scratch (normalize maxima)
    int num_sms;
    // fork/join pair for work that may overlap the main kernel of a pass (the fused pipeline's edge kernels)
    cudaStream_t side_stream;
    cudaEvent_t ev_fork, ev_join;
    int *d_hint;              // device word: epoch of the last fused peak pass that saw the channel clamp act
    int epoch;                // incremented by every aukit_cuda_dev_pipeline_peak call
};

struct aukit_block { void *d; long refs; };   // a device allocation shared by several Audio handles (batch results)

struct aukit_audio {
    float *data;
    int channels;
    size_t frames;   // #data[1]
    size_t stride;   // floats between channel rows (multiple of 32)
    double sampleRate;
    bool owned;
    size_t *ch_frames;  // per-channel lengths when ragged (G.711), else nullptr
    aukit_block *block; // non-null: `data` points into a shared allocation released with its last handle
};

int aukit_fail(const char *fmt, ...);
int aukit_cuda_check(cudaError_t e, const char *what);
#define AUKIT_CUDA_TRY(expr)                                     \
    do {                                                         \
        if (aukit_cuda_check((expr), #expr)) return -1;          \
    } while (0)

static inline size_t aukit_round_stride(size_t frames) { return (frames + 31) / 32 * 32; }

// Allocates an owned Audio (uninitialised samples).
int aukit_audio_alloc(aukit_ctx *ctx, int channels, size_t frames, double rate, aukit_audio **out);
// Uploads host bytes to a temporary device buffer on the ctx stream (freed with aukit_dev_free).
int aukit_upload_bytes(aukit_ctx *ctx, const void *h, size_t nbytes, void **d_out);
// The same copy into existing device memory (pageable sources are staged through the context's pinned slices).
int aukit_upload_into(aukit_ctx *ctx, const void *h, size_t nbytes, void *d);
int aukit_dev_alloc(aukit_ctx *ctx, size_t nbytes, void **d_out);
void aukit_dev_free(aukit_ctx *ctx, void *d);

// Host proof that the 3-operation FMA quotient equals RN(n / r) for every integer n with n / r < 2^max_k
// (resample.cu; used by the resampler and the fused pipeline).
bool aukit_quotient_fma_is_exact(double r, int max_k);

// grid sizing: enough CTAs to cover `work_items` at `per_cta`, capped to waves*num_sms*occupancy
static inline unsigned aukit_grid(size_t work_items, size_t per_cta, size_t cap) {
    size_t g = (work_items + per_cta - 1) / per_cta;
    if (g < 1) g = 1;
    if (cap && g > cap) g = cap;
    return (unsigned)g;
}

// ------------------------------------------------------------------ device helpers
#ifdef __CUDACC__

// clamp of A:228-232: NaN passes through (both comparisons false).
__device__ __forceinline__ float clamp_ref(float v) { return v < -1.0f ? -1.0f : (v > 1.0f ? 1.0f : v); }
__device__ __forceinline__ double clamp_ref(double v) { return v < -1.0 ? -1.0 : (v > 1.0 ? 1.0 : v); }

// 16-bit predictor / sample -> float, A:1133 / A:1255 / A:1312: p / (p < 0 and 32768 or 32767).
// Branch-free: p * 2^-15 is exact, and for p > 0 one FMA adds p / (32767 * 32768); the result is
// bit-identical to (float)((double)p / 32767.0) for every p in [0, 32767] (checked exhaustively
// on the host with exact rationals and on the device by tests/test_gpu_decode.py).
// max(p, 0) * 2^-15 is taken as saturate(p * 2^-15): the clamp rides on the multiply (FMA pipe) instead
// of a separate FMNMX on the half-rate ALU pipe; the constant is scaled by 2^15 to match, so the FMA sees
// the same real product and rounds identically.
__device__ __forceinline__ float s16_to_float(int p) {
    const float lo = __fmul_rn((float)p, 1.0f / 32768.0f);
    return __fmaf_rn(__saturatef(lo), (1.0f / (32767.0f * 32768.0f)) * 32768.0f, lo);
}

// streaming 128-bit accesses: inputs are read once, outputs written once
__device__ __forceinline__ uint4 ldg_stream(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 ldg_stream_f4(const float4 *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream(float4 *p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Non-negative floats order like their bit patterns; NaN never reaches here (fmaxf drops it).
__device__ __forceinline__ void atomic_max_nonneg(float *addr, float v) {
    atomicMax(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v));
}

#endif  // __CUDACC__
