// comm.cu -- the path's only collective, owned by the library: MAX of a few floats over the ranks that hold the time
// shards of one buffer (effects.normalize on a sharded Audio, A:3431-3459: `max` of A:3438-3443 must be global).
//
// It is 4 bytes per rank, so everything about it is latency: an NCCL all-reduce costs a launch of its own plus a ring /
// tree protocol (measured 18-30 us between the two passes of the fused chain).  Here every rank owns an exchange block
// in ITS device memory, peer-mapped by all other ranks (CUDA IPC between processes, cudaDeviceEnablePeerAccess inside
// one process), and ONE tiny kernel per rank does the whole exchange over NVLink / NVSwitch:
//     thread r < world : ONE 8-byte store {epoch tag, my local max} into slot [my rank] of rank r's block (a posted
//                        write over NVLink; value and validity tag cannot tear), then polls slot [r] of MY block until
//                        rank r's tag for this epoch shows up;
//     after a barrier  : the `world` values are max-combined locally and written where the apply pass reads them.
// No remote atomics, no fences, no host round trip, no NCCL launch; the kernel sits in stream order between the peak
// pass and the apply pass.  Float MAX over non-negative values is exact and order-free, so N-GPU results stay
// bit-identical to one GPU.  Slots form a ring of 4 epochs (a rank can be at most one exchange ahead of the slowest
// one, and the tag is the full epoch number), so nothing is ever cleared.  A peer that never arrives trips a 10 s
// timeout that raises an error instead of hanging the GPU.
#include "common.cuh"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace {

constexpr int COMM_RING = 4;              // exchanges in flight at most: a rank is never more than one ahead of the slowest
constexpr int COMM_MAXVALS = 16;          // floats per exchange (global normalize: 1; independent: channels <= 16)
constexpr int COMM_MAXWORLD = 64;

// word[src][k] = (epoch + 1) << 32 | float bits: value and its validity tag travel in ONE 8-byte store, so a rank's
// contribution is a single posted write per peer -- no remote atomics, no fences, nothing to clear.
struct comm_slot {
    unsigned long long word[COMM_MAXWORLD][COMM_MAXVALS];
};
struct comm_block {
    comm_slot slot[COMM_RING];
};

struct exchange_args {
    comm_block *peer[COMM_MAXWORLD];
    int world, rank, nvals;
    unsigned int epoch;
    float *vals;                          // in: local maxima; out: global maxima
    int *d_status;                        // context status word (deferred errors)
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(64) exchange_max_kernel(exchange_args a) {
    __shared__ float part[COMM_MAXWORLD][COMM_MAXVALS];
    __shared__ int failed;
    const int r = threadIdx.x;
    const unsigned int e = a.epoch % COMM_RING;
    const unsigned long long tag = (unsigned long long)(a.epoch + 1u) << 32;
    if (r == 0) failed = 0;
    __syncthreads();
    if (r < a.world) {
        // push: my maxima into slot [e][my rank] of rank r's block (NVLink posted writes; r == rank is a local store)
        unsigned long long *dst = a.peer[r]->slot[e].word[a.rank];
        for (int k = 0; k < a.nvals; k++)
            asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst + k), "l"(tag | (unsigned long long)__float_as_uint(a.vals[k])) : "memory");
        // pull: rank r's maxima out of MY block, as soon as its tag says this epoch
        const unsigned long long *src = a.peer[a.rank]->slot[e].word[r];
        const unsigned long long t0 = globaltimer_ns();
        for (int k = 0; k < a.nvals; k++) {
            unsigned long long w;
            for (;;) {
                asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(src + k) : "memory");
                if ((w >> 32) == (tag >> 32)) break;
                if (globaltimer_ns() - t0 > 10000000000ull) { failed = 1; w = 0; break; }
                __nanosleep(100);
            }
            part[r][k] = __uint_as_float((unsigned int)w);
        }
    }
    __syncthreads();
    if (r < a.nvals) {
        float m = 0.f;
        for (int q = 0; q < a.world; q++) m = fmaxf(m, part[q][r]);     // non-negative, NaN-free (fmaxf upstream): exact, order-free
        a.vals[r] = m;
    }
    if (r == 0 && failed) atomicOr(a.d_status, AUKIT_DEVERR_COMM_TIMEOUT);
}

}  // namespace

struct aukit_comm {
    aukit_ctx *ctx;
    int world, rank;
    comm_block *mine;                     // cudaMalloc'ed (IPC-exportable)
    comm_block *peer[COMM_MAXWORLD];
    bool ipc_opened[COMM_MAXWORLD];
    bool connected;
    unsigned int epoch;
    float *d_vals;                        // local / combined maxima (device)
};

extern "C" int aukit_cuda_comm_create(aukit_ctx *ctx, int world, int rank, aukit_comm **out) {
    if (!ctx || !out) return aukit_fail("aukit_cuda: null argument");
    if (world < 1 || world > COMM_MAXWORLD || rank < 0 || rank >= world) return aukit_fail("aukit_cuda: rank %d outside world %d (max %d)", rank, world, COMM_MAXWORLD);
    AUKIT_CUDA_TRY(cudaSetDevice(ctx->device));
    aukit_comm *c = static_cast<aukit_comm *>(calloc(1, sizeof(aukit_comm)));
    if (!c) return aukit_fail("aukit_cuda: out of host memory");
    c->ctx = ctx; c->world = world; c->rank = rank;
    if (aukit_cuda_check(cudaMalloc(&c->mine, sizeof(comm_block)), "cudaMalloc") ||
        aukit_cuda_check(cudaMemset(c->mine, 0, sizeof(comm_block)), "cudaMemset") ||
        aukit_cuda_check(cudaMalloc(&c->d_vals, sizeof(float) * COMM_MAXVALS), "cudaMalloc") ||
        aukit_cuda_check(cudaMemset(c->d_vals, 0, sizeof(float) * COMM_MAXVALS), "cudaMemset")) {
        cudaFree(c->mine); cudaFree(c->d_vals); free(c);
        return -1;
    }
    c->peer[rank] = c->mine;
    c->connected = world == 1;
    *out = c;
    return 0;
}

extern "C" size_t aukit_cuda_comm_handle_bytes(void) { return sizeof(cudaIpcMemHandle_t); }

extern "C" int aukit_cuda_comm_handle(aukit_comm *c, void *handle_out) {
    if (!c || !handle_out) return aukit_fail("aukit_cuda: null argument");
    AUKIT_CUDA_TRY(cudaSetDevice(c->ctx->device));
    cudaIpcMemHandle_t h;
    AUKIT_CUDA_TRY(cudaIpcGetMemHandle(&h, c->mine));
    memcpy(handle_out, &h, sizeof h);
    return 0;
}

// One process per GPU: `handles` = the world's handles in rank order (exchanged by the host, e.g. an all-gather).
extern "C" int aukit_cuda_comm_connect(aukit_comm *c, const void *handles) {
    if (!c || !handles) return aukit_fail("aukit_cuda: null argument");
    AUKIT_CUDA_TRY(cudaSetDevice(c->ctx->device));
    for (int r = 0; r < c->world; r++) {
        if (r == c->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char *>(handles) + (size_t)r * sizeof h, sizeof h);
        void *p = nullptr;
        if (aukit_cuda_check(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle (peer exchange block)")) return -1;
        c->peer[r] = static_cast<comm_block *>(p);
        c->ipc_opened[r] = true;
    }
    c->connected = true;
    return 0;
}

// One process driving several GPUs (the Lua module: a single host thread): the peers are plain pointers.
extern "C" int aukit_cuda_comm_connect_local(aukit_comm *const *comms, int world) {
    if (!comms || world < 1 || world > COMM_MAXWORLD) return aukit_fail("aukit_cuda: null argument");
    for (int i = 0; i < world; i++) {
        if (!comms[i] || comms[i]->world != world || comms[i]->rank != i) return aukit_fail("aukit_cuda: comm %d does not belong to this group", i);
        if (aukit_cuda_check(cudaSetDevice(comms[i]->ctx->device), "cudaSetDevice")) return -1;
        for (int j = 0; j < world; j++) {
            if (i != j && comms[i]->ctx->device != comms[j]->ctx->device) {
                int can = 0;
                cudaDeviceCanAccessPeer(&can, comms[i]->ctx->device, comms[j]->ctx->device);
                if (!can) return aukit_fail("aukit_cuda: device %d cannot access device %d", comms[i]->ctx->device, comms[j]->ctx->device);
                const cudaError_t e = cudaDeviceEnablePeerAccess(comms[j]->ctx->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return aukit_cuda_check(e, "cudaDeviceEnablePeerAccess");
                cudaGetLastError();
            }
            comms[i]->peer[j] = comms[j]->mine;
        }
        comms[i]->connected = true;
    }
    return 0;
}

extern "C" void aukit_cuda_comm_destroy(aukit_comm *c) {
    if (!c) return;
    cudaSetDevice(c->ctx->device);
    cudaStreamSynchronize(c->ctx->stream);
    for (int r = 0; r < c->world; r++)
        if (c->ipc_opened[r]) cudaIpcCloseMemHandle(c->peer[r]);
    cudaFree(c->mine);
    cudaFree(c->d_vals);
    free(c);
}

extern "C" float *aukit_cuda_comm_values(aukit_comm *c) { return c ? c->d_vals : nullptr; }

// MAX-combines d_vals[0..nvals) (device floats, non-negative) over the ranks, in place, in stream order on the
// context's stream.  Every rank must call it the same number of times.
extern "C" int aukit_cuda_comm_allreduce_max(aukit_comm *c, float *d_vals, int nvals) {
    if (!c || !d_vals) return aukit_fail("aukit_cuda: null argument");
    if (nvals < 1 || nvals > COMM_MAXVALS) return aukit_fail("aukit_cuda: 1..%d values per exchange", COMM_MAXVALS);
    if (!c->connected) return aukit_fail("aukit_cuda: communicator is not connected");
    if (c->world == 1) return 0;
    exchange_args a{};
    for (int r = 0; r < c->world; r++) a.peer[r] = c->peer[r];
    a.world = c->world; a.rank = c->rank; a.nvals = nvals;
    a.epoch = c->epoch++;
    a.vals = d_vals;
    a.d_status = c->ctx->d_status;
    exchange_max_kernel<<<1, 64, 0, c->ctx->stream>>>(a);
    c->ctx->launches++;
    return aukit_cuda_check(cudaGetLastError(), "exchange_max_kernel launch");
}

// effects.normalize (A:3431) on a time-sharded Audio: local abs-max, exchange, scale + clamp -- one call per rank.
extern "C" int aukit_cuda_comm_normalize(aukit_comm *c, aukit_audio *a, double peakAmplitude, int independent) {
    if (!c || !a) return aukit_fail("aukit_cuda: null argument");
    const int nmax = independent ? a->channels : 1;
    if (nmax > COMM_MAXVALS) return aukit_fail("aukit_cuda: independent normalize over more than %d channels", COMM_MAXVALS);
    aukit_ctx *ctx = c->ctx;
    AUKIT_CUDA_TRY(cudaMemsetAsync(c->d_vals, 0, sizeof(float) * (size_t)nmax, ctx->stream));
    if (aukit_cuda_absmax(ctx, a, independent, c->d_vals)) return -1;
    if (aukit_cuda_comm_allreduce_max(c, c->d_vals, nmax)) return -1;
    return aukit_cuda_scale_clamp(ctx, a, peakAmplitude, independent, c->d_vals);
}

// The fused chain on one rank's time shard: peak pass -> exchange -> apply pass, all enqueued, no host sync.
extern "C" int aukit_cuda_comm_pipeline(aukit_comm *c, const aukit_pipeline_desc *p, const void *d_in, double peakAmplitude,
                                        float *d_out, size_t out_stride) {
    if (!c || !p) return aukit_fail("aukit_cuda: null argument");
    aukit_ctx *ctx = c->ctx;
    AUKIT_CUDA_TRY(cudaMemsetAsync(c->d_vals, 0, sizeof(float), ctx->stream));
    if (aukit_cuda_dev_pipeline_peak(ctx, p, d_in, c->d_vals)) return -1;
    if (aukit_cuda_comm_allreduce_max(c, c->d_vals, 1)) return -1;
    return aukit_cuda_dev_pipeline_apply(ctx, p, d_in, peakAmplitude, c->d_vals, d_out, out_stride);
}

// ------------------------------------------------------------------ one host thread, all GPUs
// The reference is a single Lua coroutine, so the Lua module drives every GPU of the box from one thread: a group is
// one context + one communicator per device, connected through plain peer pointers.
// the group calls hop between devices; the caller's current device is restored when they return
struct device_guard {
    int prev = -1;
    device_guard() { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; }
    ~device_guard() { if (prev >= 0) cudaSetDevice(prev); }
};

struct aukit_group {
    int n;
    aukit_ctx *ctx[COMM_MAXWORLD];
    aukit_comm *comm[COMM_MAXWORLD];
};

extern "C" void aukit_cuda_group_destroy(aukit_group *g) {
    if (!g) return;
    device_guard guard;
    for (int i = 0; i < g->n; i++) {
        if (g->comm[i]) aukit_cuda_comm_destroy(g->comm[i]);
        if (g->ctx[i]) aukit_cuda_shutdown(g->ctx[i]);
    }
    free(g);
}

// devices == NULL: devices 0 .. ndev-1 (ndev <= 0: every visible device).  The same device may be listed more than once
// (each entry gets its own context and stream): that is how the exchange is exercised on a one-GPU box.
extern "C" int aukit_cuda_group_create(const int *devices, int ndev, aukit_group **out) {
    if (!out) return aukit_fail("aukit_cuda: null argument");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
        return aukit_fail("aukit_cuda: no CUDA device available; there is no CPU fallback");
    if (ndev <= 0) { ndev = count; devices = nullptr; }
    if (ndev > COMM_MAXWORLD) return aukit_fail("aukit_cuda: at most %d shards per group", COMM_MAXWORLD);
    device_guard guard;
    aukit_group *g = static_cast<aukit_group *>(calloc(1, sizeof(aukit_group)));
    if (!g) return aukit_fail("aukit_cuda: out of host memory");
    g->n = ndev;
    for (int i = 0; i < ndev; i++) {
        const int dev = devices ? devices[i] : i;
        if (aukit_cuda_init(dev, &g->ctx[i]) || aukit_cuda_comm_create(g->ctx[i], ndev, i, &g->comm[i])) { aukit_cuda_group_destroy(g); return -1; }
    }
    if (aukit_cuda_comm_connect_local(g->comm, ndev)) { aukit_cuda_group_destroy(g); return -1; }
    *out = g;
    return 0;
}

extern "C" int aukit_cuda_group_size(const aukit_group *g) { return g ? g->n : 0; }
extern "C" aukit_ctx *aukit_cuda_group_ctx(aukit_group *g, int i) { return (g && i >= 0 && i < g->n) ? g->ctx[i] : nullptr; }
extern "C" aukit_comm *aukit_cuda_group_comm(aukit_group *g, int i) { return (g && i >= 0 && i < g->n) ? g->comm[i] : nullptr; }

// effects.normalize (A:3431) on an Audio whose time shards live one per group member: shards[i] belongs to context i.
extern "C" int aukit_cuda_group_normalize(aukit_group *g, aukit_audio *const *shards, double peakAmplitude, int independent) {
    if (!g || !shards) return aukit_fail("aukit_cuda: null argument");
    device_guard guard;
    // Three sweeps -- every abs-max, every exchange, every scale -- NOT member by member: between the first and the
    // last exchange launch nothing else may be launched for the first time.  (CUDA loads a kernel's code lazily at its
    // first launch and may wait for the device to drain to do so; an exchange kernel that is already spinning for a
    // peer whose own exchange is still behind that load would wait for its full timeout.)
    int nmax = 1;
    for (int i = 0; i < g->n; i++) {
        if (!shards[i]) return aukit_fail("aukit_cuda: null shard");
        nmax = independent ? shards[i]->channels : 1;
        if (nmax > COMM_MAXVALS) return aukit_fail("aukit_cuda: independent normalize over more than %d channels", COMM_MAXVALS);
        aukit_ctx *ctx = g->ctx[i];
        if (aukit_cuda_check(cudaSetDevice(ctx->device), "cudaSetDevice")) return -1;
        AUKIT_CUDA_TRY(cudaMemsetAsync(g->comm[i]->d_vals, 0, sizeof(float) * (size_t)nmax, ctx->stream));
        if (aukit_cuda_absmax(ctx, shards[i], independent, g->comm[i]->d_vals)) return -1;
    }
    for (int i = 0; i < g->n; i++) {
        if (aukit_cuda_check(cudaSetDevice(g->ctx[i]->device), "cudaSetDevice")) return -1;
        if (aukit_cuda_comm_allreduce_max(g->comm[i], g->comm[i]->d_vals, nmax)) return -1;
    }
    for (int i = 0; i < g->n; i++) {
        if (aukit_cuda_check(cudaSetDevice(g->ctx[i]->device), "cudaSetDevice")) return -1;
        if (aukit_cuda_scale_clamp(g->ctx[i], shards[i], peakAmplitude, independent, g->comm[i]->d_vals)) return -1;
    }
    return 0;
}

// auplay's chain (aukit.pcm -> Audio:resample -> [Audio:mono] -> effects.normalize; A:1049, A:653, A:677, A:3431) on ONE
// host buffer, time-sharded over the group: contiguous output ranges on whole warp tiles, each shard's input window with
// its interpolation halo, local peak passes, the MAX exchange, local apply passes, results gathered into
// h_out[c * n_out + i].  `whole` describes the unsharded call (in_first = 0, in_avail = n_in_total, out_first = 0,
// n_out = floor(n_in_total * ratio)).  Bit-identical to the same call on one GPU.
// dst[c * dst_pitch + i]: host memory (kind = DeviceToHost) or device memory of any device (kind = Default: peer copies)
static int group_preload_into(aukit_group *g, const aukit_pipeline_desc *whole, const void *h_in, size_t nbytes,
                              double peakAmplitude, float *dst, size_t dst_pitch, cudaMemcpyKind kind, int dst_device) {
    if (!g || !whole || !dst) return aukit_fail("aukit_cuda: null argument");
    device_guard guard;
    const size_t FB = (size_t)whole->channels * (size_t)(whole->bitDepth / 8);
    if (FB == 0 || nbytes < whole->n_in_total * FB) return aukit_fail("aukit_cuda: host buffer smaller than n_in_total frames");
    const uint64_t n_out = aukit_resample_out_len(whole->n_in_total, whole->srcRate, whole->dstRate);
    const int out_ch = whole->mono ? 1 : whole->channels;
    const int W = g->n;
    // warp-tile aligned shard boundaries (see aukit_b200/sharding.py: shard_alignment)
    uint64_t align = 4;
    if (whole->srcRate == floor(whole->srcRate) && whole->dstRate == floor(whole->dstRate) && whole->srcRate >= 1 && whole->dstRate >= 1) {
        long long x = (long long)whole->srcRate, y = (long long)whole->dstRate;
        while (y) { const long long r = x % y; x = y; y = r; }
        const long long L = (long long)whole->dstRate / x;
        if (L > 1 && L <= 512) align = 32ull * (uint64_t)L;
    }
    if (n_out / (uint64_t)W < 4 * align) align = 4;
    struct part { aukit_pipeline_desc d; void *d_in; float *d_out; size_t stride; };
    part parts[COMM_MAXWORLD];
    memset(parts, 0, sizeof parts);
    int rc = 0;
    // 1. upload every window and enqueue every peak pass (pageable sources are staged synchronously: do it before any
    //    exchange kernel starts to wait)
    for (int i = 0; i < W && !rc; i++) {
        aukit_ctx *ctx = g->ctx[i];
        part &p = parts[i];
        const uint64_t o0 = n_out * (uint64_t)i / (uint64_t)W / align * align;
        const uint64_t o1 = i == W - 1 ? n_out : n_out * (uint64_t)(i + 1) / (uint64_t)W / align * align;
        p.d = *whole;
        p.d.out_first = o0; p.d.n_out = (size_t)(o1 - o0);
        uint64_t f = 0, c = 0;
        if (o1 > o0 && aukit_resample_window(whole->n_in_total, whole->srcRate, whole->dstRate, whole->interpolation, o0, o1 - o0, &f, &c)) { rc = -1; break; }
        uint64_t first = f, end = f + c;
        if (c) {                                                       // the slack the bulk copies of the run kernels want
            if (first >= 8) first = (first - 8) / 4 * 4;
            end = end + 8 < whole->n_in_total ? end + 8 : whole->n_in_total;
        }
        p.d.in_first = first; p.d.in_avail = (size_t)(end - first);
        p.stride = aukit_round_stride(p.d.n_out ? p.d.n_out : 1);
        if ((rc = aukit_cuda_check(cudaSetDevice(ctx->device), "cudaSetDevice"))) break;
        if ((rc = aukit_upload_bytes(ctx, static_cast<const char *>(h_in) + first * FB, p.d.in_avail * FB, &p.d_in))) break;
        void *o = nullptr;
        if ((rc = aukit_dev_alloc(ctx, p.stride * (size_t)out_ch * sizeof(float), &o))) break;
        p.d_out = static_cast<float *>(o);
        float *mx = aukit_cuda_comm_values(g->comm[i]);
        if ((rc = aukit_cuda_check(cudaMemsetAsync(mx, 0, sizeof(float), ctx->stream), "memset"))) break;
        rc = aukit_cuda_dev_pipeline_peak(ctx, &p.d, p.d_in, mx);
    }
    // 2. the exchange, 3. apply passes and downloads
    for (int i = 0; i < W && !rc; i++) {
        if ((rc = aukit_cuda_check(cudaSetDevice(g->ctx[i]->device), "cudaSetDevice"))) break;
        rc = aukit_cuda_comm_allreduce_max(g->comm[i], aukit_cuda_comm_values(g->comm[i]), 1);
    }
    for (int i = 0; i < W && !rc; i++) {
        aukit_ctx *ctx = g->ctx[i];
        part &p = parts[i];
        if ((rc = aukit_cuda_check(cudaSetDevice(ctx->device), "cudaSetDevice"))) break;
        if ((rc = aukit_cuda_dev_pipeline_apply(ctx, &p.d, p.d_in, peakAmplitude, aukit_cuda_comm_values(g->comm[i]), p.d_out, p.stride))) break;
        if (p.d.n_out && kind == cudaMemcpyDeviceToHost) {
            rc = aukit_cuda_check(cudaMemcpy2DAsync(dst + p.d.out_first, dst_pitch * sizeof(float), p.d_out, p.stride * sizeof(float),
                                                    p.d.n_out * sizeof(float), (size_t)out_ch, kind, ctx->stream), "gather");
        } else if (p.d.n_out) {
            // device to device, possibly across devices: row by row (a 2-D copy between two devices is not accepted)
            for (int c = 0; c < out_ch && !rc; c++)
                rc = aukit_cuda_check(cudaMemcpyPeerAsync(dst + (size_t)c * dst_pitch + p.d.out_first, dst_device, p.d_out + (size_t)c * p.stride,
                                                          ctx->device, p.d.n_out * sizeof(float), ctx->stream), "gather");
        }
    }
    for (int i = 0; i < W; i++) {
        cudaSetDevice(g->ctx[i]->device);
        aukit_dev_free(g->ctx[i], parts[i].d_in);
        aukit_dev_free(g->ctx[i], parts[i].d_out);
        const int s = aukit_cuda_synchronize(g->ctx[i]);
        if (!rc && s) rc = s;
    }
    return rc;
}

extern "C" int aukit_cuda_group_preload(aukit_group *g, const aukit_pipeline_desc *whole, const void *h_in, size_t nbytes,
                                        double peakAmplitude, float *h_out) {
    if (!whole) return aukit_fail("aukit_cuda: null argument");
    return group_preload_into(g, whole, h_in, nbytes, peakAmplitude, h_out,
                              (size_t)aukit_resample_out_len(whole->n_in_total, whole->srcRate, whole->dstRate), cudaMemcpyDeviceToHost, -1);
}

// The same, gathered into ONE device-resident Audio that belongs to `owner` (any context, normally the caller's own on
// device 0): the shards travel device to device.  This is what the Lua module's cu.preload returns on a box with several GPUs.
extern "C" int aukit_cuda_group_preload_audio(aukit_group *g, aukit_ctx *owner, const aukit_pipeline_desc *whole, const void *h_in,
                                              size_t nbytes, double peakAmplitude, aukit_audio **out) {
    if (!g || !owner || !whole || !out) return aukit_fail("aukit_cuda: null argument");
    const uint64_t n_out = aukit_resample_out_len(whole->n_in_total, whole->srcRate, whole->dstRate);
    aukit_audio *a = nullptr;
    {
        device_guard guard;
        if (aukit_cuda_check(cudaSetDevice(owner->device), "cudaSetDevice")) return -1;
        if (aukit_audio_alloc(owner, whole->mono ? 1 : whole->channels, (size_t)n_out, whole->dstRate, &a)) return -1;
        if (aukit_cuda_synchronize(owner)) { aukit_cuda_audio_free(owner, a); return -1; }   // the allocation is stream-ordered on owner's stream
    }
    // Audio buffers come from the devices' stream-ordered pools, which peers cannot touch by default: open the owner's pool
    // to every member device (and the members' pools to the owner) once per call; the copies themselves are peer copies
    for (int i = 0; i < g->n; i++) {
        const int d = g->ctx[i]->device;
        if (d == owner->device) continue;
        cudaMemPool_t po, pm;
        if (cudaDeviceGetDefaultMemPool(&po, owner->device) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pm, d) == cudaSuccess) {
            cudaMemAccessDesc to_member{}, to_owner{};
            to_member.location.type = cudaMemLocationTypeDevice; to_member.location.id = d; to_member.flags = cudaMemAccessFlagsProtReadWrite;
            to_owner.location.type = cudaMemLocationTypeDevice; to_owner.location.id = owner->device; to_owner.flags = cudaMemAccessFlagsProtReadWrite;
            cudaMemPoolSetAccess(po, &to_member, 1);
            cudaMemPoolSetAccess(pm, &to_owner, 1);
        }
        cudaGetLastError();                                             // without peer capability the peer copy below stages through the host
    }
    const int rc = group_preload_into(g, whole, h_in, nbytes, peakAmplitude, a->data, a->stride, cudaMemcpyDefault, owner->device);
    if (rc) { device_guard guard; cudaSetDevice(owner->device); aukit_cuda_audio_free(owner, a); return rc; }
    *out = a;
    return 0;
}
