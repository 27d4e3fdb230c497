// comm.cu -- the path's only collective, owned by the library: MAX of a few floats over the ranks that hold the time
// shards of one buffer (effects.normalize on a sharded Audio, A:3431-3459: `max` of A:3438-3443 must be global).
//
// It is 4 bytes per rank, so everything about it is latency: an NCCL all-reduce costs a launch of its own plus a ring /
// tree protocol (measured 18-30 us between the two passes of the fused chain).  Here every rank owns an exchange block
// in ITS device memory, peer-mapped by all other ranks (CUDA IPC between processes, cudaDeviceEnablePeerAccess inside
// one process), and ONE tiny kernel per rank does the whole exchange over NVLink / NVSwitch:
//     thread r < world : atomicMax of my local maxima into rank r's block (remote atomics), __threadfence_system,
//                        then +1 on rank r's arrival counter
//     thread 0         : spins on MY arrival counter until all `world` ranks have arrived, then copies the combined
//                        maxima to where the apply pass reads them.
// No host round trip, no NCCL launch; the kernel sits in stream order between the peak pass and the apply pass.
// Float MAX over non-negative values is exact and order-free, so N-GPU results stay bit-identical to one GPU.
// Blocks are rings of EPOCHS slots (a slot is reused every EPOCHS exchanges and cleared half a ring ahead), so no
// reset can race with a slow peer.  A peer that never arrives trips a 10 s timeout that raises an error instead of
// hanging the GPU.
#include "common.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace {

constexpr int COMM_EPOCHS = 64;
constexpr int COMM_MAXVALS = 16;          // floats per exchange (global normalize: 1; independent: channels <= 16)
constexpr int COMM_MAXWORLD = 64;

struct comm_slot {
    unsigned int max_bits[COMM_MAXVALS];  // non-negative floats order like their bit patterns
    unsigned int arrived;
    unsigned int pad[15];
};
struct comm_block {
    comm_slot slot[COMM_EPOCHS];
    unsigned int error;                   // set by the timeout
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
    const int r = threadIdx.x;
    const unsigned int e = a.epoch % COMM_EPOCHS;
    if (r < a.world) {
        comm_slot *s = &a.peer[r]->slot[e];
        for (int k = 0; k < a.nvals; k++) atomicMax(&s->max_bits[k], __float_as_uint(a.vals[k]));   // NaN never reaches here (fmaxf)
        __threadfence_system();
        atomicAdd(&s->arrived, 1u);
    }
    __syncthreads();
    if (r == 0) {
        comm_block *me = a.peer[a.rank];
        volatile unsigned int *arrived = &me->slot[e].arrived;
        const unsigned long long t0 = globaltimer_ns();
        bool ok = true;
        while (*arrived < (unsigned)a.world) {
            if (globaltimer_ns() - t0 > 10000000000ull) { ok = false; break; }
            __nanosleep(200);
        }
        __threadfence_system();
        if (ok) {
            for (int k = 0; k < a.nvals; k++) a.vals[k] = __uint_as_float(*(volatile unsigned int *)&me->slot[e].max_bits[k]);
        } else {
            me->error = 1u;
            atomicOr(a.d_status, AUKIT_DEVERR_COMM_TIMEOUT);
        }
        // clear the slot half a ring ahead: every rank passed it at least EPOCHS / 2 exchanges ago
        comm_slot *z = &me->slot[(e + COMM_EPOCHS / 2) % COMM_EPOCHS];
        for (int k = 0; k < COMM_MAXVALS; k++) z->max_bits[k] = 0u;
        z->arrived = 0u;
    }
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
