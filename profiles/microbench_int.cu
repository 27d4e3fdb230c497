// micro-benchmark: issue cost and pipe of the integer instructions the ADPCM chains are made of (sm_100a)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mbi profiles/microbench_int.cu && /tmp/mbi
// 8 independent chains per thread, 16 warps per SM (4 per SMSP), clock64 around 4096 x 8 x (ops per step) instructions.
// "cyc/SMSP" = cycles one warp instruction occupies its scheduler's issue slot / pipe (1.0 = full rate, 2.0 = half rate).
// Pairs A+B: if cost(A+B) ~ cost(A) + cost(B) they share a pipe, if ~ max(...) (+ issue) they do not.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(const int *w, int *out, int iters) {
    int a[8], b[8];
    float f[8];
    for (int i = 0; i < 8; i++) { a[i] = w[i] + threadIdx.x; b[i] = w[8 + i] ^ threadIdx.x; f[i] = (float)a[i]; }
    const int m = w[16], sh = w[17] & 31;
    const unsigned um = (unsigned)w[18];
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) a[i] = a[i] * m + b[i];                                            // IMAD
            if (MODE == 1) a[i] = __mulhi(a[i], m) + 1;                                       // IMAD.HI (+ add)
            if (MODE == 2) a[i] = (int)__umulhi((unsigned)a[i], um);                          // IMAD.HI.U32
            if (MODE == 3) a[i] = (a[i] >> sh) ^ b[i];                                        // SHF + LOP3
            if (MODE == 4) a[i] = a[i] >> 8;                                                  // SHF imm (may fold)
            if (MODE == 5) a[i] = min(max(a[i] + b[i], -32768), 32767);                       // IADD + 2 x VIMNMX
            if (MODE == 6) a[i] = __vimin_s32_relu(a[i] + b[i], 65535);                       // IADD + VIMNMX.RELU ?
            if (MODE == 7) { f[i] = (float)a[i]; a[i] = __float_as_int(f[i]) + b[i]; }        // I2FP + IADD
            if (MODE == 8) { a[i] = a[i] * m + b[i]; b[i] = (b[i] >> sh) + 1; }                // IMAD || SHF+IADD
            if (MODE == 9) { a[i] = __mulhi(a[i], m); b[i] = (b[i] >> sh) + 1; }               // IMAD.HI || SHF+IADD
            if (MODE == 10) { a[i] = a[i] * m + b[i]; f[i] = fmaf(f[i], 1.0001f, 0.5f); }      // IMAD || FFMA (same pipe?)
            if (MODE == 11) { a[i] = __mulhi(a[i], m); f[i] = fmaf(f[i], 1.0001f, 0.5f); }     // IMAD.HI || FFMA
            if (MODE == 12) { a[i] = max(a[i], b[i]); b[i] = (b[i] >> sh) + 1; }               // VIMNMX || SHF+IADD (same pipe?)
            if (MODE == 13) { f[i] = (float)b[i]; b[i] = (b[i] >> sh) + 1; }                   // I2FP || SHF+IADD
            if (MODE == 14) { int t; asm("cvt.sat.s16.s32 %0, %1;" : "=r"(t) : "r"(a[i] + b[i])); a[i] = t; }   // IADD + I2I.SAT
            if (MODE == 15) { int t; asm("cvt.sat.s16.s32 %0, %1;" : "=r"(t) : "r"(a[i])); a[i] = t ^ m; b[i] = (b[i] >> sh) + 1; }
            if (MODE == 16) a[i] = (a[i] << sh) >> 28;                                        // SHF.L + SHF.R
            if (MODE == 17) a[i] = (int)((unsigned)a[i] * um) >> 28;                          // IMAD + SHF.R
            if (MODE == 18) a[i] = __vimax3_s32(a[i], b[i], m);                               // VIMNMX3
        }
    }
    long long t1 = clock64();
    int acc = 0;
    for (int i = 0; i < 8; i++) acc += a[i] + b[i] + (int)f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + (int)(t1 - t0);
    if (threadIdx.x == 0 && blockIdx.x == 0)
        printf("mode %2d  %d warps/SM: %7.3f cycles per step per warp -> %6.3f cyc/SMSP per step\n", MODE, blockDim.x / 32,
               (double)(t1 - t0) / (iters * 8.0), (double)(t1 - t0) / (iters * 8.0) / (blockDim.x / 128.0));
}

template <int M> void run(const int *w, int *o) { k<M><<<148, 512>>>(w, o, 4096); cudaDeviceSynchronize(); }

int main() {
    int *w, *o;
    cudaMalloc(&w, 256); cudaMalloc(&o, 148 * 1024 * 4);
    int h[64];
    for (int i = 0; i < 64; i++) h[i] = 1000 + 37 * i;
    h[16] = 16777216 + 3; h[17] = 5; h[18] = 0x3000001;
    cudaMemcpy(w, h, 256, cudaMemcpyHostToDevice);
    run<0>(w, o); run<1>(w, o); run<2>(w, o); run<3>(w, o); run<4>(w, o); run<5>(w, o); run<6>(w, o); run<7>(w, o); run<8>(w, o); run<9>(w, o);
    run<10>(w, o); run<11>(w, o); run<12>(w, o); run<13>(w, o); run<14>(w, o); run<15>(w, o); run<16>(w, o); run<17>(w, o); run<18>(w, o);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
