// micro-benchmark: FFMA2 issue cost with a register scalar vs a constant-bank scalar vs scalar FFMA
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f32x2;
__constant__ float cW[64];
__device__ __forceinline__ f32x2 pack2(float x, float y) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y)); return r; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
template <int MODE>
__global__ void k(const float *w, float *out, int iters) {
    f32x2 a[8];
    float s[8];
    for (int i = 0; i < 8; i++) { a[i] = pack2(threadIdx.x * 0.001f + i, i * 0.5f); s[i] = threadIdx.x * 0.002f + i; }
    float wr[8];
    for (int i = 0; i < 8; i++) wr[i] = w[i + (threadIdx.x & 1)];
    f32x2 x = pack2(1.0f + threadIdx.x, 2.0f);
    float xs = 1.0f + threadIdx.x, ys = 0.5f;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) a[i] = fma2(x, pack2(wr[i], wr[i]), a[i]);            // register scalar
            if (MODE == 1) a[i] = fma2(x, pack2(cW[i], cW[i]), a[i]);            // constant scalar
            if (MODE == 2) { s[i] = fmaf(xs, wr[i], s[i]); }                     // scalar FFMA, register
            if (MODE == 3) { s[i] = fmaf(xs, cW[i], s[i]); }                     // scalar FFMA, constant
            if (MODE == 4) { s[i] = fmaf(xs, wr[i], s[i]); a[i] = fma2(x, pack2(wr[i], wr[i]), a[i]); }
        }
    }
    long long t1 = clock64();
    float acc = 0;
    for (int i = 0; i < 8; i++) { float p, q; asm("mov.b64 {%0, %1}, %2;" : "=f"(p), "=f"(q) : "l"(a[i])); acc += p + q + s[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + (float)(t1 - t0);
    if (threadIdx.x == 0 && blockIdx.x == 0) printf("mode %d warps/SM %d: %.3f cycles per instruction-slot per warp-group\n", MODE, blockDim.x / 32, (double)(t1 - t0) / (iters * 8.0));
}
int main() {
    float *w, *o; cudaMalloc(&w, 256); cudaMalloc(&o, 148 * 1024 * 4);
    float h[64]; for (int i = 0; i < 64; i++) h[i] = 0.001f * i; cudaMemcpy(w, h, 256, cudaMemcpyHostToDevice); cudaMemcpyToSymbol(cW, h, 256);
    for (int th = 128; th <= 512; th *= 2) {
        k<0><<<148, th>>>(w, o, 4096); cudaDeviceSynchronize();
        k<1><<<148, th>>>(w, o, 4096); cudaDeviceSynchronize();
        k<2><<<148, th>>>(w, o, 4096); cudaDeviceSynchronize();
        k<3><<<148, th>>>(w, o, 4096); cudaDeviceSynchronize();
        k<4><<<148, th>>>(w, o, 4096); cudaDeviceSynchronize();
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
