#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    unsigned long long ra, rb, rc, rd;
    ra = *reinterpret_cast<unsigned long long*>(&a); rb = *reinterpret_cast<unsigned long long*>(&b); rc = *reinterpret_cast<unsigned long long*>(&c);
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    unsigned long long ra, rb, rd;
    ra = *reinterpret_cast<unsigned long long*>(&a); rb = *reinterpret_cast<unsigned long long*>(&b);
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2*>(&rd);
}
template <int MODE>
__global__ void k(float2* out, int iters, float s) {
    float2 a[8];
    for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x + i, threadIdx.x * 0.5f + i);
    float2 b = make_float2(s, s * 1.0001f), c = make_float2(0.001f, 0.002f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) { a[i].x = fmaf(a[i].x, b.x, c.x); a[i].y = fmaf(a[i].y, b.y, c.y); }
                if (MODE == 1) a[i] = fma2(a[i], b, c);
                if (MODE == 2) { a[i].x = a[i].x + c.x; a[i].y = a[i].y + c.y; }
                if (MODE == 3) a[i] = add2(a[i], c);
            }
        }
    }
    float2 r = make_float2(0, 0);
    for (int i = 0; i < 8; ++i) { r.x += a[i].x; r.y += a[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> void run(const char* name) {
    float2* out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(float2));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 4096;
    k<MODE><<<148 * 4, 512>>>(out, 16, 1.0f);
    cudaEventRecord(e0);
    k<MODE><<<148 * 4, 512>>>(out, iters, 1.0f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = 148.0 * 4 * 512 * (double)iters * 64 * 2;  // scalar fp ops (fma = 1 op)
    printf("%s: %.3f ms  %.1f Gop/s  per SM per clk @1.9GHz: %.1f\n", name, ms, ops / ms / 1e6, ops / ms / 1e6 / 148 / 1.9);
}
int main() { run<0>("FFMA"); run<1>("FFMA2"); run<2>("FADD"); run<3>("FADD2"); return 0; }
