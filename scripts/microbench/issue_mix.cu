// What can issue next to packed fp32x2 math on sm_100a, and what does mixing cost?  Each mode runs a loop body of independent
// instructions per warp; the kernel time gives cycles per body per SM (16 warps/SM unless noted).  SASS checked with cuobjdump.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_mix issue_mix.cu ; run on one B200.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ int imad(int a, int b, int c) { int d; asm volatile("mad.lo.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ int iadd(int a, int b) { int d; asm volatile("add.s32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ u64 lds64(unsigned addr) { u64 d; asm volatile("ld.shared.b64 %0, [%1];" : "=l"(d) : "r"(addr) : "memory"); return d; }
__device__ __forceinline__ void sts64(unsigned addr, u64 v) { asm volatile("st.shared.b64 [%0], %1;" ::"r"(addr), "l"(v) : "memory"); }

// modes: 0 16xFFMA2 | 1 16xFFMA2 + 16xFFMA(scalar, independent) | 2 16xFFMA2 + 16xIMAD | 3 16xFFMA2 + 16xIADD
//        4 16xLDS.64 (consumed by 16 FADD2) | 5 16xLDS.64 + 16 FADD2 + 16 more FFMA2 | 6 32xFFMA scalar + 16 LDS.64 consumed by 32 FADD
//        7 16xFFMA2 then 16xFFMA scalar (grouped, one switch per body) | 8 16 STS.64 + 16 FFMA2 | 9 48 FFMA2 + 16 LDS.64(consumed by 16 of them)
template <int MODE>
__global__ void __launch_bounds__(512) k(u64* out, int iters, float s, int zero) {
    __shared__ u64 sm[32 * 66 + 64];
    for (int i = threadIdx.x; i < 32 * 66 + 64; i += blockDim.x) sm[i] = (u64)i * 0x3f8000003f800000ull;
    __syncthreads();
    u64 a[16], e[16];
    float f[16];
    int n[16];
    for (int i = 0; i < 16; ++i) { a[i] = (u64)(threadIdx.x + i) << 20; e[i] = a[i] + 7; f[i] = threadIdx.x + i; n[i] = threadIdx.x * i; }
    const u64 b = 0x3f8000003f800001ull, c = 0x3a83126f3a83126full;
    const unsigned p = (unsigned)__cvta_generic_to_shared(sm + (threadIdx.x & 31) * 65 + zero);   // lane stride 520 B: conflict-free 64-bit
    const unsigned p0 = p;
    for (int it = 0; it < iters; ++it) {
        const unsigned p = p0 + ((it & 1) << 7);
        if (MODE == 7) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fma2(a[i], b, c);
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = ffma(f[i], s, s);
        } else if (MODE == 4 || MODE == 5 || MODE == 9) {
            u64 v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = lds64(p + 8 * i);
            if (MODE == 5 || MODE == 9) {
#pragma unroll
                for (int i = 0; i < 16; ++i) e[i] = fma2(e[i], b, c);
            }
            if (MODE == 9) {
#pragma unroll
                for (int i = 0; i < 16; ++i) e[i] = fma2(e[i], b, c);
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = add2(a[i], v[i]);
        } else if (MODE == 6) {
            u64 v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = lds64(p + 8 * i);
#pragma unroll
            for (int i = 0; i < 16; ++i) { f[i] = ffma(f[i], s, s); n[i] = __float_as_int(ffma(__int_as_float(n[i]), s, s)); }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                float lo = __uint_as_float((unsigned)v[i]), hi = __uint_as_float((unsigned)(v[i] >> 32));
                float x = __uint_as_float((unsigned)a[i]), y = __uint_as_float((unsigned)(a[i] >> 32));
                x = ffma(lo, 1.0f, x); y = ffma(hi, 1.0f, y);
                a[i] = (u64)__float_as_uint(x) | ((u64)__float_as_uint(y) << 32);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                a[i] = fma2(a[i], b, c);
                if (MODE == 1) f[i] = ffma(f[i], s, s);
                if (MODE == 2) n[i] = imad(n[i], it, it);
                if (MODE == 3) n[i] = iadd(n[i], it);
                if (MODE == 8) sts64(p + 8 * i, a[(i + 8) & 15]);
            }
        }
    }
    u64 r = 0;
    for (int i = 0; i < 16; ++i) r += a[i] + e[i] + (u64)__float_as_uint(f[i]) + (u64)n[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> void run(const char* name, int warps_per_sm) {
    u64* out; cudaMalloc(&out, (148 * 1024 + 8) * sizeof(u64));
    const int iters = 20000, threads = warps_per_sm * 32;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148, threads>>>(out, 16, 1.0f, 0);
    cudaEventRecord(e0);
    k<MODE><<<148, threads>>>(out, iters, 1.0f, 0);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-58s warps/SM %2d: %.3f ms = %6.1f cycles per body-round per SM @1.9GHz\n", name, warps_per_sm, ms, ms * 1.9e6 / iters);
    cudaFree(out);
}
int main() {
    for (int w : {16, 8}) {
        run<0>("16 FFMA2", w);
        run<1>("16 FFMA2 + 16 FFMA interleaved", w);
        run<7>("16 FFMA2 then 16 FFMA grouped", w);
        run<2>("16 FFMA2 + 16 IMAD interleaved", w);
        run<3>("16 FFMA2 + 16 IADD interleaved", w);
        run<4>("16 LDS.64 -> 16 FADD2", w);
        run<5>("16 LDS.64 + 16 FFMA2 -> 16 FADD2", w);
        run<9>("16 LDS.64 + 32 FFMA2 -> 16 FADD2", w);
        run<6>("16 LDS.64 + 32 FFMA -> 32 FFMA (all scalar)", w);
        run<8>("16 FFMA2 + 16 STS.64 interleaved", w);
    }
    return 0;
}
