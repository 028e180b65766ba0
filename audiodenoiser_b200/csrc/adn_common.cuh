// Shared helpers for libadn_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "../../include/adn_b200.h"

namespace adn {

void set_last_cuda_error(cudaError_t e);   // abi.cu (thread-local)
int check_device();                        // abi.cu: ADN_OK iff current device is sm_100
// conv_c1_tc.cu: first-layer conv (Cin = 1) as a TF32 tcgen05 implicit GEMM; ADN_ERR_ARG when the output cannot take a TMA store
int conv3x3_c1_tc(const float* x, int n, int h, int w, const float* weight, const float* scale, const float* shift, float relu_floor,
                  void* out, cudaStream_t stream);

#define ADN_CUDA_TRY(expr)                                  \
    do {                                                    \
        cudaError_t _e = (expr);                            \
        if (_e != cudaSuccess) {                            \
            ::adn::set_last_cuda_error(_e);                 \
            return ADN_ERR_CUDA;                            \
        }                                                   \
    } while (0)

#define ADN_LAUNCH_CHECK() ADN_CUDA_TRY(cudaGetLastError())

inline int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return dev;
}

inline int num_sms() {
    static int cache[64] = {0};
    const int dev = current_device();
    if (dev >= 0 && dev < 64 && cache[dev]) return cache[dev];
    int n = 148;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (dev >= 0 && dev < 64) cache[dev] = n;
    return n;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device): keeps launches free of non-stream calls so
// they can be captured into CUDA graphs.  `flags` is a per-kernel static array of 64 bytes.
template <typename K>
inline cudaError_t ensure_dyn_smem(K kernel, int bytes, unsigned char* flags) {
    const int dev = current_device();
    if (dev >= 0 && dev < 64 && flags[dev]) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess && dev >= 0 && dev < 64) flags[dev] = 1;
    return e;
}

// ---------------------------------------------------------------- packed fp32x2 arithmetic (Blackwell FADD2 / FMUL2 / FFMA2)
// A complex number IS a float2, so every complex add / scale is one packed instruction.  The packed instructions run at the
// same lane rate as their scalar forms (scripts/microbench/fp32x2_rate.cu) and -- round-2 measurement, scripts/microbench/
// issue_mix.cu -- hold the ISSUE PORT for two cycles as well: nothing co-issues beside them, so they halve the instruction count,
// not the issue slots.  What they do buy: operand modifiers (half swap, per-half negation, scalar broadcast) make the x(-i)
// rotations and the real-by-complex products of an FFT free (ptxas folds make_float2(a.y, -a.x) style operands into the instruction).
typedef unsigned long long pk64;
__device__ __forceinline__ pk64 pk_bits(float2 a) { return *reinterpret_cast<pk64*>(&a); }
__device__ __forceinline__ float2 pk_val(pk64 a) { return *reinterpret_cast<float2*>(&a); }
__device__ __forceinline__ float2 pk_add(float2 a, float2 b) { pk64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk_bits(a)), "l"(pk_bits(b))); return pk_val(d); }
__device__ __forceinline__ float2 pk_sub(float2 a, float2 b) { pk64 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk_bits(a)), "l"(pk_bits(b))); return pk_val(d); }
__device__ __forceinline__ float2 pk_mul(float2 a, float2 b) { pk64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk_bits(a)), "l"(pk_bits(b))); return pk_val(d); }
__device__ __forceinline__ float2 pk_fma(float2 a, float2 b, float2 c) {
    pk64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(pk_bits(a)), "l"(pk_bits(b)), "l"(pk_bits(c))); return pk_val(d);
}
__device__ __forceinline__ float2 pk_bcast(float s) { return make_float2(s, s); }
__device__ __forceinline__ float2 rot_mi(float2 v) { return make_float2(v.y, -v.x); }     // v * (-i)
__device__ __forceinline__ float2 rot_pi(float2 v) { return make_float2(-v.y, v.x); }     // v * (+i)

// Two flavours of the complex / FFT helpers.  A translation unit that defines ADN_PACKED_FP32 before including this header gets
// the packed ones (stft.cu: -27 % executed instructions, 0.427 -> 0.409 ms); the default is scalar fp32 (istft.cu: its seeded-phase
// variant is register-bound at 128 and the 64-bit operand pairs of the packed form push it into spills).
#ifdef ADN_PACKED_FP32
// ---------------------------------------------------------------- small complex helpers (float2 = re, im)
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return pk_add(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return pk_sub(a, b); }
// a * b = b.x * a + b.y * (i a): two packed instructions with b's components as scalar-broadcast operands
__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return pk_fma(rot_pi(a), pk_bcast(b.y), pk_mul(a, pk_bcast(b.x))); }
__device__ __forceinline__ float2 cmul_conj(float2 a, float2 b) {   // a * conj(b) = b.x * a - b.y * (i a)
    return pk_fma(rot_mi(a), pk_bcast(b.y), pk_mul(a, pk_bcast(b.x)));
}

// 4-point DFT, in place.  INV = false: forward (W4 = -i); INV = true: inverse (W4 = +i), unnormalised.  8 packed adds.
template <bool INV>
__device__ __forceinline__ void dft4(float2& x0, float2& x1, float2& x2, float2& x3) {
    const float2 a0 = pk_add(x0, x2), a1 = pk_sub(x0, x2), a2 = pk_add(x1, x3), a3 = pk_sub(x1, x3);
    x0 = pk_add(a0, a2);
    x2 = pk_sub(a0, a2);
    const float2 r = INV ? rot_pi(a3) : rot_mi(a3);      // forward: x1 = a1 - i*a3, x3 = a1 + i*a3
    x1 = pk_add(a1, r);
    x3 = pk_sub(a1, r);
}

// multiply by W16^M (forward, exp(-2 pi i M/16)) or its conjugate (INV)
template <int M, bool INV>
__device__ __forceinline__ float2 tw16(float2 v) {
    constexpr float C8 = 0.70710678118654752440f;   // cos(pi/4)
    constexpr float C1 = 0.92387953251128675613f;   // cos(pi/8)
    constexpr float S1 = 0.38268343236508977173f;   // sin(pi/8)
    if constexpr (M == 0) return v;
    else if constexpr (M == 4) return INV ? rot_pi(v) : rot_mi(v);
    else if constexpr (M == 2) return pk_mul(INV ? pk_add(v, rot_pi(v)) : pk_add(v, rot_mi(v)), pk_bcast(C8));      // (1 -+ i)/sqrt2
    else if constexpr (M == 6) return pk_mul(INV ? pk_sub(rot_pi(v), v) : pk_sub(rot_mi(v), v), pk_bcast(C8));      // (-1 -+ i)/sqrt2
    else {
        // general: w = wr + i wi (forward wi < 0 for M = 1, 3): v * w = wr * v + wi * (i v)
        constexpr float wr = (M == 1) ? C1 : (M == 3) ? S1 : /*M == 9*/ -C1;
        constexpr float wi_f = (M == 1) ? -S1 : (M == 3) ? -C1 : /*M == 9*/ S1;
        constexpr float wi = INV ? -wi_f : wi_f;
        return pk_fma(rot_pi(v), pk_bcast(wi), pk_mul(v, pk_bcast(wr)));
    }
}

// 16-point DFT in registers, natural order in and out (v[k] <- sum_n v[n] W16^{nk}).
template <bool INV>
__device__ __forceinline__ void dft16(float2 (&v)[16]) {
    // n = 4*n1 + n2 ; k = k1 + 4*k2.  Stage A: DFT over n1 for each n2.
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) dft4<INV>(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);
    // now v[4*k1 + n2] holds y[k1][n2]; twiddle by W16^{n2*k1}
    v[5] = tw16<1, INV>(v[5]);   v[6] = tw16<2, INV>(v[6]);   v[7] = tw16<3, INV>(v[7]);
    v[9] = tw16<2, INV>(v[9]);   v[10] = tw16<4, INV>(v[10]); v[11] = tw16<6, INV>(v[11]);
    v[13] = tw16<3, INV>(v[13]); v[14] = tw16<6, INV>(v[14]); v[15] = tw16<9, INV>(v[15]);
    // Stage B: DFT over n2 for each k1: results X[k1 + 4*k2] land in v[4*k1 + k2]
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) dft4<INV>(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
    // transpose 4x4 so that v[k1 + 4*k2] = X[k1 + 4*k2]
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = a + 1; b < 4; ++b) {
            float2 t = v[4 * a + b];
            v[4 * a + b] = v[4 * b + a];
            v[4 * b + a] = t;
        }
}

#else
// ---------------------------------------------------------------- small complex helpers (float2 = re, im)
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 cmul_conj(float2 a, float2 b) {   // a * conj(b)
    return make_float2(fmaf(a.x, b.x, a.y * b.y), fmaf(a.y, b.x, -a.x * b.y));
}

// 4-point DFT, in place.  INV = false: forward (W4 = -i); INV = true: inverse (W4 = +i), unnormalised.
template <bool INV>
__device__ __forceinline__ void dft4(float2& x0, float2& x1, float2& x2, float2& x3) {
    float2 a0 = cadd(x0, x2), a1 = csub(x0, x2), a2 = cadd(x1, x3), a3 = csub(x1, x3);
    x0 = cadd(a0, a2);
    x2 = csub(a0, a2);
    // forward: x1 = a1 - i*a3, x3 = a1 + i*a3 ; (-i)*(re,im) = (im,-re)
    float2 r = INV ? make_float2(-a3.y, a3.x) : make_float2(a3.y, -a3.x);
    x1 = cadd(a1, r);
    x3 = csub(a1, r);
}

// multiply by W16^M (forward, exp(-2 pi i M/16)) or its conjugate (INV)
template <int M, bool INV>
__device__ __forceinline__ float2 tw16(float2 v) {
    constexpr float C8 = 0.70710678118654752440f;   // cos(pi/4)
    constexpr float C1 = 0.92387953251128675613f;   // cos(pi/8)
    constexpr float S1 = 0.38268343236508977173f;   // sin(pi/8)
    if constexpr (M == 0) return v;
    else if constexpr (M == 4) return INV ? make_float2(-v.y, v.x) : make_float2(v.y, -v.x);
    else if constexpr (M == 2) return INV ? make_float2((v.x - v.y) * C8, (v.x + v.y) * C8)
                                           : make_float2((v.x + v.y) * C8, (v.y - v.x) * C8);
    else if constexpr (M == 6) return INV ? make_float2(-(v.x + v.y) * C8, (v.x - v.y) * C8)
                                           : make_float2((v.y - v.x) * C8, -(v.x + v.y) * C8);
    else {
        // general: w = (cos a, -sin a) forward with a = 2 pi M / 16
        constexpr float wr = (M == 1) ? C1 : (M == 3) ? S1 : /*M == 9*/ -C1;
        constexpr float wi_f = (M == 1) ? -S1 : (M == 3) ? -C1 : /*M == 9*/ S1;
        const float wi = INV ? -wi_f : wi_f;
        return make_float2(fmaf(v.x, wr, -v.y * wi), fmaf(v.x, wi, v.y * wr));
    }
}

// 16-point DFT in registers, natural order in and out (v[k] <- sum_n v[n] W16^{nk}).
template <bool INV>
__device__ __forceinline__ void dft16(float2 (&v)[16]) {
    // n = 4*n1 + n2 ; k = k1 + 4*k2.  Stage A: DFT over n1 for each n2.
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) dft4<INV>(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);
    // now v[4*k1 + n2] holds y[k1][n2]; twiddle by W16^{n2*k1}
    v[5] = tw16<1, INV>(v[5]);   v[6] = tw16<2, INV>(v[6]);   v[7] = tw16<3, INV>(v[7]);
    v[9] = tw16<2, INV>(v[9]);   v[10] = tw16<4, INV>(v[10]); v[11] = tw16<6, INV>(v[11]);
    v[13] = tw16<3, INV>(v[13]); v[14] = tw16<6, INV>(v[14]); v[15] = tw16<9, INV>(v[15]);
    // Stage B: DFT over n2 for each k1: results X[k1 + 4*k2] land in v[4*k1 + k2]
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) dft4<INV>(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
    // transpose 4x4 so that v[k1 + 4*k2] = X[k1 + 4*k2]
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = a + 1; b < 4; ++b) {
            float2 t = v[4 * a + b];
            v[4 * a + b] = v[4 * b + a];
            v[4 * b + a] = t;
        }
}

#endif

}  // namespace adn
