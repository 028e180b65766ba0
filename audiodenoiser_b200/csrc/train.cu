// train.cu -- the HBM-bound kernels of the training step (reference code/train.py:65-72 around code/model.py in train mode):
// train-mode BatchNorm2d (batch statistics, running-stat update, model.py:12,15), its backward fused with the ReLU backward,
// MaxPool2d backward fused with the skip-connection gradient add (model.py:31,49), the 1x1 head (model.py:68,93) forward and
// backward, the first-layer (Cin = 1) weight gradient, per-channel sums (ConvTranspose2d bias gradient), the global gradient
// norm of clip_grad_norm_(max_norm) (train.py:70) and the AdamW update (train.py:71,124).
//
// Activations and activation gradients are NHWC bf16 with C a multiple of 8 (64 in this network); every reduction is a
// deterministic two-stage tree (per-block partials in a workspace, then a fixed-order finalise) -- no atomics.
#include "adn_common.cuh"

namespace adn {

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(p[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint4 u;
    __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return u;
}
__device__ __forceinline__ void load8f(const float* p, float (&f)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

constexpr int TR_THREADS = 256;
constexpr int TR_MAX_BLOCKS = 592;            // 148 SMs x 4: upper bound of every partial-sum grid

// ------------------------------------------------------------------------------------------------ per-channel reductions
// Thread (g = tid % G, pl = tid / G) owns the 8 channels of group g (G = C / 8) and walks pixels pl, pl + PL, ... of the block's
// pixel range; the block then folds its PL pixel lanes through shared memory.  partial[block][slot][C], NS slots per mode.
enum { RED_STATS = 0, RED_BNBWD = 1, RED_SUM = 2, RED_HEAD = 3, RED_C1W = 4 };

struct RedArgs {
    const uint4* a;            // STATS: z ; BNBWD / SUM: dy ; HEAD: y ; C1W: dz
    long long a_ld8;           // pixel stride of `a` in uint4 units (>= C/8: a channel slice of a wider tensor)
    const uint4* z;            // BNBWD: pre-BN activations (dense)
    const float* scale; const float* shift; const float* mean; const float* invstd;   // BNBWD
    const float* f32;          // HEAD: d_out per pixel ; C1W: the (n,1,h,w) input
    int h, w;                  // C1W
    long long pixels;
    int c;
    float* partial;
};

template <int MODE> struct RedSlots { static constexpr int value = (MODE == RED_SUM) ? 1 : (MODE == RED_C1W) ? 9 : 2; };

template <int MODE>
__global__ void __launch_bounds__(TR_THREADS)
channel_reduce_kernel(const RedArgs a) {
    constexpr int NS = RedSlots<MODE>::value;
    extern __shared__ float s_red[];                       // [NS][TR_THREADS][8]
    const int G = a.c >> 3, PL = TR_THREADS / G;
    const int g = threadIdx.x % G, pl = threadIdx.x / G;
    const long long per_block = (a.pixels + gridDim.x - 1) / gridDim.x;
    const long long p0 = blockIdx.x * per_block, p1 = min(a.pixels, p0 + per_block);
    float acc[NS][8];
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[s][i] = 0.f;
    float sc[8], sh[8], mu[8], is[8];
    if (MODE == RED_BNBWD && pl < PL) {
        load8f(a.scale + g * 8, sc); load8f(a.shift + g * 8, sh); load8f(a.mean + g * 8, mu); load8f(a.invstd + g * 8, is);
    }
    if (pl < PL) {
        for (long long p = p0 + pl; p < p1; p += PL) {
            float x[8];
            unpack8(__ldg(a.a + p * a.a_ld8 + g), x);
            if (MODE == RED_STATS) {
#pragma unroll
                for (int i = 0; i < 8; ++i) { acc[0][i] += x[i]; acc[1][i] = fmaf(x[i], x[i], acc[1][i]); }
            } else if (MODE == RED_SUM) {
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[0][i] += x[i];
            } else if (MODE == RED_BNBWD) {
                float z[8];
                unpack8(__ldg(a.z + p * G + g), z);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float gate = fmaf(z[i], sc[i], sh[i]) > 0.f ? x[i] : 0.f;      // dy through the ReLU
                    acc[0][i] += gate;
                    acc[1][i] = fmaf(gate, (z[i] - mu[i]) * is[i], acc[1][i]);
                }
            } else if (MODE == RED_HEAD) {
                const float d = __ldg(a.f32 + p);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[0][i] = fmaf(d, x[i], acc[0][i]);
                if (g == 0) acc[1][0] += d;                                              // slot 1, channel 0: sum of d_out (bias)
            } else {                                                                     // RED_C1W: dW[c][tap] += dz[p][c] * x[p + tap]
                const int px = (int)(p % a.w), py = (int)((p / a.w) % a.h);
                const float* xr = a.f32 + p;
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const int yy = py + ky - 1, xx = px + kx - 1;
                        const float v = (yy >= 0 && yy < a.h && xx >= 0 && xx < a.w) ? __ldg(xr + (ky - 1) * a.w + (kx - 1)) : 0.f;
#pragma unroll
                        for (int i = 0; i < 8; ++i) acc[ky * 3 + kx][i] = fmaf(x[i], v, acc[ky * 3 + kx][i]);
                    }
            }
        }
    }
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int i = 0; i < 8; ++i) s_red[(s * TR_THREADS + threadIdx.x) * 8 + i] = acc[s][i];
    __syncthreads();
    // thread t < NS * C folds pixel lanes in ascending order for (slot, channel) = (t / C, t % C)
    for (int t = threadIdx.x; t < NS * a.c; t += TR_THREADS) {
        const int s = t / a.c, ch = t - s * a.c;
        float sum = 0.f;
        for (int l = 0; l < PL; ++l) sum += s_red[(s * TR_THREADS + l * G + (ch >> 3)) * 8 + (ch & 7)];
        a.partial[((long long)blockIdx.x * NS + s) * a.c + ch] = sum;
    }
}

// fixed-order fold of the per-block partials (double accumulation) by ONE WARP per (slot, channel): lane l adds blocks l, l + 32,
// ... in ascending order, then a butterfly combines the 32 lane sums -- the same order every run, and every lane gets the result.
// Must be called by all 32 lanes of a warp with the same (slot, ch).
__device__ __forceinline__ double fold_partials(const float* partial, int nb, int ns, int c, int slot, int ch) {
    const int lane = threadIdx.x & 31;
    double s = 0.0;
    // eight loads in flight per lane (a one-at-a-time loop made each finalise launch cost 5-7 us of pure memory latency); the
    // additions keep the ascending block order, so the result is bit-identical to the sequential loop
    for (int b0 = lane; b0 < nb; b0 += 32 * 8) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int b = b0 + 32 * j;
            v[j] = b < nb ? partial[((long long)b * ns + slot) * c + ch] : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) s += (double)v[j];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;
}

// train-mode BatchNorm2d statistics (model.py:12,15; torch defaults eps = 1e-5, momentum = 0.1): biased variance for the
// normalisation, unbiased for the running estimate.
__global__ void bn_finalize_fwd_kernel(const float* __restrict__ partial, int nb, int c, double count, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, float eps, float momentum, float* __restrict__ running_mean,
                                       float* __restrict__ running_var, float* __restrict__ scale, float* __restrict__ shift,
                                       float* __restrict__ mean_out, float* __restrict__ invstd_out) {
    const int ch = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;       // one warp per channel
    if (ch >= c) return;
    const double s1 = fold_partials(partial, nb, 2, c, 0, ch), s2 = fold_partials(partial, nb, 2, c, 1, ch);
    if (threadIdx.x & 31) return;
    const double mean = s1 / count;
    double var = s2 / count - mean * mean;
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = gamma[ch] * invstd;
    scale[ch] = sc;
    shift[ch] = fmaf(-(float)mean, sc, beta[ch]);
    mean_out[ch] = (float)mean;
    invstd_out[ch] = invstd;
    if (running_mean) {
        const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
        running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * (float)mean;
        running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * (float)unbiased;
    }
}

// out[slot][ch] = fold of slot (d_gamma = S2 / d_beta = S1 of the BN backward, bias gradients, head gradients)
__global__ void fold_kernel(const float* __restrict__ partial, int nb, int ns, int c, int slot, float* __restrict__ out) {
    const int ch = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;       // one warp per channel
    if (ch >= c) return;
    const double v = fold_partials(partial, nb, ns, c, slot, ch);
    if ((threadIdx.x & 31) == 0) out[ch] = (float)v;
}
// both slots of a 2-slot reduction in one launch (BN backward: slot 0 -> d_beta, slot 1 -> d_gamma): one warp per (slot, channel)
__global__ void fold2_kernel(const float* __restrict__ partial, int nb, int c, float* __restrict__ out0, float* __restrict__ out1) {
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (t >= 2 * c) return;
    const int slot = t / c, ch = t - slot * c;
    const double v = fold_partials(partial, nb, 2, c, slot, ch);
    if ((threadIdx.x & 31) == 0) (slot ? out1 : out0)[ch] = (float)v;
}
// first-layer weight gradient in the reference layout (64, 1, 3, 3): out[ch * 9 + tap]
__global__ void fold_c1w_kernel(const float* __restrict__ partial, int nb, float* __restrict__ out) {
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;        // one warp per (tap, channel)
    if (t >= 576) return;
    const double v = fold_partials(partial, nb, 9, 64, t >> 6, t & 63);
    if ((threadIdx.x & 31) == 0) out[(t & 63) * 9 + (t >> 6)] = (float)v;
}

// ------------------------------------------------------------------------------------------------ elementwise passes
// y = max(z * scale + shift, 0): the normalise + affine + ReLU of train-mode BN (model.py:12-13)
__global__ void __launch_bounds__(TR_THREADS)
bn_relu_apply_kernel(const uint4* __restrict__ z, const float* __restrict__ scale, const float* __restrict__ shift, long long pixels,
                     int c8, uint4* __restrict__ y) {
    const long long total = pixels * c8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(i % c8);
        float x[8], sc[8], sh[8];
        unpack8(__ldg(z + i), x);
        load8f(scale + g * 8, sc); load8f(shift + g * 8, sh);
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = fmaxf(fmaf(x[k], sc[k], sh[k]), 0.f);
        y[i] = pack8(x);
    }
}

// BatchNorm2d + ReLU backward: dz = gamma * invstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * [y > 0]
__global__ void __launch_bounds__(TR_THREADS)
bn_relu_bwd_apply_kernel(const uint4* __restrict__ dy, long long dy_ld8, const uint4* __restrict__ z, const float* __restrict__ scale,
                         const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ invstd,
                         const float* __restrict__ s1, const float* __restrict__ s2, float inv_count, long long pixels, int c8,
                         uint4* __restrict__ dz) {
    const long long total = pixels * c8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(i % c8);
        const long long p = i / c8;
        float d[8], x[8], sc[8], sh[8], mu[8], is[8], a1[8], a2[8];
        unpack8(__ldg(dy + p * dy_ld8 + g), d);
        unpack8(__ldg(z + i), x);
        load8f(scale + g * 8, sc); load8f(shift + g * 8, sh); load8f(mean + g * 8, mu); load8f(invstd + g * 8, is);
        load8f(s1 + g * 8, a1); load8f(s2 + g * 8, a2);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float gate = fmaf(x[k], sc[k], sh[k]) > 0.f ? d[k] : 0.f;
            const float xh = (x[k] - mu[k]) * is[k];
            d[k] = sc[k] * (gate - a1[k] * inv_count - xh * a2[k] * inv_count);
        }
        dz[i] = pack8(d);
    }
}

// MaxPool2d(2) backward + skip add (model.py:31,49): d_skip = d_from_decoder + route(d_pool -> first maximum of each 2x2 window).
// One thread per (pooled pixel, 8-channel group); h, w even.  d_dec may be a channel slice (pixel stride dec_ld8) or NULL.
__global__ void __launch_bounds__(TR_THREADS)
maxpool_bwd_add_kernel(const uint4* __restrict__ y, const uint4* __restrict__ d_pool, const uint4* __restrict__ d_dec, long long dec_ld8,
                       int n, int h, int w, int c8, uint4* __restrict__ out) {
    const int ho = h >> 1, wo = w >> 1;
    const long long total = (long long)n * ho * wo * c8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(i % c8);
        const int xo = (int)((i / c8) % wo);
        const int yo = (int)((i / ((long long)c8 * wo)) % ho);
        const long long img = i / ((long long)c8 * wo * ho);
        const long long q0 = (img * h + 2 * yo) * w + 2 * xo;            // pixel index of the window's top-left corner
        const long long q[4] = {q0, q0 + 1, q0 + w, q0 + w + 1};
        float v[4][8], dp[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) unpack8(__ldg(y + q[j] * c8 + g), v[j]);
        unpack8(__ldg(d_pool + i), dp);
        int arg[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            int best = 0; float bv = v[0][k];
#pragma unroll
            for (int j = 1; j < 4; ++j) if (v[j][k] > bv) { bv = v[j][k]; best = j; }
            arg[k] = best;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float o[8];
            if (d_dec) unpack8(__ldg(d_dec + q[j] * dec_ld8 + g), o);
            else {
#pragma unroll
                for (int k = 0; k < 8; ++k) o[k] = 0.f;
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] += (arg[k] == j) ? dp[k] : 0.f;
            out[q[j] * c8 + g] = pack8(o);
        }
    }
}

// 1x1 head (model.py:68,93): out[p] = b + sum_c y[p][c] w[c], 64 channels, 8 lanes per pixel
__global__ void __launch_bounds__(TR_THREADS)
head_fwd_kernel(const uint4* __restrict__ y, const float* __restrict__ w, const float* __restrict__ b, long long pixels, float* __restrict__ out) {
    const int g = threadIdx.x & 7;
    float wv[8];
    load8f(w + g * 8, wv);
    const float bias = b[0];
    const long long p_stride = (long long)gridDim.x * (TR_THREADS / 8);
    const long long iters = (pixels + p_stride - 1) / p_stride;
    for (long long it = 0; it < iters; ++it) {                            // uniform trip count: the shuffles stay converged
        const long long p = it * p_stride + blockIdx.x * (TR_THREADS / 8) + (threadIdx.x >> 3);
        float s = 0.f;
        if (p < pixels) {
            float x[8];
            unpack8(__ldg(y + p * 8 + g), x);
#pragma unroll
            for (int k = 0; k < 8; ++k) s = fmaf(x[k], wv[k], s);
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        if (g == 0 && p < pixels) out[p] = s + bias;
    }
}
// d_y[p][c] = d_out[p] * w[c]
__global__ void __launch_bounds__(TR_THREADS)
head_bwd_kernel(const float* __restrict__ d_out, const float* __restrict__ w, long long pixels, uint4* __restrict__ dy) {
    const long long total = pixels * 8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(i & 7);
        const float d = __ldg(d_out + (i >> 3));
        float wv[8];
        load8f(w + g * 8, wv);
#pragma unroll
        for (int k = 0; k < 8; ++k) wv[k] *= d;
        dy[i] = pack8(wv);
    }
}

// ------------------------------------------------------------------------------------------------ clip_grad_norm_ + AdamW
__global__ void __launch_bounds__(TR_THREADS)
sumsq_kernel(const float* __restrict__ g, long long n, double* __restrict__ partial) {
    double s = 0.0;
    const long long n4 = (reinterpret_cast<uintptr_t>(g) & 15) == 0 ? n >> 2 : 0;      // 16-byte loads (fixed order: still deterministic)
    const float4* g4 = reinterpret_cast<const float4*>(g);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = g4[i];
        s += (double)v.x * (double)v.x + (double)v.y * (double)v.y + ((double)v.z * (double)v.z + (double)v.w * (double)v.w);
    }
    for (long long i = 4 * n4 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = g[i];
        s += (double)v * (double)v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __shared__ double sh[TR_THREADS / 32];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < TR_THREADS / 32; ++k) t += sh[k];
        partial[blockIdx.x] = t;
    }
}
// norm_out[0] = total L2 norm, norm_out[1] = clip coefficient min(1, max_norm / (norm + 1e-6))  (torch.nn.utils.clip_grad_norm_)
__global__ void grad_norm_finalize_kernel(const double* __restrict__ partial, int nb, float max_norm, float* __restrict__ norm_out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double t = 0.0;
        for (int b = 0; b < nb; ++b) t += partial[b];
        const float norm = (float)sqrt(t);
        const float coef = max_norm / (norm + 1e-6f);
        norm_out[0] = norm;
        norm_out[1] = coef < 1.f ? coef : 1.f;
    }
}
// torch.optim.AdamW (decoupled weight decay, bias-corrected): g is scaled by the clip coefficient read from the device
__global__ void __launch_bounds__(TR_THREADS)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
             const float* __restrict__ clip, float lr, float beta1, float beta2, float eps, float weight_decay, float bc1, float bc2) {
    const float coef = clip ? clip[1] : 1.f;
    const float step = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float gi = g[i] * coef;
        float pi = p[i] * (1.f - lr * weight_decay);
        const float mi = fmaf(beta1, m[i], (1.f - beta1) * gi);
        const float vi = fmaf(beta2, v[i], (1.f - beta2) * gi * gi);
        m[i] = mi; v[i] = vi;
        pi -= step * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
        p[i] = pi;
    }
}

// graph-capturable variant: the 1-based step count lives on the device (step_counter[0], advanced by the launch itself), so a
// captured training step can be replayed without baking the bias corrections into the graph
__global__ void adamw_advance_step_kernel(float* __restrict__ step_counter) {
    if (threadIdx.x == 0 && blockIdx.x == 0) step_counter[0] += 1.f;
}
__global__ void __launch_bounds__(TR_THREADS)
adamw_dev_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
                      const float* __restrict__ clip, const float* __restrict__ step_counter, float lr, float beta1, float beta2, float eps,
                      float weight_decay) {
    const float t = step_counter[0];
    const float bc1 = 1.f - powf(beta1, t), bc2 = 1.f - powf(beta2, t);
    const float coef = clip ? clip[1] : 1.f;
    const float step = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
    auto upd = [&](float& pi, float gi, float& mi, float& vi) {        // one element, the arithmetic of adamw_kernel
        gi *= coef;
        pi *= (1.f - lr * weight_decay);
        mi = fmaf(beta1, mi, (1.f - beta1) * gi);
        vi = fmaf(beta2, vi, (1.f - beta2) * gi * gi);
        pi -= step * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
    };
    // 16-byte accesses: four streams of 4-byte loads from 150 k threads keep ~2.4 MB in flight, half of what HBM needs (the flat
    // buffers are 256-byte aligned)
    const long long n4 = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                           reinterpret_cast<uintptr_t>(v)) & 15) == 0 ? n >> 2 : 0;
    float4* p4 = reinterpret_cast<float4*>(p); const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* m4 = reinterpret_cast<float4*>(m); float4* v4 = reinterpret_cast<float4*>(v);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 pp = p4[i], gg = g4[i], mm = m4[i], vv = v4[i];
        upd(pp.x, gg.x, mm.x, vv.x); upd(pp.y, gg.y, mm.y, vv.y); upd(pp.z, gg.z, mm.z, vv.z); upd(pp.w, gg.w, mm.w, vv.w);
        m4[i] = mm; v4[i] = vv; p4[i] = pp;
    }
    for (long long i = 4 * n4 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float pi = p[i], mi = m[i], vi = v[i];
        upd(pi, g[i], mi, vi);
        m[i] = mi; v[i] = vi; p[i] = pi;
    }
}

static inline int tr_grid(long long work_items) {
    long long g = (work_items + TR_THREADS - 1) / TR_THREADS;
    long long cap = (long long)num_sms() * 4;
    if (cap > TR_MAX_BLOCKS) cap = TR_MAX_BLOCKS;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}
static inline int red_blocks(long long pixels, int c) {
    const int pl = TR_THREADS / (c / 8);
    long long g = (pixels + (long long)pl * 16 - 1) / ((long long)pl * 16);      // >= 16 pixels per lane
    long long cap = (long long)num_sms() * 4;
    if (cap > TR_MAX_BLOCKS) cap = TR_MAX_BLOCKS;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}
template <int MODE>
static int launch_reduce(RedArgs& a, int* nb_out, cudaStream_t stream) {
    if (a.c <= 0 || (a.c & 7) || a.c / 8 > TR_THREADS || a.pixels <= 0 || !a.a || !a.partial) return ADN_ERR_ARG;
    int st = check_device(); if (st != ADN_OK) return st;
    const int nb = red_blocks(a.pixels, a.c);
    const size_t smem = (size_t)RedSlots<MODE>::value * TR_THREADS * 8 * sizeof(float);
    static unsigned char smem_set[64] = {0};
    ADN_CUDA_TRY(ensure_dyn_smem(channel_reduce_kernel<MODE>, (int)smem, smem_set));
    channel_reduce_kernel<MODE><<<nb, TR_THREADS, smem, stream>>>(a);
    ADN_LAUNCH_CHECK();
    *nb_out = nb;
    return ADN_OK;
}

}  // namespace adn

using namespace adn;

extern "C" int64_t adn_train_workspace_bytes(void) { return (int64_t)TR_MAX_BLOCKS * 2 * 2048 * sizeof(float); }

extern "C" int adn_bn_train_stats_f32(const void* z, int64_t pixels, int c, const float* gamma, const float* beta, float eps,
                                      float momentum, float* running_mean, float* running_var, float* scale, float* shift,
                                      float* mean, float* invstd, void* workspace, void* stream) {
    if (!gamma || !beta || !scale || !shift || !mean || !invstd || !workspace) return ADN_ERR_ARG;
    RedArgs a{}; a.a = (const uint4*)z; a.a_ld8 = c / 8; a.pixels = pixels; a.c = c; a.partial = (float*)workspace;
    int nb = 0;
    int st = launch_reduce<RED_STATS>(a, &nb, (cudaStream_t)stream);
    if (st != ADN_OK) return st;
    bn_finalize_fwd_kernel<<<(c * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const float*)workspace, nb, c, (double)pixels, gamma, beta, eps,
                                                                              momentum, running_mean, running_var, scale, shift, mean, invstd);
    ADN_LAUNCH_CHECK();
    return ADN_OK;
}

extern "C" int adn_bn_relu_apply_bf16(const void* z, const float* scale, const float* shift, int64_t pixels, int c, void* y, void* stream) {
    if (!z || !scale || !shift || !y || pixels <= 0 || c <= 0 || (c & 7)) return ADN_ERR_ARG;
    int st = check_device(); if (st != ADN_OK) return st;
    bn_relu_apply_kernel<<<tr_grid(pixels * (c / 8)), TR_THREADS, 0, (cudaStream_t)stream>>>((const uint4*)z, scale, shift, pixels, c / 8, (uint4*)y);
    ADN_LAUNCH_CHECK();
    return ADN_OK;
}

extern "C" int adn_bn_relu_backward_bf16(const void* dy, int dy_ld, const void* z, int64_t pixels, int c, const float* scale,
                                         const float* shift, const float* mean, const float* invstd, float* d_gamma, float* d_beta,
                                         void* dz, void* workspace, void* stream) {
    if (!dy || !z || !scale || !shift || !mean || !invstd || !d_gamma || !d_beta || !dz || !workspace || dy_ld < c || (dy_ld & 7)) return ADN_ERR_ARG;
    RedArgs a{}; a.a = (const uint4*)dy; a.a_ld8 = dy_ld / 8; a.z = (const uint4*)z; a.scale = scale; a.shift = shift; a.mean = mean;
    a.invstd = invstd; a.pixels = pixels; a.c = c; a.partial = (float*)workspace;
    int nb = 0;
    int st = launch_reduce<RED_BNBWD>(a, &nb, (cudaStream_t)stream);
    if (st != ADN_OK) return st;
    cudaStream_t s = (cudaStream_t)stream;
    fold2_kernel<<<(2 * c * 32 + 255) / 256, 256, 0, s>>>((const float*)workspace, nb, c, d_beta, d_gamma);
    bn_relu_bwd_apply_kernel<<<tr_grid(pixels * (c / 8)), TR_THREADS, 0, s>>>((const uint4*)dy, dy_ld / 8, (const uint4*)z, scale, shift, mean, invstd,
                                                                             d_beta, d_gamma, (float)(1.0 / (double)pixels), pixels, c / 8, (uint4*)dz);
    ADN_LAUNCH_CHECK();
    return ADN_OK;
}

extern "C" int adn_channel_sum_f32(const void* x, int x_ld, int64_t pixels, int c, float* out, void* workspace, void* stream) {
    if (!out || !workspace || x_ld < c || (x_ld & 7)) return ADN_ERR_ARG;
    RedArgs a{}; a.a = (const uint4*)x; a.a_ld8 = x_ld / 8; a.pixels = pixels; a.c = c; a.partial = (float*)workspace;
    int nb = 0;
    int st = launch_reduce<RED_SUM>(a, &nb, (cudaStream_t)stream);
    if (st != ADN_OK) return st;
    fold_kernel<<<(c * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const float*)workspace, nb, 1, c, 0, out);
    ADN_LAUNCH_CHECK();
    return ADN_OK;
}

extern "C" int adn_maxpool2x2_backward_add_bf16(const void* y, const void* d_pool, const void* d_dec, int dec_ld, int n, int h, int w,
                                                int c, void* out, void* stream) {
    if (!y || !d_pool || !out || n <= 0 || h < 2 || w < 2 || (h & 1) || (w & 1) || c <= 0 || (c & 7)) return ADN_ERR_ARG;
    if (d_dec && (dec_ld < c || (dec_ld & 7))) return ADN_ERR_ARG;
    int st = check_device(); if (st != ADN_OK) return st;
    maxpool_bwd_add_kernel<<<tr_grid((long long)n * (h / 2) * (w / 2) * (c / 8)), TR_THREADS, 0, (cudaStream_t)stream>>>(
        (const uint4*)y, (const uint4*)d_pool, (const uint4*)d_dec, dec_ld / 8, n, h, w, c / 8, (uint4*)out);
    ADN_LAUNCH_CHECK();
    return ADN_OK;
}

extern "C" int adn_head1x1_forward_f32(const void* y, const float* w, const float* b, int64_t pixels, float* out, void* stream) {
    if (!y || !w || !b || !out || pixels <= 0) return ADN_ERR_ARG;
    int st = check_device(); if (st != ADN_OK) return st;
    head_fwd_kernel<<<tr_grid(pixels * 8), TR_THREADS, 0, (cudaStream_t)stream>>>((const uint4*)y, w, b, pixels, out);
    ADN_LAUNCH_CHECK();
    return ADN_OK;
}

extern "C" int adn_head1x1_backward(const void* y, const float* d_out, const float* w, int64_t pixels, void* dy, float* d_w, float* d_b,
                                    void* workspace, void* stream) {
    if (!y || !d_out || !w || !dy || !d_w || !d_b || !workspace || pixels <= 0) return ADN_ERR_ARG;
    RedArgs a{}; a.a = (const uint4*)y; a.a_ld8 = 8; a.f32 = d_out; a.pixels = pixels; a.c = 64; a.partial = (float*)workspace;
    int nb = 0;
    int st = launch_reduce<RED_HEAD>(a, &nb, (cudaStream_t)stream);
    if (st != ADN_OK) return st;
    cudaStream_t s = (cudaStream_t)stream;
    fold_kernel<<<8, 256, 0, s>>>((const float*)workspace, nb, 2, 64, 0, d_w);
    fold_kernel<<<8, 256, 0, s>>>((const float*)workspace, nb, 2, 64, 1, d_b);      // d_b: 64 floats, only element 0 is meaningful
    head_bwd_kernel<<<tr_grid(pixels * 8), TR_THREADS, 0, s>>>(d_out, w, pixels, (uint4*)dy);
    ADN_LAUNCH_CHECK();
    return ADN_OK;
}

extern "C" int adn_conv3x3_c1_wgrad_f32(const void* dz, const float* x, int n, int h, int w, float* d_weight, void* workspace, void* stream) {
    if (!dz || !x || !d_weight || !workspace || n <= 0 || h <= 0 || w <= 0) return ADN_ERR_ARG;
    RedArgs a{}; a.a = (const uint4*)dz; a.a_ld8 = 8; a.f32 = x; a.h = h; a.w = w; a.pixels = (long long)n * h * w; a.c = 64;
    a.partial = (float*)workspace;
    int nb = 0;
    int st = launch_reduce<RED_C1W>(a, &nb, (cudaStream_t)stream);
    if (st != ADN_OK) return st;
    fold_c1w_kernel<<<72, 256, 0, (cudaStream_t)stream>>>((const float*)workspace, nb, d_weight);
    ADN_LAUNCH_CHECK();
    return ADN_OK;
}

extern "C" int adn_grad_norm_f32(const float* grads, int64_t count, float max_norm, float* norm_and_coef, void* workspace, void* stream) {
    if (!grads || !norm_and_coef || !workspace || count <= 0) return ADN_ERR_ARG;
    int st = check_device(); if (st != ADN_OK) return st;
    const int nb = tr_grid(count);
    sumsq_kernel<<<nb, TR_THREADS, 0, (cudaStream_t)stream>>>(grads, count, (double*)workspace);
    grad_norm_finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((const double*)workspace, nb, max_norm, norm_and_coef);
    ADN_LAUNCH_CHECK();
    return ADN_OK;
}

extern "C" int adn_adamw_step_f32(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t count,
                                  const float* norm_and_coef, float lr, float beta1, float beta2, float eps, float weight_decay,
                                  int64_t step, void* stream) {
    if (!params || !grads || !exp_avg || !exp_avg_sq || count <= 0 || step < 1) return ADN_ERR_ARG;
    int st = check_device(); if (st != ADN_OK) return st;
    const float bc1 = (float)(1.0 - pow((double)beta1, (double)step)), bc2 = (float)(1.0 - pow((double)beta2, (double)step));
    adamw_kernel<<<tr_grid(count), TR_THREADS, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, count, norm_and_coef, lr, beta1,
                                                                         beta2, eps, weight_decay, bc1, bc2);
    ADN_LAUNCH_CHECK();
    return ADN_OK;
}

extern "C" int adn_adamw_step_dev_f32(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t count,
                                      const float* norm_and_coef, float* step_counter, float lr, float beta1, float beta2, float eps,
                                      float weight_decay, void* stream) {
    if (!params || !grads || !exp_avg || !exp_avg_sq || !step_counter || count <= 0) return ADN_ERR_ARG;
    int st = check_device(); if (st != ADN_OK) return st;
    adamw_advance_step_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(step_counter);
    adamw_dev_step_kernel<<<tr_grid(count), TR_THREADS, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, count, norm_and_coef,
                                                                                  step_counter, lr, beta1, beta2, eps, weight_decay);
    ADN_LAUNCH_CHECK();
    return ADN_OK;
}
