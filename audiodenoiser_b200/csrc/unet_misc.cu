// unet_misc.cu -- the HBM-bound pieces around the tcgen05 convolutions of the reference UNet (code/model.py):
// checkpoint packing, BN folding, the Cin=1 first layer (direct conv), 2x2 max-pool, layout converters, the
// SpectrogramDataset transform (code/data_loader.py:41-72) and error statistics.
#include <cstdlib>
#include "adn_common.cuh"

namespace adn {

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

// ------------------------------------------------------------------------------------------------ packing
__global__ void pack_conv3x3_kernel(const float* __restrict__ w, int co_n, int ci_n, __nv_bfloat16* __restrict__ out) {
    const long long total = (long long)co_n * 9 * ci_n;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ci = (int)(i % ci_n);
        const int tap = (int)((i / ci_n) % 9);
        const int co = (int)(i / ((long long)ci_n * 9));
        out[i] = __float2bfloat16_rn(w[((long long)co * ci_n + ci) * 9 + tap]);
    }
}

__global__ void pack_convt2x2_kernel(const float* __restrict__ w, int ci_n, int co_n, __nv_bfloat16* __restrict__ out) {
    const long long total = 4ll * co_n * ci_n;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ci = (int)(i % ci_n);
        const int co = (int)((i / ci_n) % co_n);
        const int q = (int)(i / ((long long)ci_n * co_n));
        out[i] = __float2bfloat16_rn(w[((long long)ci * co_n + co) * 4 + q]);
    }
}

// data-gradient weights of Conv2d 3x3: Wd[ci][tap'][co] = W[co][ci][2-ky][2-kx] (flipped taps, transposed channels), so that
// d_in = conv3x3(d_out, Wd) runs on the forward kernel
__global__ void pack_conv3x3_dgrad_kernel(const float* __restrict__ w, int co_n, int ci_n, __nv_bfloat16* __restrict__ out) {
    const long long total = (long long)ci_n * 9 * co_n;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int co = (int)(i % co_n);
        const int tap = (int)((i / co_n) % 9);
        const int ci = (int)(i / ((long long)co_n * 9));
        out[i] = __float2bfloat16_rn(w[((long long)co * ci_n + ci) * 9 + (8 - tap)]);
    }
}
// data-gradient weights of ConvTranspose2d 2x2: Wd[ci][q * Co + co] = W[ci][co][q]
__global__ void pack_convt2x2_dgrad_kernel(const float* __restrict__ w, int ci_n, int co_n, __nv_bfloat16* __restrict__ out) {
    const long long total = 4ll * co_n * ci_n;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int co = (int)(i % co_n);
        const int q = (int)((i / co_n) % 4);
        const int ci = (int)(i / (4ll * co_n));
        out[i] = __float2bfloat16_rn(w[((long long)ci * co_n + co) * 4 + q]);
    }
}

// One launch for every layer: forward and data-gradient packs of all conv / convT weights after an optimizer step.
// blockIdx.y = table entry, blockIdx.x strides over that tensor's elements (destination-ordered: coalesced bf16 writes).
struct PackEntry {
    const float* w;
    __nv_bfloat16* fwd;
    __nv_bfloat16* dgrad;
    int c_out, c_in, kind, pad;          // kind 0: Conv2d 3x3 (Co,Ci,3,3); 1: ConvTranspose2d 2x2 (Ci,Co,2,2)
};
__global__ void __launch_bounds__(256)
pack_table_kernel(const PackEntry* __restrict__ table, int which, int n_flat) {      // which: bit 0 = forward packs, bit 1 = data-gradient packs
    __shared__ float tile[32][32 * 9 + 1];                        // [co][ci * 9 + tap], odd pitch: conflict-free both ways
    // n_flat > 0: ONE-dimensional grid, entry i owns the `pad` blocks after those of the entries before it (the host sized the grid
    // to the sum).  The 2-D form launched 1 024 blocks per entry, of which the small layers use a handful: 21 504 block launches
    // for 3 400 tiles cost more than the copies themselves.
    int ent = blockIdx.y, bx = blockIdx.x, gx = gridDim.x;
    if (n_flat > 0) {
        int b = blockIdx.x;
        ent = 0;
        while (ent < n_flat - 1 && b >= table[ent].pad) { b -= table[ent].pad; ++ent; }
        bx = b; gx = table[ent].pad;
    }
    const PackEntry e = table[ent];
    const int co_n = e.c_out, ci_n = e.c_in;
    if (e.kind == 0) {
        // 32 x 32 x 9 tiles through shared memory: the fp32 weights are read in 1 152-byte runs and both bf16 packs leave as
        // 64-byte runs (the destination-ordered version gathered 4-byte elements 36 B / Ci*36 B apart: 12-28 % sector use)
        const int ci_tiles = ci_n >> 5, tiles = (co_n >> 5) * ci_tiles;
        const int l = threadIdx.x & 31, grp = threadIdx.x >> 5;
        for (int t = bx; t < tiles; t += gx) {
            const int co0 = (t / ci_tiles) << 5, ci0 = (t % ci_tiles) << 5;
            // 32 rows x 288 floats = 36 loads per thread, issued nine at a time before they are used (the row-by-row loop kept one or
            // two loads in flight per thread and the kernel ran at a fifth of its memory floor)
            const float* src0 = e.w + ((long long)co0 * ci_n + ci0) * 9;
            const long long row_stride = (long long)ci_n * 9;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                float v[9];
#pragma unroll
                for (int j = 0; j < 9; ++j) {
                    const int idx = (b * 9 + j) * 256 + threadIdx.x;          // < 9 216
                    const int r = idx / 288, i = idx - r * 288;
                    v[j] = src0[r * row_stride + i];
                }
#pragma unroll
                for (int j = 0; j < 9; ++j) {
                    const int idx = (b * 9 + j) * 256 + threadIdx.x;
                    const int r = idx / 288, i = idx - r * 288;
                    tile[r][i] = v[j];
                }
            }
            __syncthreads();
            for (int it = grp; it < 288 && (which & 1); it += 8) {               // it = co_l * 9 + tap: 32 consecutive ci
                const int co_l = it / 9, tap = it - co_l * 9;
                e.fwd[((long long)(co0 + co_l) * 9 + tap) * ci_n + ci0 + l] = __float2bfloat16_rn(tile[co_l][l * 9 + tap]);
            }
            for (int it = grp; it < 288 && (which & 2); it += 8) {               // it = ci_l * 9 + tapd: 32 consecutive co, flipped tap
                const int ci_l = it / 9, tapd = it - ci_l * 9;
                e.dgrad[((long long)(ci0 + ci_l) * 9 + tapd) * co_n + co0 + l] = __float2bfloat16_rn(tile[l][ci_l * 9 + (8 - tapd)]);
            }
            __syncthreads();
        }
    } else {
        const long long stride = (long long)gx * blockDim.x;
        const long long i0 = bx * (long long)blockDim.x + threadIdx.x;
        const long long total = 4ll * co_n * ci_n;
        for (long long i = i0; i < total && (which & 1); i += stride) {          // fwd [q][co][ci]
            const int ci = (int)(i % ci_n);
            const int co = (int)((i / ci_n) % co_n);
            const int q = (int)(i / ((long long)ci_n * co_n));
            e.fwd[i] = __float2bfloat16_rn(e.w[((long long)ci * co_n + co) * 4 + q]);
        }
        for (long long i = i0; i < total && (which & 2); i += stride) {          // dgrad [ci][q * Co + co]
            const int co = (int)(i % co_n);
            const int q = (int)((i / co_n) % 4);
            const int ci = (int)(i / (4ll * co_n));
            e.dgrad[i] = __float2bfloat16_rn(e.w[((long long)ci * co_n + co) * 4 + q]);
        }
    }
}

__global__ void fold_bn_kernel(const float* __restrict__ conv_bias, const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ mean, const float* __restrict__ var, float eps, int c,
                               float* __restrict__ scale, float* __restrict__ shift) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c) return;
    const float s = gamma[i] / sqrtf(var[i] + eps);
    scale[i] = s;
    shift[i] = fmaf((conv_bias ? conv_bias[i] : 0.f) - mean[i], s, beta[i]);
}

// ------------------------------------------------------------------------------------------------ first layer
// (n,1,h,w) fp32 -> conv3x3 pad 1 (64 filters, fp32 math) -> folded BN -> ReLU -> NHWC bf16.
// HBM-bound on the 128 B/pixel it writes (algorithmic: 4 B in + 128 B out per pixel).  8 lanes per pixel group, 8 output
// channels per lane, C1_RUN consecutive pixels per lane with a sliding 3x(C1_RUN+2) input window.  What limits it below the
// HBM floor is L1 wavefronts, not FMA issue (ncu r1i: l1tex 97 %): a weight LDS.128 costs one wavefront per quarter warp, so
// (a) the weights are laid out [tap][half][lane][4] -- the 8 lanes of a quarter warp read one contiguous 128-byte line instead
// of eight 16-byte pieces 32 bytes apart (2-way bank conflict) -- and (b) every weight vector feeds C1_RUN pixels (8 on wide
// rows: 64 FFMA2 per 2 LDS.128).  Each store instruction of a warp writes four complete 128-byte lines.
// STAGED: the block first copies its row span (+ one halo row either side, one zero column either side, ragged tail zeroed)
// into shared memory, so the pixel loop holds no long-latency load and no column bounds check; rows_per_block / pitch come
// from the host.  !STAGED reads the window straight from global memory (rows too wide for 45 KB of staging).
template <int C1_RUN, bool STAGED>
__global__ void __launch_bounds__(256, 2)
conv3x3_c1_kernel(const float* __restrict__ x, int n, int h, int w, int rows_per_block, int pitch, const float* __restrict__ weight,
                  const float* __restrict__ scale, const float* __restrict__ shift, float relu_floor, uint4* __restrict__ out) {
    extern __shared__ __align__(16) float s_x[];            // STAGED: [rows_per_block + 2][pitch]
    __shared__ __align__(16) float s_w[9][2][8][4];
    for (int i = threadIdx.x; i < 576; i += blockDim.x) {                                  // weight[c][tap]
        const int c = i / 9;
        s_w[i % 9][(c >> 2) & 1][c >> 3][c & 3] = weight[i];
    }
    const int cg = threadIdx.x & 7, grp = threadIdx.x >> 3;
    float2 sc[4], sh[4];                                    // channel pairs: the math below is packed fp32x2 (FFMA2)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        sc[c] = make_float2(__ldg(scale + cg * 8 + 2 * c), __ldg(scale + cg * 8 + 2 * c + 1));
        sh[c] = make_float2(__ldg(shift + cg * 8 + 2 * c), __ldg(shift + cg * 8 + 2 * c + 1));
    }
    // a block owns a contiguous span of rows and its 32 pixel groups stride through the span's (row, run) units, so the
    // ragged last run of a row (w = 1034: two of 130) does not leave one warp with an extra pass per row
    const int n_rows = n * h;
    const int row0 = blockIdx.x * rows_per_block;
    const int runs_per_row = (w + C1_RUN - 1) / C1_RUN;
    const int n_own = max(0, min(rows_per_block, n_rows - row0));
    const int n_units = n_own * runs_per_row;
    if (STAGED) {
        for (int r = 0; r < n_own + 2; ++r) {
            const int g = row0 - 1 + r;
            const bool gok = g >= 0 && g < n_rows;
            const float* __restrict__ src = x + (long long)g * w - 1;
#pragma unroll 5
            for (int j = threadIdx.x; j < pitch; j += 256) s_x[r * pitch + j] = (gok && j >= 1 && j <= w) ? __ldg(src + j) : 0.f;
        }
    }
    __syncthreads();
    {
        for (int u = grp; u < n_units; u += 32) {
            const int row = row0 + u / runs_per_row;
            const int px0 = (u % runs_per_row) * C1_RUN;
            const int py = row % h;
            const float* __restrict__ xr = x + (long long)row * w;
            const bool rok[3] = {py > 0, true, py + 1 < h};
            uint4* __restrict__ orow = out + (long long)row * w * 8 + cg;
            float v[3][C1_RUN + 2];
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int i = 0; i < C1_RUN + 2; ++i) {
                    const int xx = px0 + i - 1;
                    if (STAGED) v[ky][i] = rok[ky] ? s_x[(row - row0 + ky) * pitch + px0 + i] : 0.f;
                    else        v[ky][i] = (rok[ky] && xx >= 0 && xx < w) ? __ldg(xr + (ky - 1) * w + xx) : 0.f;
                }
            float2 acc[C1_RUN][4];
#pragma unroll
            for (int r = 0; r < C1_RUN; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[r][c] = make_float2(0.f, 0.f);
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const float4 w0 = *reinterpret_cast<const float4*>(s_w[ky * 3 + kx][0][cg]);
                    const float4 w1 = *reinterpret_cast<const float4*>(s_w[ky * 3 + kx][1][cg]);
                    const float2 wv[4] = {make_float2(w0.x, w0.y), make_float2(w0.z, w0.w), make_float2(w1.x, w1.y), make_float2(w1.z, w1.w)};
#pragma unroll
                    for (int r = 0; r < C1_RUN; ++r)
#pragma unroll
                        for (int c = 0; c < 4; ++c) acc[r][c] = pk_fma(pk_bcast(v[ky][r + kx]), wv[c], acc[r][c]);
                }
#pragma unroll
            for (int r = 0; r < C1_RUN; ++r) {
                if (px0 + r < w) {
                    uint32_t pk[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float2 y = pk_fma(acc[r][c], sc[c], sh[c]);
                        pk[c] = pack_bf16x2(fmaxf(y.x, relu_floor), fmaxf(y.y, relu_floor));
                    }
                    __stcs(orow + (long long)(px0 + r) * 8, make_uint4(pk[0], pk[1], pk[2], pk[3]));
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ max-pool
__device__ __forceinline__ uint32_t bmax2(uint32_t a, uint32_t b) {
    __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
}
__device__ __forceinline__ uint4 bmax8(uint4 a, uint4 b) {
    return make_uint4(bmax2(a.x, b.x), bmax2(a.y, b.y), bmax2(a.z, b.z), bmax2(a.w, b.w));
}

__global__ void __launch_bounds__(256)
maxpool2x2_kernel(const uint4* __restrict__ src, int n, int h, int w, int c8, uint4* __restrict__ dst) {
    const int ho = h >> 1, wo = w >> 1;
    const long long total = (long long)n * ho * wo * c8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % c8);
        const int xo = (int)((i / c8) % wo);
        const int yo = (int)((i / ((long long)c8 * wo)) % ho);
        const long long img = i / ((long long)c8 * wo * ho);
        const long long r0 = ((img * h + 2 * yo) * w + 2 * xo) * c8 + c;
        const long long r1 = r0 + (long long)w * c8;
        const uint4 a = __ldg(src + r0), b = __ldg(src + r0 + c8), cc = __ldg(src + r1), d = __ldg(src + r1 + c8);
        dst[i] = bmax8(bmax8(a, b), bmax8(cc, d));
    }
}

// ------------------------------------------------------------------------------------------------ layout converters
__global__ void nhwc_bf16_to_nchw_f32_kernel(const __nv_bfloat16* __restrict__ src, int n, int h, int w, int c, float* __restrict__ dst) {
    const long long total = (long long)n * h * w * c;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int xx = (int)(i % w);
        const int yy = (int)((i / w) % h);
        const int cc = (int)((i / ((long long)w * h)) % c);
        const long long img = i / ((long long)w * h * c);
        dst[i] = __bfloat162float(src[((img * h + yy) * w + xx) * c + cc]);
    }
}

__global__ void nchw_f32_to_nhwc_bf16_kernel(const float* __restrict__ src, int n, int c, int h, int w, __nv_bfloat16* __restrict__ dst) {
    const long long total = (long long)n * h * w * c;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int cc = (int)(i % c);
        const int xx = (int)((i / c) % w);
        const int yy = (int)((i / ((long long)c * w)) % h);
        const long long img = i / ((long long)c * w * h);
        dst[i] = __float2bfloat16_rn(src[((img * c + cc) * h + yy) * w + xx]);
    }
}

// ------------------------------------------------------------------------------------------------ loader transform
__global__ void spec_f16_crop_kernel(const float* __restrict__ src, long long n, int f_in, int t_in, int f_out, int t_out,
                                     float* __restrict__ dst) {
    const long long total = n * f_out * t_out;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(i % t_out);
        const int f = (int)((i / t_out) % f_out);
        const long long img = i / ((long long)t_out * f_out);
        float v = 0.f;
        if (f < f_in && t < t_in) v = __half2float(__float2half_rn(src[(img * f_in + f) * t_in + t]));
        dst[i] = v;
    }
}

// ------------------------------------------------------------------------------------------------ error statistics
__global__ void __launch_bounds__(256)
spec_error_sums_kernel(const float* __restrict__ pred, const float* __restrict__ target, long long count, double* __restrict__ sums) {
    double s_abs = 0.0, s_sig = 0.0, s_err = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
        const float p = pred[i], t = target[i];
        const float d = t - p;
        s_abs += (double)fabsf(d);
        s_sig += (double)t * (double)t;
        s_err += (double)d * (double)d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s_abs += __shfl_xor_sync(0xffffffffu, s_abs, o);
        s_sig += __shfl_xor_sync(0xffffffffu, s_sig, o);
        s_err += __shfl_xor_sync(0xffffffffu, s_err, o);
    }
    __shared__ double sh[3][8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sh[0][warp] = s_abs; sh[1][warp] = s_sig; sh[2][warp] = s_err; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = 0.0;
        for (int k = 0; k < 8; ++k) t += sh[threadIdx.x][k];
        atomicAdd(&sums[threadIdx.x], t);
    }
}

static inline int grid_for(long long total, int block = 256) {
    long long g = (total + block - 1) / block;
    const long long cap = (long long)num_sms() * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace adn

using namespace adn;

extern "C" int adn_pack_conv3x3_weight_bf16(const float* w, int c_out, int c_in, void* packed, void* stream) {
    if (!w || !packed || c_out <= 0 || c_in <= 0) return ADN_ERR_ARG;
    int st = check_device(); if (st != ADN_OK) return st;
    const long long total = (long long)c_out * 9 * c_in;
    pack_conv3x3_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(w, c_out, c_in, (__nv_bfloat16*)packed);
    ADN_LAUNCH_CHECK();
    return ADN_OK;
}

extern "C" int adn_pack_convt2x2_weight_bf16(const float* w, int c_in, int c_out, void* packed, void* stream) {
    if (!w || !packed || c_out <= 0 || c_in <= 0) return ADN_ERR_ARG;
    int st = check_device(); if (st != ADN_OK) return st;
    const long long total = 4ll * c_out * c_in;
    pack_convt2x2_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(w, c_in, c_out, (__nv_bfloat16*)packed);
    ADN_LAUNCH_CHECK();
    return ADN_OK;
}

extern "C" int adn_pack_conv3x3_dgrad_weight_bf16(const float* w, int c_out, int c_in, void* packed, void* stream) {
    if (!w || !packed || c_out <= 0 || c_in <= 0) return ADN_ERR_ARG;
    int st = check_device(); if (st != ADN_OK) return st;
    pack_conv3x3_dgrad_kernel<<<grid_for((long long)c_out * 9 * c_in), 256, 0, (cudaStream_t)stream>>>(w, c_out, c_in, (__nv_bfloat16*)packed);
    ADN_LAUNCH_CHECK();
    return ADN_OK;
}

extern "C" int adn_pack_convt2x2_dgrad_weight_bf16(const float* w, int c_in, int c_out, void* packed, void* stream) {
    if (!w || !packed || c_out <= 0 || c_in <= 0) return ADN_ERR_ARG;
    int st = check_device(); if (st != ADN_OK) return st;
    pack_convt2x2_dgrad_kernel<<<grid_for(4ll * c_out * c_in), 256, 0, (cudaStream_t)stream>>>(w, c_in, c_out, (__nv_bfloat16*)packed);
    ADN_LAUNCH_CHECK();
    return ADN_OK;
}

extern "C" int adn_pack_weights_table_sel_bf16(const void* table_dev, int n_entries, int which, void* stream) {
    if (!table_dev || n_entries <= 0 || n_entries > 65535 || which < 1 || which > 3) return ADN_ERR_ARG;
    int st = check_device(); if (st != ADN_OK) return st;
    pack_table_kernel<<<dim3(1024, (unsigned)n_entries), 256, 0, (cudaStream_t)stream>>>(static_cast<const PackEntry*>(table_dev), which, 0);
    ADN_LAUNCH_CHECK();
    return ADN_OK;
}

extern "C" int adn_pack_weights_table_flat_bf16(const void* table_dev, int n_entries, int total_blocks, int which, void* stream) {
    if (!table_dev || n_entries <= 0 || n_entries > 65535 || total_blocks <= 0 || which < 1 || which > 3) return ADN_ERR_ARG;
    int st = check_device(); if (st != ADN_OK) return st;
    pack_table_kernel<<<dim3((unsigned)total_blocks, 1), 256, 0, (cudaStream_t)stream>>>(static_cast<const PackEntry*>(table_dev), which, n_entries);
    ADN_LAUNCH_CHECK();
    return ADN_OK;
}

extern "C" int adn_pack_weights_table_bf16(const void* table_dev, int n_entries, void* stream) {
    return adn_pack_weights_table_sel_bf16(table_dev, n_entries, 3, stream);
}

extern "C" int adn_fold_bn_f32(const float* conv_bias, const float* gamma, const float* beta, const float* mean, const float* var,
                               float eps, int channels, float* scale, float* shift, void* stream) {
    if (!gamma || !beta || !mean || !var || !scale || !shift || channels <= 0) return ADN_ERR_ARG;
    int st = check_device(); if (st != ADN_OK) return st;
    fold_bn_kernel<<<(channels + 127) / 128, 128, 0, (cudaStream_t)stream>>>(conv_bias, gamma, beta, mean, var, eps, channels, scale, shift);
    ADN_LAUNCH_CHECK();
    return ADN_OK;
}

template <int RUN>
static void launch_c1_run(const float* x, int n, int h, int w, const float* weight, const float* scale, const float* shift,
                          float relu_floor, uint4* out, cudaStream_t stream) {
    const int n_rows = n * h, pitch = (w + RUN - 1) / RUN * RUN + 2;
    const int max_rows = (45 * 1024) / (pitch * 4) - 2;                       // 48 KB default limit minus the weight table
    if (max_rows >= 1) {
        const int resident = num_sms() * 2;                                    // 2 CTAs of 256 threads per SM (register-limited)
        int rows = max_rows;
        for (int waves = 1; waves <= 64; ++waves) {                            // fewest whole waves whose span fits the staging
            const int r = (n_rows + resident * waves - 1) / (resident * waves);
            if (r <= max_rows) { rows = r; break; }
        }
        const int grid = (n_rows + rows - 1) / rows;
        conv3x3_c1_kernel<RUN, true><<<grid, 256, (size_t)(rows + 2) * pitch * 4, stream>>>(x, n, h, w, rows, pitch, weight, scale, shift,
                                                                                         relu_floor, out);
    } else {
        const int grid = grid_for((long long)n_rows * w * 8, 256);
        conv3x3_c1_kernel<RUN, false><<<grid, 256, 0, stream>>>(x, n, h, w, (n_rows + grid - 1) / grid, 0, weight, scale, shift, relu_floor, out);
    }
}

static void launch_c1(const float* x, int n, int h, int w, const float* weight, const float* scale, const float* shift,
                      float relu_floor, uint4* out, cudaStream_t stream) {
    if (w >= 512) launch_c1_run<8>(x, n, h, w, weight, scale, shift, relu_floor, out, stream);
    else          launch_c1_run<4>(x, n, h, w, weight, scale, shift, relu_floor, out, stream);
}

// ADN_C1_IMPL=fma keeps the CUDA-core first-layer kernel (A/B measurements); default: the TF32 tensor-core kernel of conv_c1_tc.cu
static bool c1_use_tc() {
    static const bool v = !(getenv("ADN_C1_IMPL") && getenv("ADN_C1_IMPL")[0] == 'f');
    return v;
}

extern "C" int adn_conv3x3_c1_bn_relu_bf16(const float* x, int n, int h, int w, const float* weight, const float* scale,
                                           const float* shift, void* out, void* stream) {
    if (!x || !weight || !scale || !shift || !out || n <= 0 || h <= 0 || w <= 0) return ADN_ERR_ARG;
    int st = check_device(); if (st != ADN_OK) return st;
    if (c1_use_tc() && conv3x3_c1_tc(x, n, h, w, weight, scale, shift, 0.f, out, (cudaStream_t)stream) == ADN_OK) return ADN_OK;
    launch_c1(x, n, h, w, weight, scale, shift, 0.f, (uint4*)out, (cudaStream_t)stream);
    ADN_LAUNCH_CHECK();
    return ADN_OK;
}

// first layer without the activation (train-mode pre-BatchNorm output: scale = 1, shift = conv bias)
extern "C" int adn_conv3x3_c1_affine_bf16(const float* x, int n, int h, int w, const float* weight, const float* scale,
                                          const float* shift, int relu, void* out, void* stream) {
    if (!x || !weight || !scale || !shift || !out || n <= 0 || h <= 0 || w <= 0) return ADN_ERR_ARG;
    int st = check_device(); if (st != ADN_OK) return st;
    if (c1_use_tc() && conv3x3_c1_tc(x, n, h, w, weight, scale, shift, relu ? 0.f : -INFINITY, out, (cudaStream_t)stream) == ADN_OK) return ADN_OK;
    launch_c1(x, n, h, w, weight, scale, shift, relu ? 0.f : -INFINITY, (uint4*)out, (cudaStream_t)stream);
    ADN_LAUNCH_CHECK();
    return ADN_OK;
}

extern "C" int adn_maxpool2x2_bf16(const void* src, int n, int h, int w, int c, void* out, void* stream) {
    if (!src || !out || n <= 0 || h < 2 || w < 2 || c <= 0 || (c & 7)) return ADN_ERR_ARG;
    int st = check_device(); if (st != ADN_OK) return st;
    maxpool2x2_kernel<<<grid_for((long long)n * (h / 2) * (w / 2) * (c / 8)), 256, 0, (cudaStream_t)stream>>>(
        (const uint4*)src, n, h, w, c / 8, (uint4*)out);
    ADN_LAUNCH_CHECK();
    return ADN_OK;
}

extern "C" int adn_nhwc_bf16_to_nchw_f32(const void* src, int n, int h, int w, int c, float* dst, void* stream) {
    if (!src || !dst || n <= 0 || h <= 0 || w <= 0 || c <= 0) return ADN_ERR_ARG;
    int st = check_device(); if (st != ADN_OK) return st;
    nhwc_bf16_to_nchw_f32_kernel<<<grid_for((long long)n * h * w * c), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)src, n, h, w, c, dst);
    ADN_LAUNCH_CHECK();
    return ADN_OK;
}

extern "C" int adn_nchw_f32_to_nhwc_bf16(const float* src, int n, int c, int h, int w, void* dst, void* stream) {
    if (!src || !dst || n <= 0 || h <= 0 || w <= 0 || c <= 0) return ADN_ERR_ARG;
    int st = check_device(); if (st != ADN_OK) return st;
    nchw_f32_to_nhwc_bf16_kernel<<<grid_for((long long)n * h * w * c), 256, 0, (cudaStream_t)stream>>>(src, n, c, h, w, (__nv_bfloat16*)dst);
    ADN_LAUNCH_CHECK();
    return ADN_OK;
}

extern "C" int adn_spec_f16_crop_f32(const float* src, int64_t n, int f_in, int t_in, int f_out, int t_out, float* dst, void* stream) {
    if (n < 0 || f_in <= 0 || t_in <= 0 || f_out <= 0 || t_out <= 0) return ADN_ERR_ARG;
    if (n == 0) return ADN_OK;
    if (!src || !dst) return ADN_ERR_ARG;
    int st = check_device(); if (st != ADN_OK) return st;
    spec_f16_crop_kernel<<<grid_for((long long)n * f_out * t_out), 256, 0, (cudaStream_t)stream>>>(src, n, f_in, t_in, f_out, t_out, dst);
    ADN_LAUNCH_CHECK();
    return ADN_OK;
}

// sums[3] = element count, sums[4] = clip count, sums[5..7] = clip count x (stft, mel, l1) loss terms: the bookkeeping slots of the
// statistics vector that sharding.ShardedDenoiser all-reduces, written by one thread (no eager framework kernels on the path)
__global__ void stats_pack_kernel(double* __restrict__ sums, double numel, double n_clips, const float* __restrict__ loss4) {
    sums[3] = numel;
    if (loss4 != nullptr) {
        sums[4] = n_clips;
        sums[5] = n_clips * (double)loss4[1];
        sums[6] = n_clips * (double)loss4[2];
        sums[7] = n_clips * (double)loss4[3];
    }
}

extern "C" int adn_stats_pack_f64(double* sums8, int64_t numel, int64_t n_clips, const float* loss4, void* stream) {
    if (!sums8 || numel < 0 || n_clips < 0) return ADN_ERR_ARG;
    int st = check_device(); if (st != ADN_OK) return st;
    stats_pack_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(sums8, (double)numel, (double)n_clips, loss4);
    ADN_LAUNCH_CHECK();
    return ADN_OK;
}

extern "C" int adn_spec_error_sums_f64(const float* pred, const float* target, int64_t count, double* sums, void* stream) {
    if (count < 0 || !sums) return ADN_ERR_ARG;
    if (count == 0) return ADN_OK;
    if (!pred || !target) return ADN_ERR_ARG;
    int st = check_device(); if (st != ADN_OK) return st;
    spec_error_sums_kernel<<<grid_for(count), 256, 0, (cudaStream_t)stream>>>(pred, target, count, sums);
    ADN_LAUNCH_CHECK();
    return ADN_OK;
}
