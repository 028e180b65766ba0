// wgrad_tc.cu -- weight gradients of Conv2d 3x3 and ConvTranspose2d 2x2 (the backward of code/model.py:11,14,38 driven by
// loss.backward() at code/train.py:69) as a tcgen05/TMEM GEMM that CONTRACTS OVER PIXELS, for sm_100a.
//
//   GEMM view    D_tap[m, n] += sum_px A[px, m] * B_tap[px, n]        (bf16 x bf16 -> fp32 in TMEM)
//                Conv3x3:   A = dz (gradient of the conv output, m = co), B_tap = the conv input shifted by the tap
//                           (n = ci), three taps (one kernel row ky) per CTA:   dW[co][ci][ky][kx] = D_kx[co][ci];
//                ConvT 2x2: A = the convT input (m = ci), B_q = quadrant q of the output gradient (n = co), four taps:
//                           dW[ci][co][dy][dx] = D_q[ci][co].
//   operands     both are NHWC activations, so the contraction index (pixel) is the SLOW index of the shared-memory tile and
//                the M / N index (channel) is contiguous: MN-major UMMA operands (instruction-descriptor bits 15/16).  A k-step
//                is an 8x8 pixel tile = 64 rows of 128 B; one 4-D TMA box {64 ch, 8, 8, 1} per 64-channel block lands as
//                MN-major SWIZZLE_128B atoms (8 pixel rows x 64 channels = 1024 B): SBO = 1024 B between 8-pixel groups,
//                LBO = 8192 B between 64-channel blocks.  Tap shifts are TMA coordinates; out-of-image pixels arrive as zeros
//                (= the conv's zero padding).
//   split-K      the pixel range is split across CTAs; every CTA stores its fp32 partial tile [split][tap][m][n] to a workspace
//                with 128-byte row segments, and a second kernel folds the splits in ascending order and WRITES the result into
//                the gradient buffer in the reference layout: deterministic, and no scattered atomics (the first version spent
//                5x its GEMM time in red.global.add).
//   roles        warp 0: TMA producer - warp 1: tcgen05.mma issuer - warps 2..5: epilogue (tcgen05.ld -> st.global.v4).
#include "tc_common.cuh"

namespace adn {

constexpr int WG_SUB = 64 * 128;               // one [64 px][64 ch] bf16 sub-tile
constexpr int WG_THREADS = 192;
constexpr int WG_MAX_STAGES = 6;
constexpr int WG_MAX_TAPS = 4;                 // taps per CTA (TMEM: 4 x 128 columns)
constexpr int WG_MAX_ALL_TAPS = 9;

struct WgradArgs {
    int tiles_x, tiles_y, kt_total;            // 8x8 pixel tiling of the (n, h, w) grid
    int m_blocks, n_blocks, groups, splits;    // blockIdx.x = ((mb * n_blocks + nb) * groups + grp) * splits + s
    int taps, stages;                          // taps per group (a CTA accumulates one group); groups * taps taps in all
    int tap_dy[WG_MAX_ALL_TAPS], tap_dx[WG_MAX_ALL_TAPS], tap_map[WG_MAX_ALL_TAPS];
    long long tap_off[WG_MAX_ALL_TAPS];        // output element offset of each tap
    int m_total, n_total;
    long long sm, sn;                          // output strides in elements
    float* out;
    float* partial;                            // workspace [splits][groups * taps][m_pad][n_pad]
    int m_pad, n_pad;
    uint32_t lbo, sbo;                         // MN-major descriptor strides (bytes)
};

__device__ __forceinline__ uint64_t make_mn_sw128_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc_mn(int n) { return make_idesc(n) | (1u << 15) | (1u << 16); }

template <int BN>
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1,
             const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmB3, const WgradArgs a) {
    constexpr int NSUB = BN / 64;
    extern __shared__ uint8_t smem_dyn[];
    const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_dyn + (smem_base - smem_u32(smem_dyn));
    const uint32_t stage_bytes = (uint32_t)(2 + a.taps * NSUB) * WG_SUB;
    const uint32_t aux_off = (uint32_t)a.stages * stage_bytes;
    const uint32_t bar_base = smem_base + aux_off;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (WG_MAX_STAGES + s); };
    const uint32_t tfull = bar_base + 8u * (2 * WG_MAX_STAGES);
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem_gen + aux_off + (2 * WG_MAX_STAGES + 1) * 8);
    constexpr int TMEM_COLS = (WG_MAX_TAPS * BN) < 32 ? 32 : WG_MAX_TAPS * BN;      // 256 or 512

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB0); tma_prefetch_desc(&tmB1); tma_prefetch_desc(&tmB2); tma_prefetch_desc(&tmB3);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < WG_MAX_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(tfull, 1);
        fence_barrier_init();
    }
    if (warp == 0) { tmem_alloc(smem_u32(tmem_ptr_smem), TMEM_COLS); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const int s_idx = blockIdx.x % a.splits;
    const int rest = blockIdx.x / a.splits;
    const int grp = rest % a.groups, mn = rest / a.groups;
    const int nb = mn % a.n_blocks, mb = mn / a.n_blocks;
    const int tap0 = grp * a.taps;                                       // first tap of this CTA's group
    const int per = (a.kt_total + a.splits - 1) / a.splits;
    const int kt0 = s_idx * per, kt1 = min(a.kt_total, kt0 + per);
    const int tiles_per_img = a.tiles_x * a.tiles_y;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int kt = kt0; kt < kt1; ++kt) {
                const int img = kt / tiles_per_img;
                const int rem = kt - img * tiles_per_img;
                const int ty = rem / a.tiles_x, tx = rem - ty * a.tiles_x;
                const int x0 = tx * 8, y0 = ty * 8;
                mbar_wait(empty_bar(stage), phase ^ 1u);
                mbar_arrive_expect_tx(full_bar(stage), stage_bytes);
                const uint32_t sa = smem_base + (uint32_t)stage * stage_bytes;
                tma_load_4d(sa, &tmA, full_bar(stage), mb * 128, x0, y0, img);
                tma_load_4d(sa + WG_SUB, &tmA, full_bar(stage), mb * 128 + 64, x0, y0, img);
                for (int t = 0; t < a.taps; ++t) {
                    const int gt = tap0 + t;
                    const CUtensorMap* mp = (a.tap_map[gt] == 0) ? &tmB0 : (a.tap_map[gt] == 1) ? &tmB1 : (a.tap_map[gt] == 2) ? &tmB2 : &tmB3;
                    for (int j = 0; j < NSUB; ++j)
                        tma_load_4d(sa + (uint32_t)(2 + t * NSUB + j) * WG_SUB, mp, full_bar(stage), nb * BN + j * 64, x0 + a.tap_dx[gt],
                                    y0 + a.tap_dy[gt], img);
                }
                if (++stage == a.stages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // The taps of a group sit in consecutive 8 KB sub-tiles, i.e. they ARE one MN-major operand with taps*BN columns (LBO = 8 KB
        // between 64-column blocks): one UMMA per k-slice covers up to 256 columns (BN = 64: all 3 or 4 taps, N = 192 / 256;
        // BN = 128: two taps, then the rest) instead of one N = BN instruction per tap -- per-MMA smem operand traffic drops from
        // A + B to A + B over 3-4x the work, which took the 64-wide layers off the shared-memory read limit.
        const int taps_first = (a.taps * BN <= 256) ? a.taps : 256 / BN;
        const int taps_rest = a.taps - taps_first;
        const uint32_t idesc0 = make_idesc_mn(taps_first * BN);
        const uint32_t idesc1 = make_idesc_mn(taps_rest > 0 ? taps_rest * BN : BN);
        int stage = 0; uint32_t phase = 0;
        for (int kt = kt0; kt < kt1; ++kt) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint32_t sa = smem_base + (uint32_t)stage * stage_bytes;
            if (elect_one()) {
                const uint32_t sb0 = sa + 2u * WG_SUB;
                const uint32_t sb1 = sb0 + (uint32_t)(taps_first * NSUB) * WG_SUB;
#pragma unroll
                for (int k = 0; k < 4; ++k) {                // UMMA_K = 16 pixels = 16 rows of 128 B
                    const uint64_t da = make_mn_sw128_desc(sa + k * 2048, a.lbo, a.sbo);
                    const uint32_t accf = (kt > kt0 || k > 0) ? 1u : 0u;
                    umma_bf16(tmem_base, da, make_mn_sw128_desc(sb0 + k * 2048, a.lbo, a.sbo), idesc0, accf);
                    if (taps_rest > 0)
                        umma_bf16(tmem_base + (uint32_t)(taps_first * BN), da, make_mn_sw128_desc(sb1 + k * 2048, a.lbo, a.sbo), idesc1, accf);
                }
                umma_commit(empty_bar(stage));
            }
            __syncwarp();
            if (++stage == a.stages) { stage = 0; phase ^= 1u; }
        }
        if (elect_one()) umma_commit(tfull);
        __syncwarp();
    } else if (kt1 > kt0) {
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const int m = mb * 128 + row;
        mbar_wait(tfull, 0);
        tc_fence_after();
        for (int t = 0; t < a.taps; ++t) {
            float4* o = reinterpret_cast<float4*>(a.partial + (((long long)s_idx * (a.groups * a.taps) + tap0 + t) * a.m_pad + m) * a.n_pad + nb * BN);
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(t * BN + c0), r);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    o[(c0 >> 2) + i] = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]),
                                                   __uint_as_float(r[4 * i + 3]));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS);
}

// out[m * sm + n * sn + tap_off[0] + t] = sum over splits (ascending) of partial[s][t][m][n]  (taps are innermost and contiguous
// in both reference layouts: sn == number of taps).  One block = one output row m x up to 256 columns n: the partial reads are
// coalesced along n, the results are transposed through shared memory and leave as ONE contiguous run of 256 x taps floats.
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const WgradArgs a) {
    __shared__ float s_out[256 * WG_MAX_ALL_TAPS];
    const int all_taps = a.groups * a.taps;
    const long long plane = (long long)a.m_pad * a.n_pad;
    const int chunks = (a.n_total + 255) / 256;
    for (int blk = blockIdx.x; blk < a.m_total * chunks; blk += gridDim.x) {
        const int m = blk / chunks, n0 = (blk - m * chunks) * 256;
        const int cols = min(256, a.n_total - n0);
        const int n = n0 + threadIdx.x;
        if ((int)threadIdx.x < cols) {
            const float* p = a.partial + (long long)m * a.n_pad + n;
            for (int t = 0; t < all_taps; ++t) {
                float acc = 0.f;
                for (int s0 = 0; s0 < a.splits; s0 += 4) {             // four loads in flight; additions stay in ascending split order
                    float v[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) v[j] = s0 + j < a.splits ? p[((long long)(s0 + j) * all_taps + t) * plane] : 0.f;
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc += v[j];
                }
                s_out[threadIdx.x * all_taps + t] = acc;
            }
        }
        __syncthreads();
        float* o = a.out + (long long)m * a.sm + (long long)n0 * a.sn + a.tap_off[0];
        for (int i = threadIdx.x; i < cols * all_taps; i += 256) o[i] = s_out[i];
        __syncthreads();
    }
}

// Same fold for the layers with MANY splits and a small gradient (the full-resolution levels: 148 partial tiles for a 64 x 64 x 9
// gradient): one block per (output row m, tap, 64 columns); its 256 threads are 64 columns x 4 split groups, so the partial reads
// are spread over 9 x 4 times more threads than above.  Split group j adds splits j, j+4, ... in ascending order and the four
// group sums are combined in a fixed order: deterministic.
__global__ void __launch_bounds__(256)
wgrad_reduce_splits_kernel(const WgradArgs a) {
    __shared__ float s_part[4][64];
    const int all_taps = a.groups * a.taps;
    const long long plane = (long long)a.m_pad * a.n_pad;
    const int chunks = (a.n_total + 63) / 64;
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    for (int blk = blockIdx.x; blk < a.m_total * all_taps * chunks; blk += gridDim.x) {
        const int ch = blk % chunks;
        const int t = (blk / chunks) % all_taps;
        const int m = blk / (chunks * all_taps);
        const int n = ch * 64 + tx;
        float acc = 0.f;
        if (n < a.n_total) {
            const float* p = a.partial + (long long)t * plane + (long long)m * a.n_pad + n;
            for (int s0 = ty; s0 < a.splits; s0 += 32) {               // eight loads in flight; additions stay in ascending split order
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = s0 + 4 * j < a.splits ? p[(long long)(s0 + 4 * j) * all_taps * plane] : 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) acc += v[j];
            }
        }
        s_part[ty][tx] = acc;
        __syncthreads();
        if (ty == 0 && n < a.n_total)
            a.out[(long long)m * a.sm + (long long)n * a.sn + a.tap_off[t]] = ((s_part[0][tx] + s_part[1][tx]) + s_part[2][tx]) + s_part[3][tx];
        __syncthreads();
    }
}

// MN-major SWIZZLE_128B descriptor strides, confirmed by a sweep on B200 (profiles/README.md): LBO = bytes between 64-channel
// blocks, SBO = bytes between 8-pixel row groups; every other combination produces garbage.
constexpr uint32_t g_wg_lbo = 8192, g_wg_sbo = 1024;

// a channel slice [c_off, c_off + c) of an NHWC tensor whose pixels are `ld` channels apart, sampled on the (h, w) grid with
// pixel steps (sy, sx) starting at (oy, ox): element (ch, x, y, img) = base[((img * H_full + oy + sy*y) * W_full + ox + sx*x) * ld + c_off + ch]
static int make_view_map(CUtensorMap* map, const void* ptr, int n, int h, int w, int c, int ld, int c_off, int h_full, int w_full,
                         int sy, int sx, int oy, int ox) {
    PFN_tmapEncodeTiled enc = get_encode_fn();
    if (!enc) return ADN_ERR_DRIVER;
    const char* base = static_cast<const char*>(ptr) + (((size_t)oy * w_full + ox) * ld + c_off) * 2;
    cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
    cuuint64_t strides[3] = {(cuuint64_t)sx * ld * 2, (cuuint64_t)sy * w_full * ld * 2, (cuuint64_t)h_full * w_full * ld * 2};
    cuuint32_t box[4] = {64u, 8u, 8u, 1u};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<char*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? ADN_OK : ADN_ERR_DRIVER;
}

constexpr long long WG_WORKSPACE_BYTES = 96ll << 20;

static int launch_wgrad(const CUtensorMap& mA, const CUtensorMap* mB, WgradArgs& args, int n, int h, int w, void* workspace, cudaStream_t stream) {
    if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 15)) return ADN_ERR_ARG;
    args.tiles_x = (w + 7) / 8; args.tiles_y = (h + 7) / 8;
    const long long kt = (long long)n * args.tiles_x * args.tiles_y;
    if (kt > 0x7fffffffLL) return ADN_ERR_ARG;
    args.kt_total = (int)kt;
    const int bn = (args.n_total % 128 == 0) ? 128 : 64;
    args.m_blocks = (args.m_total + 127) / 128;
    args.n_blocks = (args.n_total + bn - 1) / bn;
    const int tiles = args.m_blocks * args.n_blocks * args.groups;
    // split K so that one wave of CTAs covers the chip, but keep >= 12 k-steps per CTA: every CTA writes a full fp32 partial
    // tile, so over-splitting a short K turns the GEMM into a workspace-bandwidth problem
    int splits = (num_sms() + tiles - 1) / tiles;
    if (splits > args.kt_total / 12) splits = args.kt_total / 12;
    if (splits < 1) splits = 1;
    args.m_pad = args.m_blocks * 128; args.n_pad = args.n_blocks * bn;
    const long long per_split = (long long)args.groups * args.taps * args.m_pad * args.n_pad * 4;
    if (per_split > WG_WORKSPACE_BYTES) return ADN_ERR_ARG;
    if ((long long)splits * per_split > WG_WORKSPACE_BYTES) splits = (int)(WG_WORKSPACE_BYTES / per_split);
    // every split must own at least one k-step
    while (splits > 1 && (long long)((args.kt_total + splits - 1) / splits) * (splits - 1) >= args.kt_total) --splits;
    args.splits = splits;
    args.partial = static_cast<float*>(workspace);
    args.lbo = g_wg_lbo; args.sbo = g_wg_sbo;
    const int stage_bytes = (2 + args.taps * (bn / 64)) * WG_SUB;
    int stages = (225 * 1024 - 2048) / stage_bytes;
    if (stages > WG_MAX_STAGES) stages = WG_MAX_STAGES;
    if (stages < 2) return ADN_ERR_ARG;
    args.stages = stages;
    const int smem = 1024 + stages * stage_bytes + (2 * WG_MAX_STAGES + 1) * 8 + 16;
    const int grid = tiles * splits;
    if (bn == 128) {
        static unsigned char smem_set[64] = {0};
        ADN_CUDA_TRY(ensure_dyn_smem(wgrad_kernel<128>, 232448, smem_set));
        wgrad_kernel<128><<<grid, WG_THREADS, smem, stream>>>(mA, mB[0], mB[1], mB[2], mB[3], args);
    } else {
        static unsigned char smem_set[64] = {0};
        ADN_CUDA_TRY(ensure_dyn_smem(wgrad_kernel<64>, 232448, smem_set));
        wgrad_kernel<64><<<grid, WG_THREADS, smem, stream>>>(mA, mB[0], mB[1], mB[2], mB[3], args);
    }
    ADN_LAUNCH_CHECK();
    if (args.sn != args.groups * args.taps) return ADN_ERR_ARG;          // the fold writes contiguous tap runs
    const long long cap = (long long)num_sms() * 16;
    if (args.splits >= 8) {
        long long rg = (long long)args.m_total * args.groups * args.taps * ((args.n_total + 63) / 64); if (rg > cap) rg = cap;
        wgrad_reduce_splits_kernel<<<(int)rg, 256, 0, stream>>>(args);
    } else {
        long long rg = (long long)args.m_total * ((args.n_total + 255) / 256); if (rg > cap) rg = cap;
        wgrad_reduce_kernel<<<(int)rg, 256, 0, stream>>>(args);
    }
    ADN_LAUNCH_CHECK();
    return ADN_OK;
}

}  // namespace adn

using namespace adn;

extern "C" int64_t adn_wgrad_workspace_bytes(void) { return WG_WORKSPACE_BYTES; }

// d_weight[(co * ci_total + ci_off + ci) * 9 + ky * 3 + kx] = sum_p dz[p][co] * x[p + (ky-1, kx-1)][ci]   (reference layout
// (Co, Ci, 3, 3) of model.py:11,14).  dz: (n,h,w,c_out) dense; x: (n,h1,w1,c_in) dense, (h1,w1) <= (h,w) (a zero-padded
// up-sampled map, model.py:44-47).  ci_off / ci_total address the slice of a concatenated input (model.py:49).
extern "C" int adn_conv3x3_wgrad_f32(const void* dz, int c_out, const void* x, int c_in, int h1, int w1, int n, int h, int w,
                                     float* d_weight, int ci_off, int ci_total, void* workspace, void* stream) {
    if (!dz || !x || !d_weight || n <= 0 || h <= 0 || w <= 0 || c_out <= 0 || (c_out % 64) || c_in <= 0 || (c_in % 64)) return ADN_ERR_ARG;
    if (h1 <= 0 || w1 <= 0 || h1 > h || w1 > w || ci_off < 0 || ci_off + c_in > ci_total) return ADN_ERR_ARG;
    if (!aligned16(dz) || !aligned16(x)) return ADN_ERR_ARG;
    int st = check_device(); if (st != ADN_OK) return st;
    CUtensorMap mA, mB[4];
    st = make_view_map(&mA, dz, n, h, w, c_out, c_out, 0, h, w, 1, 1, 0, 0);
    if (st != ADN_OK) return st;
    st = make_view_map(&mB[0], x, n, h1, w1, c_in, c_in, 0, h1, w1, 1, 1, 0, 0);
    if (st != ADN_OK) return st;
    mB[1] = mB[2] = mB[3] = mB[0];
    WgradArgs args{};
    args.groups = 3; args.taps = 3;                       // a CTA accumulates one kernel row (ky): 3 x BN TMEM columns
    for (int t = 0; t < 9; ++t) {
        args.tap_dy[t] = t / 3 - 1; args.tap_dx[t] = t % 3 - 1; args.tap_map[t] = 0;
        args.tap_off[t] = (long long)ci_off * 9 + t;
    }
    args.m_total = c_out; args.n_total = c_in;
    args.sm = (long long)ci_total * 9; args.sn = 9;
    args.out = d_weight;
    return launch_wgrad(mA, mB, args, n, h, w, workspace, (cudaStream_t)stream);
}

// d_weight[(ci * c_out + co) * 4 + dy * 2 + dx] = sum_p x[p][ci] * d_up[2p + (dy, dx)][co]   (reference layout (Ci, Co, 2, 2) of
// model.py:38).  x: (n,h,w,c_in) dense; d_up: channels [up_off, up_off + c_out) of an (n,2h,2w,up_ld) tensor.
extern "C" int adn_convt2x2_wgrad_f32(const void* x, int c_in, const void* d_up, int up_ld, int up_off, int c_out, int n, int h, int w,
                                      float* d_weight, void* workspace, void* stream) {
    if (!x || !d_up || !d_weight || n <= 0 || h <= 0 || w <= 0 || c_out <= 0 || (c_out % 64) || c_in <= 0 || (c_in % 64)) return ADN_ERR_ARG;
    if (up_off < 0 || up_off + c_out > up_ld || (up_ld % 8) || (up_off % 8)) return ADN_ERR_ARG;
    if (!aligned16(x) || !aligned16(d_up)) return ADN_ERR_ARG;
    int st = check_device(); if (st != ADN_OK) return st;
    CUtensorMap mA, mB[4];
    st = make_view_map(&mA, x, n, h, w, c_in, c_in, 0, h, w, 1, 1, 0, 0);
    if (st != ADN_OK) return st;
    WgradArgs args{};
    args.groups = 1; args.taps = 4;
    for (int q = 0; q < 4; ++q) {
        st = make_view_map(&mB[q], d_up, n, h, w, c_out, up_ld, up_off, 2 * h, 2 * w, 2, 2, q >> 1, q & 1);
        if (st != ADN_OK) return st;
        args.tap_dy[q] = 0; args.tap_dx[q] = 0; args.tap_map[q] = q; args.tap_off[q] = q;
    }
    args.m_total = c_in; args.n_total = c_out;
    args.sm = (long long)c_out * 4; args.sn = 4;
    args.out = d_weight;
    return launch_wgrad(mA, mB, args, n, h, w, workspace, (cudaStream_t)stream);
}
