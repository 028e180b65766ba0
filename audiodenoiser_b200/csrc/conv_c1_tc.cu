// conv_c1_tc.cu -- the FIRST conv of the reference UNet (code/model.py:11 with in_channels = 1: Conv2d(1, 64, 3, padding=1)) + folded
// BatchNorm + ReLU as a TF32 tcgen05 implicit GEMM, for sm_100a.
//
// Why a tensor-core kernel for a K = 9 contraction.  The layer writes 128 bytes per pixel (64 bf16 channels, 2.2 GB for the
// benchmark batch): its floor is the HBM write time (0.34 ms).  On the CUDA cores it costs 576 FMAs per pixel = 18 issue slots per
// pixel and warp -- the round-1 kernel (unet_misc.cu: conv3x3_c1_kernel) ran 0.70-0.77 ms, co-limited by the fp32 pipe, L1 and HBM.
// Measured (batch 64 x (257,1034)): 0.505 ms alone under ncu (4.3 TB/s written) against 0.70-0.77; inside the step 0.66-0.69 against
// 0.70-0.78 -- there every write-heavy kernel of the network sits near 3.2 TB/s of stores, whatever issues them (profiles/README.md).
// As a GEMM it is one 128 x 64 x 16 tile per 128 pixels: D[128 px, 64 ch] = A[128 px, K = 16] * B[64 ch, 16]^T with A the im2col
// rows (9 taps, 7 zeros) and B the weights, two UMMAs of K = 8 per tile.  TF32 keeps 10 explicit significand bits of the fp32 input
// and weights (round to nearest: cvt.rna.tf32.f32), i.e. an operand error of 2^-11 against the 2^-9 of every bf16 layer that follows
// -- the input magnitudes are NOT rounded to bf16 (that would spend a visible share of the 1e-2 budget on the first layer).
//
//   roles        warp 0: tcgen05.mma issuer (TMEM owner) - warps 1..4: im2col producers, one pixel per thread and tile: nine bounds-
//                checked loads of the fp32 input (L1 serves the 3x3 overlap), three 16-byte swizzled stores into the tile's 128-byte
//                row, fence.proxy.async, one mbarrier arrival per warp - warps 5..12: epilogue, two sets of four warps (tcgen05.ld -> scale / shift ->
//                ReLU folded into the bf16 conversion -> SWIZZLE_128B staging tile -> ONE TMA store {64 ch, TW, TH, 1} per tile; image
//                edges are clipped by the TMA engine).
//   layout       A and B rows are 128 bytes (32 TF32 slots, 16 used), SWIZZLE_128B K-major: byte for byte the operand layout of the
//                bf16 kernels (tc_common.cuh: make_sw128_desc), K advances 32 bytes per UMMA.
//   pipeline     ring of C1_STAGES A tiles, two TMEM accumulators (128 columns), two output staging tiles; two CTAs per SM keep
//                enough global loads in flight (a producer thread waits a full memory latency per tile).
#include <cstdlib>
#include "tc_common.cuh"

namespace adn {

#ifndef ADN_C1_TW
#define ADN_C1_TW 128
#endif
constexpr int C1_TW = ADN_C1_TW, C1_TH = 128 / C1_TW;     // pixel tile (128 pixels = the 128 TMEM lanes): a run of one image row
constexpr int C1_STAGES = 4;
constexpr int C1_A_BYTES = 128 * 128;                     // [128 px][128 B]
constexpr int C1_B_BYTES = 64 * 128;                      // [64 ch][128 B]
constexpr int C1_OUT_BYTES = 128 * 128;                   // [128 px][64 ch] bf16
constexpr int c1_threads(int sets) { return 32 + 128 + 128 * sets; }   // SETS epilogue sets of four warps: set e converts 64 / SETS columns
constexpr int C1_AUX_BYTES = 2 * 64 * 4 + (2 * C1_STAGES + 4) * 8 + 16;
constexpr int C1_SMEM_BYTES = C1_STAGES * C1_A_BYTES + C1_B_BYTES + 2 * C1_OUT_BYTES + C1_AUX_BYTES + 1024;

struct C1Args {
    const float* x;            // (n, 1, h, w) fp32
    const float* weight;       // (64, 1, 3, 3) fp32
    const float* scale;        // [64]
    const float* shift;        // [64]
    float relu_floor;          // 0 = ReLU, -inf = none
    int n_img, H, W;
    int tiles_x, tiles_y, num_tiles;
    FastDiv div_tpi, div_tx;
};

__device__ __forceinline__ float to_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// instruction descriptor, kind::tf32: D = f32 (bits 4-5 = 1), A = B = TF32 (format 2 in bits 7-9 / 10-12), K-major, N = 64, M = 128
__host__ __device__ constexpr uint32_t c1_idesc() {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

template <int SETS>
__global__ void __launch_bounds__(c1_threads(SETS), 2)
conv3x3_c1_tc_kernel(const __grid_constant__ CUtensorMap tmOut, const C1Args a) {
    extern __shared__ uint8_t smem_dyn[];
    const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_dyn + (smem_base - smem_u32(smem_dyn));
    // carve-up: [STAGES x A] [B] [2 x output staging] [scale 64] [shift 64] [a_full S] [a_empty S] [t_full 2] [t_empty 2] [tmem ptr]
    const uint32_t b_base = smem_base + C1_STAGES * C1_A_BYTES;
    const uint32_t out_base = b_base + C1_B_BYTES;
    constexpr uint32_t AUX_OFF = C1_STAGES * C1_A_BYTES + C1_B_BYTES + 2 * C1_OUT_BYTES;
    float* s_scale = reinterpret_cast<float*>(smem_gen + AUX_OFF);
    float* s_shift = s_scale + 64;
    const uint32_t bar_base = smem_base + AUX_OFF + 2 * 64 * 4;
    auto a_full = [&](int s) { return bar_base + 8u * s; };
    auto a_empty = [&](int s) { return bar_base + 8u * (C1_STAGES + s); };
    auto t_full = [&](int s) { return bar_base + 8u * (2 * C1_STAGES + s); };
    auto t_empty = [&](int s) { return bar_base + 8u * (2 * C1_STAGES + 2 + s); };
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem_gen + AUX_OFF + 2 * 64 * 4 + (2 * C1_STAGES + 4) * 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- one-time setup: weights -> B tile (TF32, zero-padded K), zero the never-written parts of the A tiles, barriers, TMEM
    constexpr int C1_THREADS = c1_threads(SETS), C1_EPI_THREADS = 128 * SETS, CW = 64 / SETS;   // CW: columns per epilogue thread
    for (int i = threadIdx.x; i < 64 * 8; i += C1_THREADS) {         // 16-byte chunk j of row co: k = 4 j .. 4 j + 3
        const int co = i >> 3, j = i & 7;
        float v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { const int k = 4 * j + q; v[q] = k < 9 ? to_tf32(a.weight[co * 9 + k]) : 0.f; }
        st_shared_v4(b_base + (uint32_t)co * 128u + (((uint32_t)j ^ ((uint32_t)co & 7u)) << 4), __float_as_uint(v[0]), __float_as_uint(v[1]),
                     __float_as_uint(v[2]), __float_as_uint(v[3]));
    }
    for (int i = threadIdx.x; i < C1_STAGES * 128 * 8; i += C1_THREADS) {     // chunks 3..7 of every A row stay zero for the whole kernel
        const int s = i >> 10, m = (i >> 3) & 127, j = i & 7;
        if (j >= 3) st_shared_v4(smem_base + (uint32_t)s * C1_A_BYTES + (uint32_t)m * 128u + (((uint32_t)j ^ ((uint32_t)m & 7u)) << 4), 0u, 0u, 0u, 0u);
    }
    if (threadIdx.x < 64) { s_scale[threadIdx.x] = a.scale[threadIdx.x]; s_shift[threadIdx.x] = a.shift[threadIdx.x]; }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmOut);
        for (int s = 0; s < C1_STAGES; ++s) { mbar_init(a_full(s), 4); mbar_init(a_empty(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(t_full(s), 1); mbar_init(t_empty(s), C1_EPI_THREADS / 32); }
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(smem_u32(tmem_ptr_smem), 128);
        tmem_relinquish();
    }
    fence_proxy_async();                                             // the B tile / zero chunks were written through the generic proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        // ===================================================================== MMA issuer
        constexpr uint32_t idesc = c1_idesc();
        const uint64_t db = make_sw128_desc(b_base);
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
            mbar_wait(t_empty(acc), acc_phase ^ 1u);                  // the epilogue has drained this accumulator
            mbar_wait(a_full(stage), phase);                          // the producers have written this A tile
            tc_fence_after();
            const uint64_t da = make_sw128_desc(smem_base + stage * C1_A_BYTES);
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 64);
            if (elect_one()) {
                umma_tf32(d_tmem, da, db, idesc, 0u);                 // k = 0..7
                umma_tf32(d_tmem, da + 2u, db + 2u, idesc, 1u);       // k = 8..15: +32 bytes inside the swizzle row
                umma_commit(a_empty(stage));
                umma_commit(t_full(acc));
            }
            __syncwarp();
            if (++stage == C1_STAGES) { stage = 0; phase ^= 1u; }
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
    } else if (warp <= 4) {
        // ===================================================================== im2col producers: thread m = pixel m of the tile
        const int m = threadIdx.x - 32;                               // 0..127
        const int py = m / C1_TW, px = m % C1_TW;
        const uint32_t row_off = (uint32_t)m * 128u, sw = (uint32_t)m & 7u;
        int stage = 0; uint32_t phase = 0;
        // A producer thread would otherwise wait one full memory latency per tile (ncu: long_scoreboard 8.8 warps per issue, the
        // kernel ran at 1 680 cycles per tile and SM).  The nine loads of a tile are therefore issued TWO tiles ahead: three register
        // sets rotate through (in flight, in flight, being stored).
        auto load_tile = [&](int tile, float (&v)[9]) {
            if (tile >= a.num_tiles) return;
            const int img = fast_div(tile, a.div_tpi);
            const int rem = tile - img * a.tiles_x * a.tiles_y;
            const int ty = fast_div(rem, a.div_tx), tx = rem - ty * a.tiles_x;
            const int y = ty * C1_TH + py, x = tx * C1_TW + px;
            const float* src = a.x + ((long long)img * a.H + y) * a.W + x;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int yy = y + ky - 1, xx = x + kx - 1;
                    const bool ok = (unsigned)yy < (unsigned)a.H && (unsigned)xx < (unsigned)a.W;     // zero padding (pixels past the edge too)
                    v[ky * 3 + kx] = ok ? __ldg(src + (ky - 1) * a.W + (kx - 1)) : 0.f;
                }
        };
        auto store_tile = [&](const float (&v)[9]) {
            mbar_wait(a_empty(stage), phase ^ 1u);                    // the MMAs that read this slot have retired
            const uint32_t base = smem_base + (uint32_t)stage * C1_A_BYTES + row_off;
            st_shared_v4(base + ((0u ^ sw) << 4), __float_as_uint(to_tf32(v[0])), __float_as_uint(to_tf32(v[1])), __float_as_uint(to_tf32(v[2])),
                         __float_as_uint(to_tf32(v[3])));
            st_shared_v4(base + ((1u ^ sw) << 4), __float_as_uint(to_tf32(v[4])), __float_as_uint(to_tf32(v[5])), __float_as_uint(to_tf32(v[6])),
                         __float_as_uint(to_tf32(v[7])));
            st_shared_v4(base + ((2u ^ sw) << 4), __float_as_uint(to_tf32(v[8])), 0u, 0u, 0u);
            fence_proxy_async();                                      // generic-proxy writes -> visible to the tensor core's async proxy
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full(stage));
            if (++stage == C1_STAGES) { stage = 0; phase ^= 1u; }
        };
        float va[9], vb[9], vc[9];
        const int step = (int)gridDim.x;
        int tile = blockIdx.x;
        load_tile(tile, va);
        load_tile(tile + step, vb);
        for (; tile < a.num_tiles; tile += 3 * step) {
            load_tile(tile + 2 * step, vc);
            store_tile(va);
            if (tile + step >= a.num_tiles) break;
            load_tile(tile + 3 * step, va);
            store_tile(vb);
            if (tile + 2 * step >= a.num_tiles) break;
            load_tile(tile + 4 * step, vb);
            store_tile(vc);
        }
    } else {
        // ===================================================================== epilogue: 2 sets x 4 warps.  A set covers the 128 TMEM
        // lanes (warp % 4 = lane quadrant); set e converts the 32 columns [32 e, 32 e + 32) of every tile, i.e. the 64-byte half e
        // of each staged 128-byte pixel row, so a tile's conversion latency is halved; one thread issues the tile's TMA store.
        const int quad = warp & 3;                                    // TMEM lane quadrant this warp may access
        const int eset = (SETS == 1) ? 0 : (warp - 5) >> 2;
        const int row = quad * 32 + lane;                             // accumulator row = pixel of the tile
        const int et = threadIdx.x - 160;                             // 0..255
        const bool relu = a.relu_floor == 0.f;
        int acc = 0; uint32_t acc_phase = 0;
        uint32_t n_store = 0;
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
            const int img = fast_div(tile, a.div_tpi);
            const int rem = tile - img * a.tiles_x * a.tiles_y;
            const int ty = fast_div(rem, a.div_tx), tx = rem - ty * a.tiles_x;
            mbar_wait(t_full(acc), acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * 64 + eset * CW);
            const uint32_t o_stage = out_base + (n_store & 1u) * C1_OUT_BYTES;
            uint32_t r[CW];
            {
                uint32_t (&r0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&r[0]);
                tmem_ld32(t_row, r0);
                if (CW == 64) { uint32_t (&r1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&r[CW - 32]); tmem_ld32(t_row + 32u, r1); }
            }
            if (et == 0) bulk_wait_read<1>();                         // the store that last used this staging tile has read it
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(t_empty(acc));                 // this warp's share of the accumulator is in registers
            named_bar_sync(1, C1_EPI_THREADS);                        // staging tile free (et 0 has waited for the older store)
            const uint32_t rbase = o_stage + (uint32_t)row * 128u;
            const float4* sc4 = reinterpret_cast<const float4*>(s_scale + eset * CW);
            const float4* sh4 = reinterpret_cast<const float4*>(s_shift + eset * CW);
#pragma unroll
            for (int i = 0; i < CW / 8; ++i) {
                uint32_t wv[4];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const float4 sc = sc4[2 * i + j], sh = sh4[2 * i + j];
                    const float v0 = fmaf(__uint_as_float(r[8 * i + 4 * j]), sc.x, sh.x), v1 = fmaf(__uint_as_float(r[8 * i + 4 * j + 1]), sc.y, sh.y);
                    const float v2 = fmaf(__uint_as_float(r[8 * i + 4 * j + 2]), sc.z, sh.z), v3 = fmaf(__uint_as_float(r[8 * i + 4 * j + 3]), sc.w, sh.w);
                    wv[2 * j] = relu ? pack_relu_bf16x2(v0, v1) : pack_bf16x2(v0, v1);
                    wv[2 * j + 1] = relu ? pack_relu_bf16x2(v2, v3) : pack_bf16x2(v2, v3);
                }
                st_shared_v4(rbase + ((((uint32_t)(eset * (CW / 8) + i)) ^ ((uint32_t)row & 7u)) << 4), wv[0], wv[1], wv[2], wv[3]);
            }
            fence_proxy_async();
            named_bar_sync(1, C1_EPI_THREADS);
            if (et == 0) {
                tma_store_4d(&tmOut, o_stage, 0, tx * C1_TW, ty * C1_TH, img);
                bulk_commit();
            }
            ++n_store;
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
        if (et == 0) bulk_wait<0>();                                  // smem must outlive the last bulk store
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) tmem_dealloc(tmem_base, 128);
}

// x (n,1,h,w) fp32 -> out (n,h,w,64) bf16 = max(conv3x3(x, weight) * scale + shift, relu_floor).  Returns ADN_ERR_ARG for shapes /
// pointers the TMA store cannot take (the caller then uses the CUDA-core kernel).
int conv3x3_c1_tc(const float* x, int n, int h, int w, const float* weight, const float* scale, const float* shift, float relu_floor,
                  void* out, cudaStream_t stream) {
    if (!aligned16(out) || n <= 0 || h <= 0 || w <= 0) return ADN_ERR_ARG;
    C1Args a{};
    a.x = x; a.weight = weight; a.scale = scale; a.shift = shift; a.relu_floor = relu_floor;
    a.n_img = n; a.H = h; a.W = w;
    a.tiles_x = (w + C1_TW - 1) / C1_TW; a.tiles_y = (h + C1_TH - 1) / C1_TH;
    const long long tiles = (long long)n * a.tiles_x * a.tiles_y;
    if (tiles > 0x7fffffffLL) return ADN_ERR_ARG;
    a.num_tiles = (int)tiles;
    a.div_tpi = make_fastdiv(a.tiles_x * a.tiles_y); a.div_tx = make_fastdiv(a.tiles_x);
    CUtensorMap tmOut;
    int st = make_act_map(&tmOut, out, n, h, w, 64, C1_TW, C1_TH);
    if (st != ADN_OK) return st;
    static const int sets = (getenv("ADN_C1_SETS") && getenv("ADN_C1_SETS")[0] == '2') ? 2 : 1;     // A/B knob: epilogue warp sets
    const long long cap = (long long)num_sms() * 2;                   // persistent: two CTAs per SM
    const int grid = (int)(tiles < cap ? tiles : cap);
    if (sets == 2) {
        static unsigned char smem_set[64] = {0};
        ADN_CUDA_TRY(ensure_dyn_smem(conv3x3_c1_tc_kernel<2>, C1_SMEM_BYTES, smem_set));
        conv3x3_c1_tc_kernel<2><<<grid, c1_threads(2), C1_SMEM_BYTES, stream>>>(tmOut, a);
    } else {
        static unsigned char smem_set[64] = {0};
        ADN_CUDA_TRY(ensure_dyn_smem(conv3x3_c1_tc_kernel<1>, C1_SMEM_BYTES, smem_set));
        conv3x3_c1_tc_kernel<1><<<grid, c1_threads(1), C1_SMEM_BYTES, stream>>>(tmOut, a);
    }
    ADN_LAUNCH_CHECK();
    return ADN_OK;
}

}  // namespace adn
