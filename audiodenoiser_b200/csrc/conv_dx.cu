// conv_dx.cu -- the 64-OUTPUT-CHANNEL 3x3 convs of the reference UNet (code/model.py:11-16 at full resolution: downconv1's
// second conv, upconv4's two convs, the conv in front of the 1x1 head, model.py:68,93) as a tcgen05 implicit GEMM whose N
// dimension carries the three HORIZONTAL taps:
//
//   D[(row r, halo column xh), (kx, co)] += A_ky[(r, xh), ci] * W[ky][(kx, co), ci]^T        ky = 0..2, ci in 64-channel chunks
//   out[r, x, co] = D[(r, x), kx=0] + D[(r, x+1), kx=1] + D[(r, x+2), kx=2]                   (halo column xh = x + 1 + (kx-1))
//
// Why: with N = 64 a 128x64x16 UMMA reads 4 KB of A + 2 KB of B from shared memory for 32 clk of tensor time, and the
// shared-memory operand path delivers ~96 B/clk (measured: those layers ran at 72 clk per UMMA, 0.5 of the tensor peak,
// profiles/README.md).  Folding kx into N makes it 128x192x16: 4 KB + 6 KB for 98 clk of tensor time -- the A bytes per MAC drop
// 3x and the three vertical taps are three whole-row offsets into one halo tile (1024-byte aligned descriptors).  The price: the
// tile is 8 rows x 16 halo columns of which 14 are outputs (87.5 % useful rows), and the epilogue adds three accumulator
// columns across neighbouring lanes (two shuffles per output value) -- the lanes of one image row sit in one half warp.
//
//   A operand   4-D TMA box {64 ch, 16 px, 10 rows, 1 image} per 64-channel chunk at (x0 - 1, y0 - 1): zero OOB fill = conv
//               padding and the F.pad of the up-sampled half of a concat (model.py:44-49).  UMMA A for ky = plain K-major
//               SWIZZLE_128B tile of 128 rows at base + ky * 2048 B.
//   B operand   the packed weights [co][tap][ci] (adn_pack_conv3x3_weight_bf16) loaded ONCE per CTA as 9 * chunks boxes {64, 64}
//               into [chunk][ky][kx][co] order: one (chunk, ky) operand = 192 consecutive rows.
//   roles       warp 0 A producer, warp 1 UMMA issuer, warp 2 weight loader + TMEM allocator, then SETS x 4 epilogue warps: set s
//               owns output channels [64 s / SETS, 64 (s + 1) / SETS) of all four TMEM lane quadrants (SETS = 2, or 4 for the
//               pool-fused layer whose epilogue is the bottleneck)
//               (tcgen05.ld -> neighbour-lane sum -> fp32 scale/shift -> ReLU -> bf16 -> the warp's own swizzled staging -> its own
//               TMA store of a {32 | 16 ch, 14 px, 2 rows} box, + fused 2x2 max-pool {., 7, 1} | fused 1x1 head): no CTA-wide
//               barrier anywhere in the tile loop.  Two TMEM accumulators, each released as soon as a warp holds its columns
//               in registers.  With K = 576 these layers are epilogue-bound: one warp per scheduler could not hide the shuffle /
//               TMEM latencies (measured 1.98 ms for downconv1's second conv with 4 epilogue warps, 1.60 with 8, 1.32 with 16).
#include "tc_common.cuh"

namespace adn {

constexpr int X_TW = 14, X_TH = 8;                  // output pixels of a tile
constexpr int X_PITCH = 16, X_ROWS = X_TH + 2;      // halo tile: 10 rows x 16 pixels
constexpr int X_A_STAGE = X_ROWS * X_PITCH * 128;   // 20 480 B
constexpr int X_B_TAP = 64 * 128;                   // one (tap, chunk) weight box
constexpr int X_B_BLOCK = 3 * X_B_TAP;              // one (chunk, ky) operand: 192 rows
// per epilogue warp (SETS sets of 4 warps, each warp owning 64 / SETS channels of its 2 tile rows): staging of 2 rows x 14 px and
// of 7 pooled px, rows of 64 B (SWIZZLE_64B) or 32 B (SWIZZLE_32B), padded to the swizzle period
constexpr int x_out_stage(int sets) { return 2048 * 2 / sets; }
constexpr int x_pool_stage(int sets) { return 512 * 2 / sets; }
constexpr int x_threads(int sets) { return 96 + 128 * sets; }     // 3 role warps + SETS x 4 epilogue warps
constexpr int X_MAX_A = 6;
constexpr int X_AUX_F32 = 3 * 64 + 2 * 3 * 128;    // scale, shift, head weights, head partials [2][SETS - 1][128]
constexpr int X_ACC_COLS = 256;                     // TMEM column pitch of the two accumulators (192 used)

struct DxArgs {
    int c0_chunks, c1_chunks;
    int n_img, H, W;
    int tiles_x, tiles_y, num_tiles;
    FastDiv div_tpi, div_tx;
    int head;                      // 1: fused 1x1 head (fp32 out), 0: NHWC bf16 out (+ optional pool)
    int pool;
    int a_stages;
    int stage_bufs;                // 1 or 2 staging buffers per epilogue warp (2 when the weights leave room: one input chunk)
    float relu_floor;
    const float* scale;
    const float* shift;
    const float* head_w;
    const float* head_b;
    float* head_out;
};

// NCTA == 2: a CTA pair (cluster of 2, cta_group::2) computes TWO tiles with ONE 256x192x16 UMMA per k-step; each CTA holds its own
// halo tile and the weight rows of HALF the output channels (columns [96 r, 96 r + 96) = [kx][32 channels] of CTA r), so a CTA
// reads 4 KB of A + 3 KB of B per UMMA instead of 4 + 6: the single-CTA kernel runs at the shared-memory operand bandwidth
// (tensor pipe 75 % active).  Barrier protocol as in conv_halo.cu: full / accumulator-empty barriers live in the leader, a peer's
// TMA completes its bytes there, commits are multicast to both CTAs.
template <int NCTA, int SETS>
__global__ void __launch_bounds__(x_threads(SETS), 1)
conv3x3_dx_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                  const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOut,
                  const __grid_constant__ CUtensorMap tmPool, const DxArgs a) {
    constexpr bool PAIR = (NCTA == 2);
    constexpr int X_EPI_THREADS = 128 * SETS, EW = 4 * SETS;       // epilogue threads / warps
    constexpr int X_OUT_STAGE = x_out_stage(SETS), X_POOL_STAGE = x_pool_stage(SETS);
    constexpr int CW = 64 / SETS, GW = CW / 16;                    // channels / 16-channel groups per epilogue warp
    constexpr int B_TAP = X_B_TAP / NCTA;                            // this CTA's rows of one (tap, chunk) weight box
    constexpr int B_BLOCK = 3 * B_TAP;
    const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
    const bool leader = (cta_rank == 0);
    extern __shared__ uint8_t smem_dyn[];
    const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_dyn + (smem_base - smem_u32(smem_dyn));
    const int chunks = a.c0_chunks + a.c1_chunks;

    const uint32_t a_base = smem_base;
    const uint32_t b_base = a_base + (uint32_t)a.a_stages * X_A_STAGE;
    const uint32_t stage_off = (uint32_t)a.a_stages * X_A_STAGE + (uint32_t)(chunks * 3) * B_BLOCK;
    const uint32_t stage_bytes = a.head ? 0u : (uint32_t)a.stage_bufs * (uint32_t)(EW * X_OUT_STAGE + (a.pool ? EW * X_POOL_STAGE : 0));
    const uint32_t aux_off = stage_off + stage_bytes;
    float* s_scale = reinterpret_cast<float*>(smem_gen + aux_off);
    float* s_shift = s_scale + 64;
    float* s_head = s_shift + 64;
    float* s_hpart = s_head + 64;                                     // [2][SETS - 1][128] head partial sums of sets 1..
    const uint32_t bar_base = smem_base + aux_off + X_AUX_F32 * 4;
    auto full_a = [&](int s) { return bar_base + 8u * s; };
    auto empty_a = [&](int s) { return bar_base + 8u * (X_MAX_A + s); };
    auto tfull = [&](int s) { return bar_base + 8u * (2 * X_MAX_A + s); };
    auto tempty = [&](int s) { return bar_base + 8u * (2 * X_MAX_A + 2 + s); };
    const uint32_t bres = bar_base + 8u * (2 * X_MAX_A + 4);
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem_gen + aux_off + X_AUX_F32 * 4 + (2 * X_MAX_A + 5) * 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA0); tma_prefetch_desc(&tmA1); tma_prefetch_desc(&tmB);
        if (!a.head) { tma_prefetch_desc(&tmOut); if (a.pool) tma_prefetch_desc(&tmPool); }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < X_MAX_A; ++s) { mbar_init(full_a(s), 1); mbar_init(empty_a(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull(s), 1); mbar_init(tempty(s), NCTA * X_EPI_THREADS / 32); }
        mbar_init(bres, 1);
        fence_barrier_init();
    }
    if (warp == 2) {
        if (PAIR) { tmem_alloc_pair(smem_u32(tmem_ptr_smem), 512); tmem_relinquish_pair(); }
        else { tmem_alloc(smem_u32(tmem_ptr_smem), 512); tmem_relinquish(); }
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();           // peers must see initialised barriers before any remote arrive
    tc_fence_after();
    const uint32_t tmem_own = *tmem_ptr_smem;
    const uint32_t tmem_base = PAIR ? ld_shared_cluster_u32(mapa_shared(smem_u32(tmem_ptr_smem), 0)) : tmem_own;
    const int tiles_per_img = a.tiles_x * a.tiles_y;
    // work items: single CTA -> tile; pair -> two consecutive tiles, this CTA taking 2 w + rank (the odd tail is an all-zero
    // tile of image n_img whose stores are clipped away)
    const int work_total = PAIR ? (a.num_tiles + 1) / 2 : a.num_tiles;
    const int work_first = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int work_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    auto tile_of = [&](int w) { return PAIR ? 2 * w + (int)cta_rank : w; };

    if (warp == 0) {
        // ===================================================================== A producer: one halo tile per (tile, chunk)
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int w = work_first; w < work_total; w += work_step) {
                const int t = tile_of(w);
                const int img = fast_div(t, a.div_tpi);
                const int rem = t - img * tiles_per_img;
                const int ty = fast_div(rem, a.div_tx), tx = rem - ty * a.tiles_x;
                for (int ch = 0; ch < chunks; ++ch) {
                    mbar_wait(empty_a(stage), phase ^ 1u);
                    if (leader) mbar_arrive_expect_tx(full_a(stage), NCTA * X_A_STAGE);
                    const CUtensorMap* map = (ch < a.c0_chunks) ? &tmA0 : &tmA1;
                    const int c = (ch < a.c0_chunks ? ch : ch - a.c0_chunks) * 64;
                    const uint32_t dst = a_base + (uint32_t)stage * X_A_STAGE;
                    if (PAIR) tma_load_4d_pair(dst, map, mapa_shared(full_a(stage), 0), c, tx * X_TW - 1, ty * X_TH - 1, img);
                    else tma_load_4d(dst, map, full_a(stage), c, tx * X_TW - 1, ty * X_TH - 1, img);
                    if (++stage == a.a_stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 2) {
        // ===================================================================== weights: resident for the CTA's lifetime
        if (lane == 0) {
            if (leader) mbar_arrive_expect_tx(bres, (uint32_t)(NCTA * 9 * chunks) * B_TAP);
            const uint32_t sig = PAIR ? mapa_shared(bres, 0) : bres;
            const int row0 = (int)cta_rank * (64 / NCTA);            // this CTA's output channels
            for (int ch = 0; ch < chunks; ++ch)
                for (int tap = 0; tap < 9; ++tap) {
                    const uint32_t dst = b_base + (uint32_t)(ch * 9 + tap) * B_TAP;
                    if (PAIR) tma_load_2d_pair(dst, &tmB, sig, (tap * chunks + ch) * 64, row0);
                    else tma_load_2d(dst, &tmB, sig, (tap * chunks + ch) * 64, row0);
                }
        }
    } else if (warp == 1) {
      // ===================================================================== UMMA issuer (warp-uniform flow, one elected lane;
      // in a pair only the leader's warp issues, for both CTAs)
      if (leader) {
        constexpr uint32_t idesc = make_idesc(192, 128 * NCTA);
        auto umma = [&](uint32_t d, uint64_t da, uint64_t db, uint32_t accf) {
            if (PAIR) umma_bf16_pair(d, da, db, idesc, accf); else umma_bf16(d, da, db, idesc, accf);
        };
        auto commit = [&](uint32_t bar) { if (PAIR) umma_commit_pair(bar); else umma_commit(bar); };
        int sa = 0; uint32_t pa = 0;
        int acc = 0; uint32_t acc_phase = 0;
        mbar_wait(bres, 0);
        tc_fence_after();
        const uint64_t db_base = make_sw128_desc(b_base);
        for (int w = work_first; w < work_total; w += work_step) {
            mbar_wait(tempty(acc), acc_phase ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * X_ACC_COLS);
            for (int ch = 0; ch < chunks; ++ch) {
                mbar_wait(full_a(sa), pa);
                tc_fence_after();
                const uint64_t da = make_sw128_desc(a_base + (uint32_t)sa * X_A_STAGE);
                const uint64_t db = db_base + (uint64_t)((uint32_t)(ch * 3) * (B_BLOCK >> 4));
                if (elect_one()) {
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma(d_tmem, da + (uint64_t)((ky * X_PITCH * 128 + k * 32) >> 4), db + (uint64_t)((ky * B_BLOCK + k * 32) >> 4),
                                 (ky | k) != 0 ? 1u : (ch != 0 ? 1u : 0u));
                    commit(empty_a(sa));
                }
                __syncwarp();
                if (++sa == a.a_stages) { sa = 0; pa ^= 1u; }
            }
            if (elect_one()) commit(tfull(acc));
            __syncwarp();
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
      }
    } else {
        // ===================================================================== epilogue (warps 3..10; TMEM lane quadrant = warp & 3)
        const int quad = warp & 3;
        const int et = threadIdx.x - 96;
        const int set = (warp - 3) >> 2;                             // channel slice [CW set, CW set + CW) owned by this warp
        const int r = quad * 2 + (lane >> 4), xh = lane & 15;         // tile row, halo column of this lane's accumulator row
        const bool out_lane = xh >= 1 && xh <= X_TW;
        const int xo = xh - 1;
        if (et < 64) { s_scale[et] = a.scale[et]; s_shift[et] = a.shift[et]; if (a.head) s_head[et] = a.head_w[et]; }
        named_bar_sync(1, X_EPI_THREADS);
        // 2x2 pool partners: (xo even, xo + 1) along x = lanes (xh odd, xh + 1); (r even, r + 1) along y = lanes l, l ^ 16
        const bool x_first = xh & 1, y_first = !(lane & 16);
        const int x_partner = x_first ? ((lane + 1) & 31) : ((lane + 31) & 31);
        // every epilogue warp stages and stores its own {32 ch, 14 px, 2 rows} box (+ {32, 7, 1} pooled): no CTA-wide barrier
        const int ew = warp - 3;
        const uint32_t srow = (uint32_t)((lane >> 4) * X_TW + xo);    // staging row (64 B) of this lane's output pixel
        const uint32_t prow = (uint32_t)(xo >> 1);
        const uint32_t o_stage0 = smem_base + stage_off + (uint32_t)(ew * a.stage_bufs) * X_OUT_STAGE;
        const uint32_t p_stage0 = smem_base + stage_off + (uint32_t)(EW * a.stage_bufs) * X_OUT_STAGE + (uint32_t)(ew * a.stage_bufs) * X_POOL_STAGE;
        int acc = 0; uint32_t acc_phase = 0;
        uint32_t store_groups = 0;
        const uint32_t tempty_sig0 = PAIR ? mapa_shared(tempty(0), 0) : tempty(0);
        const uint32_t tempty_sig1 = PAIR ? mapa_shared(tempty(1), 0) : tempty(1);
        // accumulator columns of (kx, 16-channel group gl of this warp's channel half)
        const int ch0 = set * CW;
        auto acc_col = [&](int kx, int gl) {
            const int c = ch0 + gl * 16;
            return (uint32_t)(PAIR ? (c >> 5) * 96 + kx * 32 + (c & 31) : kx * 64 + c);
        };
        for (int w = work_first; w < work_total; w += work_step) {
            const int t = tile_of(w);
            const int img = fast_div(t, a.div_tpi);                 // == n_img for the padding tile of an odd pair
            const int rem = t - img * tiles_per_img;
            const int ty = fast_div(rem, a.div_tx), tx = rem - ty * a.tiles_x;
            const int x = tx * X_TW + xo, y = ty * X_TH + r;
            const uint32_t buf = (a.stage_bufs == 2) ? (store_groups & 1u) : 0u;
            const uint32_t o_stage = o_stage0 + buf * X_OUT_STAGE, p_stage = p_stage0 + buf * X_POOL_STAGE;
            mbar_wait(tfull(acc), acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * X_ACC_COLS);
            float2 head_acc2 = make_float2(0.f, 0.f);

            // 16 output channels: sum the three kx column blocks across neighbouring lanes, affine (+ ReLU), pack, stage
            auto group = [&](const int g, const uint32_t (&k0)[16], const uint32_t (&k1)[16], const uint32_t (&k2)[16]) {
                float v[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float left = __shfl_up_sync(0xffffffffu, __uint_as_float(k0[i]), 1);
                    const float right = __shfl_down_sync(0xffffffffu, __uint_as_float(k2[i]), 1);
                    v[i] = (left + __uint_as_float(k1[i])) + right;
                }
                const float4* sc4 = reinterpret_cast<const float4*>(s_scale + g * 16);
                const float4* sh4 = reinterpret_cast<const float4*>(s_shift + g * 16);
                float2 y[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 sc = sc4[i], sh = sh4[i];
                    y[2 * i] = pk_fma(make_float2(v[4 * i], v[4 * i + 1]), make_float2(sc.x, sc.y), make_float2(sh.x, sh.y));
                    y[2 * i + 1] = pk_fma(make_float2(v[4 * i + 2], v[4 * i + 3]), make_float2(sc.z, sc.w), make_float2(sh.z, sh.w));
                }
                if (a.head) {                                         // warp-uniform
                    const float4* hw4 = reinterpret_cast<const float4*>(s_head + g * 16);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 hw = hw4[i];
                        float2 y0 = y[2 * i], y1 = y[2 * i + 1];
                        y0.x = fmaxf(y0.x, a.relu_floor); y0.y = fmaxf(y0.y, a.relu_floor);
                        y1.x = fmaxf(y1.x, a.relu_floor); y1.y = fmaxf(y1.y, a.relu_floor);
                        head_acc2 = pk_fma(y0, make_float2(hw.x, hw.y), head_acc2);
                        head_acc2 = pk_fma(y1, make_float2(hw.z, hw.w), head_acc2);
                    }
                    return;
                }
                uint32_t pk[8];
                if (a.relu_floor == 0.f) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) pk[i] = pack_relu_bf16x2(y[i].x, y[i].y);
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) pk[i] = pack_bf16x2(y[i].x, y[i].y);
                }
                const uint32_t c0 = (uint32_t)(g % GW) * 2u;             // 16-byte chunk of this group within the warp's channels
                if (out_lane) {
                    const uint32_t rbase = o_stage + srow * (uint32_t)(CW * 2), sw = (GW == 2) ? ((srow >> 1) & 3u) : ((srow >> 2) & 1u);
                    st_shared_v4(rbase + ((c0 ^ sw) << 4), pk[0], pk[1], pk[2], pk[3]);
                    st_shared_v4(rbase + (((c0 + 1u) ^ sw) << 4), pk[4], pk[5], pk[6], pk[7]);
                }
                if (a.pool) {                                         // warp-uniform
                    // halving exchanges: along x the first lane keeps registers 0..3, along y the first keeps 0..1; every lane
                    // ends with ONE 8-byte piece (4 channels) of its 2x2 window's pooled pixel
                    uint32_t m1[4], m2[2];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint32_t keep = x_first ? pk[i] : pk[i + 4], send = x_first ? pk[i + 4] : pk[i];
                        m1[i] = bf16x2_max(keep, __shfl_sync(0xffffffffu, send, x_partner));
                    }
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const uint32_t keep = y_first ? m1[i] : m1[i + 2], send = y_first ? m1[i + 2] : m1[i];
                        m2[i] = bf16x2_max(keep, __shfl_xor_sync(0xffffffffu, send, 16));
                    }
                    if (out_lane) {
                        const uint32_t c16 = c0 + (x_first ? 0u : 1u);
                        const uint32_t psw = (GW == 2) ? ((prow >> 1) & 3u) : ((prow >> 2) & 1u);
                        st_shared_v2(p_stage + prow * (uint32_t)(CW * 2) + ((c16 ^ psw) << 4) + (y_first ? 0u : 8u), m2[0], m2[1]);
                    }
                }
            };

            // this warp's GW x 16 output channels x 3 kx blocks -> registers, then the accumulator is free for the MMA warp
            uint32_t ra0[16], ra1[16], ra2[16], rb0[16], rb1[16], rb2[16];
            tmem_ld16(t_row + acc_col(0, 0), ra0); tmem_ld16(t_row + acc_col(1, 0), ra1); tmem_ld16(t_row + acc_col(2, 0), ra2);
            if (GW == 2) { tmem_ld16(t_row + acc_col(0, 1), rb0); tmem_ld16(t_row + acc_col(1, 1), rb1); tmem_ld16(t_row + acc_col(2, 1), rb2); }
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {                                         // one arrival per epilogue warp (of both CTAs) frees the accumulator
                if (PAIR) mbar_arrive_cluster(acc ? tempty_sig1 : tempty_sig0); else mbar_arrive(tempty(acc));
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            if (!a.head) {                                           // the TMA store that last used this staging buffer must have read it
                if (elect_one()) { if (a.stage_bufs == 2) bulk_wait_read<1>(); else bulk_wait_read<0>(); }
                __syncwarp();
            }
            group(set * GW, ra0, ra1, ra2);
            if (GW == 2) group(set * GW + 1, rb0, rb1, rb2);

            if (a.head) {                                            // sets 1.. hand their part of the dot product to set 0
                float* part = s_hpart + (store_groups & 1u) * (3 * 128) + quad * 32 + lane;
                if (set > 0) part[(set - 1) * 128] = head_acc2.x + head_acc2.y;
                named_bar_sync(2 + quad, 32 * SETS);                  // the warps of this lane quadrant
                if (set == 0 && out_lane && x < a.W && y < a.H && t < a.num_tiles) {
                    float sum = head_acc2.x + head_acc2.y;
#pragma unroll
                    for (int j = 0; j < SETS - 1; ++j) sum += part[j * 128];
                    a.head_out[((long long)img * a.H + y) * a.W + x] = sum + a.head_b[0];
                }
                ++store_groups;
            } else {
                fence_proxy_async();
                __syncwarp();
                if (elect_one()) {                                   // the same lane every time (bulk groups are per thread); uniform
                    tma_store_4d(&tmOut, o_stage, ch0, tx * X_TW, ty * X_TH + quad * 2, img);
                    if (a.pool) tma_store_4d(&tmPool, p_stage, ch0, tx * (X_TW / 2), ty * (X_TH / 2) + quad, img);
                    bulk_commit();
                }
                ++store_groups;
            }
        }
        __syncwarp();
        if (!a.head && elect_one()) bulk_wait<0>();                      // smem must outlive the last bulk stores
    }

    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();           // the peer may still read this CTA's smem / write its TMEM
    tc_fence_after();
    if (warp == 2) { if (PAIR) tmem_dealloc_pair(tmem_own, 512); else tmem_dealloc(tmem_own, 512); }
}

// ------------------------------------------------------------------------------------------------ host side
// (n,h,w,64) NHWC bf16 output written in per-warp boxes {cw ch, tw, th, 1}: 64-byte rows + SWIZZLE_64B or 32-byte rows + SWIZZLE_32B
static int make_store_map(CUtensorMap* map, const void* ptr, int n, int h, int w, int tw, int th, int cw) {
    PFN_tmapEncodeTiled enc = get_encode_fn();
    if (!enc) return ADN_ERR_DRIVER;
    cuuint64_t dims[4] = {64u, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
    cuuint64_t strides[3] = {128u, (cuuint64_t)w * 128u, (cuuint64_t)h * w * 128u};
    cuuint32_t box[4] = {(cuuint32_t)cw, (cuuint32_t)tw, (cuuint32_t)th, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, cw == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? ADN_OK : ADN_ERR_DRIVER;
}

// CTA pairs are OFF by default: measured on B200 (scripts/ab_conv_modes.py, batch 64) the 256x192x16 pair UMMA (96 weight rows per
// CTA) runs 2.7x slower per instruction than the single-CTA 128x192x16 one -- upconv4.0 2.03 -> 2.59 ms, the K = 576 layers 2x
// slower -- the same effect conv_halo.cu saw with 32 rows per CTA; only 64-row halves (N = 128 pairs) are fast.  Kept behind the
// debug hook adn__conv_dx_mode(2) because the parity tests cover it.
int g_dx_pair = 0;
// epilogue warp sets of the single-CTA kernel: 2 (32 channels per warp) or 4 (16 per warp); 0 = automatic: four sets where the
// fused pool makes the epilogue the bottleneck (measured, batch 64: downconv1.3 1.60 -> 1.32 ms), two elsewhere (upconv4.0 2.09 vs
// 2.26 ms, head conv 1.19 vs 1.28 ms with four: the extra warps take issue slots from the MMA / producer warps)
int g_dx_sets = 0;

bool conv3x3_dx_eligible(int c0, int c1, int c_out) { return c_out == 64 && (c0 + c1) / 64 <= 2; }

int conv3x3_dx(const void* src0, int c0, const void* src1, int c1, int h1, int w1, int n, int h, int w, const void* w_packed,
               const float* scale, const float* shift, float relu_floor, void* out, void* pool_out, const float* head_w,
               const float* head_b, float* head_out, cudaStream_t stream) {
    DxArgs args{};
    args.c0_chunks = c0 / 64; args.c1_chunks = c1 / 64;
    const int chunks = args.c0_chunks + args.c1_chunks;
    args.n_img = n; args.H = h; args.W = w;
    args.tiles_x = (w + X_TW - 1) / X_TW; args.tiles_y = (h + X_TH - 1) / X_TH;
    const long long tiles = (long long)n * args.tiles_x * args.tiles_y;
    if (tiles > 0x7fffffffLL) return ADN_ERR_ARG;
    args.num_tiles = (int)tiles;
    args.div_tpi = make_fastdiv(args.tiles_x * args.tiles_y); args.div_tx = make_fastdiv(args.tiles_x);
    args.head = head_out ? 1 : 0;
    args.pool = (!args.head && pool_out) ? 1 : 0;
    args.relu_floor = relu_floor;
    args.scale = scale; args.shift = shift;
    args.head_w = head_w; args.head_b = head_b; args.head_out = head_out;

    CUtensorMap mA0, mA1, mB, mOut, mPool;
    int st = make_act_map(&mA0, src0, n, h, w, c0, X_PITCH, X_ROWS);
    if (st != ADN_OK) return st;
    if (c1 > 0) st = make_act_map(&mA1, src1, n, h1, w1, c1, X_PITCH, X_ROWS); else mA1 = mA0;
    if (st != ADN_OK) return st;
    const int sms = num_sms();
    const bool pair = g_dx_pair && args.num_tiles >= 2 * sms;
    const int ncta = pair ? 2 : 1;
    const int sets = pair ? 2 : (g_dx_sets ? g_dx_sets : (args.pool ? 4 : 2));
    mOut = mA0; mPool = mA0;
    if (!args.head) {
        st = make_store_map(&mOut, out, n, h, w, X_TW, 2, 64 / sets);
        if (st != ADN_OK) return st;
        if (args.pool) st = make_store_map(&mPool, pool_out, n, h / 2, w / 2, X_TW / 2, 1, 64 / sets);
        if (st != ADN_OK) return st;
    }

    constexpr int MAX_DYN = 232448;
    const int AUX = X_AUX_F32 * 4 + (2 * X_MAX_A + 5) * 8 + 16;
    args.stage_bufs = (chunks == 1) ? 2 : 1;
    const int fixed = 1024 + chunks * 3 * (X_B_BLOCK / ncta) +
                      (args.head ? 0 : args.stage_bufs * 4 * sets * (x_out_stage(sets) + (args.pool ? x_pool_stage(sets) : 0))) + AUX;
    int stages = (MAX_DYN - fixed) / X_A_STAGE;
    if (stages < 2) return ADN_ERR_ARG;
    args.a_stages = stages > X_MAX_A ? X_MAX_A : stages;
    const int smem = fixed + args.a_stages * X_A_STAGE;
    st = make_weight_map(&mB, w_packed, 64, 9 * (c0 + c1), 64 / ncta);
    if (st != ADN_OK) return st;
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute attr[1];
    cfg.blockDim = dim3(x_threads(sets)); cfg.dynamicSmemBytes = (size_t)smem; cfg.stream = stream;
    if (pair) {
        static unsigned char smem_set[64] = {0};
        ADN_CUDA_TRY(ensure_dyn_smem(conv3x3_dx_kernel<2, 2>, MAX_DYN, smem_set));
        const int pairs = (args.num_tiles + 1) / 2;
        cfg.gridDim = dim3(2 * (pairs < sms / 2 ? pairs : sms / 2));
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        ADN_CUDA_TRY(cudaLaunchKernelEx(&cfg, conv3x3_dx_kernel<2, 2>, mA0, mA1, mB, mOut, mPool, args));
    } else {
        cfg.gridDim = dim3(args.num_tiles < sms ? args.num_tiles : sms);
        cfg.attrs = nullptr; cfg.numAttrs = 0;
        if (sets == 4) {
            static unsigned char smem_set[64] = {0};
            ADN_CUDA_TRY(ensure_dyn_smem(conv3x3_dx_kernel<1, 4>, MAX_DYN, smem_set));
            ADN_CUDA_TRY(cudaLaunchKernelEx(&cfg, conv3x3_dx_kernel<1, 4>, mA0, mA1, mB, mOut, mPool, args));
        } else {
            static unsigned char smem_set[64] = {0};
            ADN_CUDA_TRY(ensure_dyn_smem(conv3x3_dx_kernel<1, 2>, MAX_DYN, smem_set));
            ADN_CUDA_TRY(cudaLaunchKernelEx(&cfg, conv3x3_dx_kernel<1, 2>, mA0, mA1, mB, mOut, mPool, args));
        }
    }
    ADN_LAUNCH_CHECK();
    return ADN_OK;
}

}  // namespace adn
