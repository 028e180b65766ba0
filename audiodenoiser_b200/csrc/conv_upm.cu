// conv_upm.cu -- the first conv of a decoder level with the ConvTranspose2d in front of it MERGED INTO ITS WEIGHTS
// (reference: UpSampleLayer.forward, code/model.py:41-50: x1 = ConvTranspose2d(k=2,s=2)(x1); F.pad; cat([x2, x1]); Conv3x3+BN+ReLU).
//
// ConvTranspose2d(k=2,s=2) has no overlap: up(Y,X) = Wt[Y&1][X&1] * low(Y>>1, X>>1) + bt.  A 3x3 conv over `up` is therefore, for
// the output pixels of ONE parity class (py,px) = (Y&1, X&1), a 2x2 conv over `low` with class-specific weights
//     Weff[py,px][dy,dx] = sum over the taps (ky,kx) that land on low row/col (dy,dx) of  W3_up[ky,kx] * Wt[(py+ky-1)&1][(px+kx-1)&1]
// (4 x Cl MACs per output pixel and channel instead of 9 x Cl/2 + the ConvTranspose itself), and the up-sampled tensor is never
// written or read: per decoder level one kernel launch and 2-6 GB of HBM traffic disappear.  The ConvTranspose bias becomes a
// per-channel constant (folded into the BatchNorm shift on the host) minus a correction on the border pixels whose 3x3 window
// hangs over the edge of `up` (conv zero padding / the F.pad of model.py:44-47); the linear part needs no special case because
// out-of-range rows / columns of `low` arrive as zeros from TMA exactly where `up` would have been zero.
//
//   GEMM view   one M tile = 128 output pixels OF ONE PARITY CLASS: 16 x 8 pixels of the half-resolution plane (Y = 2y+py, X = 2x+px).
//               K = 9 x C0 (skip tensor, all classes share these weights) + 4 x Cl (low tensor, class-specific weights).
//   A operand   skip: the four parity planes of the NHWC skip tensor are plain strided TMA views (pixel stride doubled, base shifted);
//               per 64-channel chunk one dense {64 ch, 9 px, 17 rows} box per plane, the 4 / 2 / 2 / 1 taps that read that plane are UMMA
//               descriptors into it (SBO = 1152 B: the swizzle follows absolute address bits).  low: one {64, 9, 17} box per chunk, 4 taps.
//   B operand   packed [Cout][9*C0 + 4 classes x 4 taps x Cl] bf16, streamed block by block in consumption order.
//   output      TMA stores through four class-specific strided views of the NHWC output.
//   roles       as conv_halo.cu (A producer, MMA issuer, B producer + TMEM allocator, 4 epilogue warps), always as a CTA PAIR
//               (cta_group::2: one 256-row UMMA over two tiles of the same class, each CTA holding half of the weight rows).
#include "tc_common.cuh"

namespace adn {

constexpr int U_TW = 8, U_TH = 16;
constexpr int U_BW = U_TW + 1, U_BH = U_TH + 1;
constexpr int U_A_BYTES = U_BH * U_BW * 128;                   // 19 584 B
constexpr int U_A_STAGE = (U_A_BYTES + 1023) & ~1023;          // 20 480 B
constexpr int U_THREADS = 224, U_EPI_THREADS = 128;
constexpr int U_MAX_A = 8, U_MAX_B = 16;
constexpr int U_OUT_STAGE = 128 * 128;

struct UpmMaps { CUtensorMap skip[4]; CUtensorMap out[4]; };

struct UpmArgs {
    int c0_chunks, cl_chunks;
    int n_img, H, W;               // output (= skip) size
    int Hu, Wu;                    // extent of the up-sampled map: 2*hl, 2*wl (H - Hu, W - Wu in {0, 1})
    int tiles_y;                   // plane tile rows: ceil(ceil(H/2) / 16)
    int ntx;                       // decoded tile-column range (pair over images: plane tile columns; pair over columns: half of them)
    int pair_img;                  // 1: the CTAs of a pair take images 2a, 2a+1 of one tile; 0: tile columns 2q, 2q+1 of one image
    int c_out, n_blocks, work_total;
    FastDiv div_nb, div_ntx, div_ty;
    int a_stages, b_slots;
    const float* scale;
    const float* shift;            // BN shift with the interior ConvTranspose-bias term folded in
    const float* wb;               // [9][c_out]: scale[co] * sum_cu W3_up[co][cu][tap] * bt[cu]  (border correction)
};

template <int BLOCK_N>
__global__ void __launch_bounds__(U_THREADS, 1)
conv3x3_upm_kernel(const __grid_constant__ UpmMaps maps, const __grid_constant__ CUtensorMap tmLow,
                   const __grid_constant__ CUtensorMap tmB, const UpmArgs a) {
    constexpr int B_ROWS = BLOCK_N / 2;
    constexpr int B_BLOCK = B_ROWS * 128;
    constexpr int TMEM_COLS = 2 * BLOCK_N;
    const uint32_t cta_rank = cluster_ctarank();
    const bool leader = (cta_rank == 0);
    extern __shared__ uint8_t smem_dyn[];
    const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_dyn + (smem_base - smem_u32(smem_dyn));

    const uint32_t a_base = smem_base;
    const uint32_t b_base = a_base + (uint32_t)a.a_stages * U_A_STAGE;
    const uint32_t stage_off = (uint32_t)a.a_stages * U_A_STAGE + (uint32_t)a.b_slots * B_BLOCK;
    const uint32_t aux_off = stage_off + 2u * U_OUT_STAGE;
    float* s_scale = reinterpret_cast<float*>(smem_gen + aux_off);
    float* s_shift = s_scale + a.c_out;
    const uint32_t aux_f32 = (uint32_t)(2 * a.c_out) * 4;
    const uint32_t bar_base = smem_base + aux_off + aux_f32;
    auto full_a = [&](int s) { return bar_base + 8u * s; };
    auto empty_a = [&](int s) { return bar_base + 8u * (U_MAX_A + s); };
    auto full_b = [&](int s) { return bar_base + 8u * (2 * U_MAX_A + s); };
    auto empty_b = [&](int s) { return bar_base + 8u * (2 * U_MAX_A + U_MAX_B + s); };
    auto tfull = [&](int s) { return bar_base + 8u * (2 * U_MAX_A + 2 * U_MAX_B + s); };
    auto tempty = [&](int s) { return bar_base + 8u * (2 * U_MAX_A + 2 * U_MAX_B + 2 + s); };
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem_gen + aux_off + aux_f32 + (2 * U_MAX_A + 2 * U_MAX_B + 4) * 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        for (int p = 0; p < 4; ++p) { tma_prefetch_desc(&maps.skip[p]); tma_prefetch_desc(&maps.out[p]); }
        tma_prefetch_desc(&tmLow);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < U_MAX_A; ++s) { mbar_init(full_a(s), 1); mbar_init(empty_a(s), 1); }
        for (int s = 0; s < U_MAX_B; ++s) { mbar_init(full_b(s), 1); mbar_init(empty_b(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull(s), 1); mbar_init(tempty(s), 2 * U_EPI_THREADS / 32); }
        fence_barrier_init();
    }
    if (warp == 2) { tmem_alloc_pair(smem_u32(tmem_ptr_smem), TMEM_COLS); tmem_relinquish_pair(); }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_own = *tmem_ptr_smem;
    const uint32_t tmem_base = ld_shared_cluster_u32(mapa_shared(smem_u32(tmem_ptr_smem), 0));

    const int work_first = (int)(blockIdx.x >> 1), work_step = (int)(gridDim.x >> 1);
    const int Hc0 = (a.H + 1) >> 1, Hc1 = a.H >> 1, Wc0 = (a.W + 1) >> 1, Wc1 = a.W >> 1;     // rows / columns of parity 0 / 1

    // work item -> (n block, parity class, plane tile, image) of THIS CTA; `live` = the pair has a non-empty tile (same in both CTAs)
    struct Tile { int n_blk, py, px, ty, tx, img; bool live; };
    auto decode = [&](int w) {
        Tile t;
        int q = fast_div(w, a.div_nb);
        t.n_blk = w - q * a.n_blocks;
        const int cls = q & 3; q >>= 2;
        t.py = cls >> 1; t.px = cls & 1;
        int r = fast_div(q, a.div_ntx);
        const int txq = q - r * a.ntx;
        const int aa = fast_div(r, a.div_ty);
        t.ty = r - aa * a.tiles_y;
        t.tx = a.pair_img ? txq : 2 * txq + (int)cta_rank;
        t.img = a.pair_img ? 2 * aa + (int)cta_rank : aa;
        const int tx0 = a.pair_img ? txq : 2 * txq;
        t.live = (t.ty * U_TH < (t.py ? Hc1 : Hc0)) && (tx0 * U_TW < (t.px ? Wc1 : Wc0));
        return t;
    };
    auto full_a_sig = [&](int s) { return mapa_shared(full_a(s), 0); };
    auto full_b_sig = [&](int s) { return mapa_shared(full_b(s), 0); };

    if (warp == 0) {
        // ===================================================================== A producer
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            auto load = [&](const CUtensorMap* map, int c, int x, int y, int img) {
                mbar_wait(empty_a(stage), phase ^ 1u);
                if (leader) mbar_arrive_expect_tx(full_a(stage), 2 * U_A_BYTES);
                tma_load_4d_pair(a_base + (uint32_t)stage * U_A_STAGE, map, full_a_sig(stage), c, x, y, img);
                if (++stage == a.a_stages) { stage = 0; phase ^= 1u; }
            };
            for (int w = work_first; w < a.work_total; w += work_step) {
                const Tile t = decode(w);
                if (!t.live) continue;
                const int x0 = t.tx * U_TW, y0 = t.ty * U_TH;
                for (int ch = 0; ch < a.c0_chunks; ++ch)
                    for (int plane = 0; plane < 4; ++plane) {
                        const int qy = plane >> 1, qx = plane & 1;
                        // a plane of the OTHER parity is read at two offsets (taps 0 and 2): {-1, 0} for class parity 0, {0, +1} for 1
                        const int oy = (qy != t.py && t.py == 0) ? -1 : 0, ox = (qx != t.px && t.px == 0) ? -1 : 0;
                        load(&maps.skip[plane], ch * 64, x0 + ox, y0 + oy, t.img);
                    }
                for (int ch = 0; ch < a.cl_chunks; ++ch)
                    load(&tmLow, ch * 64, x0 - (1 - t.px), y0 - (1 - t.py), t.img);
            }
        }
    } else if (warp == 2) {
        // ===================================================================== B producer (this CTA's half of every weight block)
        if (lane == 0) {
            int slot = 0; uint32_t phase = 0;
            const int c0 = a.c0_chunks * 64, cl = a.cl_chunks * 64;
            auto load = [&](int k, int row) {
                mbar_wait(empty_b(slot), phase ^ 1u);
                if (leader) mbar_arrive_expect_tx(full_b(slot), 2 * B_BLOCK);
                tma_load_2d_pair(b_base + (uint32_t)slot * B_BLOCK, &tmB, full_b_sig(slot), k, row);
                if (++slot == a.b_slots) { slot = 0; phase ^= 1u; }
            };
            for (int w = work_first; w < a.work_total; w += work_step) {
                const Tile t = decode(w);
                if (!t.live) continue;
                const int row = t.n_blk * BLOCK_N + (int)cta_rank * B_ROWS;
                for (int ch = 0; ch < a.c0_chunks; ++ch)
                    for (int plane = 0; plane < 4; ++plane) {
                        const int nky = ((plane >> 1) != t.py) ? 2 : 1, nkx = ((plane & 1) != t.px) ? 2 : 1;
                        for (int iy = 0; iy < nky; ++iy)
                            for (int ix = 0; ix < nkx; ++ix) {
                                const int ky = nky == 2 ? 2 * iy : 1, kx = nkx == 2 ? 2 * ix : 1;
                                load((ky * 3 + kx) * c0 + ch * 64, row);
                            }
                    }
                const int kl = 9 * c0 + (t.py * 2 + t.px) * 4 * cl;
                for (int ch = 0; ch < a.cl_chunks; ++ch)
                    for (int tp = 0; tp < 4; ++tp) load(kl + tp * cl + ch * 64, row);
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer (leader CTA only, for both CTAs)
        if (leader) {
            constexpr uint32_t idesc = make_idesc(BLOCK_N, 256);
            constexpr uint64_t B_STEP = (uint64_t)(B_BLOCK >> 4);
            int sa = 0; uint32_t pa = 0;
            int sb = 0; uint32_t pb = 0;
            int acc = 0; uint32_t acc_phase = 0;
            const uint64_t db_base = make_sw128_desc(b_base);
            for (int w = work_first; w < a.work_total; w += work_step) {
                const Tile t = decode(w);
                if (!t.live) continue;
                mbar_wait(tempty(acc), acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
                uint32_t started = 0;
                const int units = a.c0_chunks * 4 + a.cl_chunks;
                for (int u = 0; u < units; ++u) {
                    int nky = 2, nkx = 2;                               // low unit: taps (dy, dx) in {0,1}^2
                    if (u < a.c0_chunks * 4) {
                        const int plane = u & 3;
                        nky = ((plane >> 1) != t.py) ? 2 : 1; nkx = ((plane & 1) != t.px) ? 2 : 1;
                    }
                    mbar_wait(full_a(sa), pa);
                    tc_fence_after();
                    const uint64_t da_stage = make_sw128_desc(a_base + (uint32_t)sa * U_A_STAGE, U_BW * 128);
                    for (int iy = 0; iy < nky; ++iy)
                        for (int ix = 0; ix < nkx; ++ix) {
                            mbar_wait(full_b(sb), pb);
                            tc_fence_after();
                            const uint64_t da = da_stage + (uint64_t)((iy * U_BW + ix) * 8);
                            const uint64_t db = db_base + (uint64_t)sb * B_STEP;
                            const bool last = (iy == nky - 1) && (ix == nkx - 1);
                            if (elect_one()) {
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    umma_bf16_pair(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (k != 0) ? 1u : started);
                                umma_commit_pair(empty_b(sb));
                                if (last) umma_commit_pair(empty_a(sa));
                            }
                            __syncwarp();
                            started = 1u;
                            if (++sb == a.b_slots) { sb = 0; pb ^= 1u; }
                        }
                    if (++sa == a.a_stages) { sa = 0; pa ^= 1u; }
                }
                if (elect_one()) umma_commit_pair(tfull(acc));
                __syncwarp();
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else {
        // ===================================================================== epilogue (warps 3..6 = TMEM lane quadrants 3,0,1,2)
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const int et = threadIdx.x - 96;
        const int lx = row & (U_TW - 1), ly = row >> 3;
        for (int c = et; c < a.c_out; c += U_EPI_THREADS) { s_scale[c] = a.scale[c]; s_shift[c] = a.shift[c]; }
        named_bar_sync(1, U_EPI_THREADS);
        int acc = 0; uint32_t acc_phase = 0;
        uint32_t store_groups = 0;
        const uint32_t tempty_sig0 = mapa_shared(tempty(0), 0), tempty_sig1 = mapa_shared(tempty(1), 0);
        for (int w = work_first; w < a.work_total; w += work_step) {
            const Tile t = decode(w);
            if (!t.live) continue;
            const int Y = 2 * (t.ty * U_TH + ly) + t.py, X = 2 * (t.tx * U_TW + lx) + t.px;
            // taps whose up-sampled pixel lies outside [0,Hu) x [0,Wu) carry no ConvTranspose bias: 9-bit mask of taps to take back out
            uint32_t emask = 0;
            if (Y < a.H && X < a.W && t.img < a.n_img) {
                uint32_t ry = 0, rx = 0;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    if (Y + k - 1 < 0 || Y + k - 1 >= a.Hu) ry |= 1u << k;
                    if (X + k - 1 < 0 || X + k - 1 >= a.Wu) rx |= 1u << k;
                }
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx)
                        if (((ry >> ky) | (rx >> kx)) & 1u) emask |= 1u << (ky * 3 + kx);
            }
            const float* t_scale = s_scale + t.n_blk * BLOCK_N;
            const float* t_shift = s_shift + t.n_blk * BLOCK_N;

            mbar_wait(tfull(acc), acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BLOCK_N);
            auto group = [&](const int c0, const bool first_half, const uint32_t (&r)[32]) {
                const float4* sc4 = reinterpret_cast<const float4*>(t_scale + c0);
                const float4* sh4 = reinterpret_cast<const float4*>(t_shift + c0);
                float2 yv[16];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 sc = sc4[i], sh = sh4[i];
                    yv[2 * i] = pk_fma(make_float2(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1])), make_float2(sc.x, sc.y), make_float2(sh.x, sh.y));
                    yv[2 * i + 1] = pk_fma(make_float2(__uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3])), make_float2(sc.z, sc.w), make_float2(sh.z, sh.w));
                }
                if (emask) {                                           // border pixels only (rare, divergent)
                    const float* wb = a.wb + t.n_blk * BLOCK_N + c0;
                    for (int tap = 0; tap < 9; ++tap)
                        if ((emask >> tap) & 1u) {
                            const float2* w2 = reinterpret_cast<const float2*>(wb + tap * a.c_out);
#pragma unroll
                            for (int i = 0; i < 16; ++i) { const float2 v = w2[i]; yv[i].x -= v.x; yv[i].y -= v.y; }
                        }
                }
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) pk[i] = pack_relu_bf16x2(yv[i].x, yv[i].y);
                const uint32_t cbase = first_half ? 0u : 4u;
                const uint32_t buf = store_groups & 1u;
                const uint32_t o_stage = smem_base + stage_off + buf * U_OUT_STAGE;
                if (first_half) {                                      // this buffer's previous TMA store must have read it
                    if (et == 0) bulk_wait_read<1>();
                    named_bar_sync(1, U_EPI_THREADS);
                }
                const uint32_t rbase = o_stage + (uint32_t)row * 128u;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    st_shared_v4(rbase + (((cbase + i) ^ ((uint32_t)row & 7u)) << 4), pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
                if (!first_half) {
                    fence_proxy_async();
                    named_bar_sync(1, U_EPI_THREADS);
                    if (et == 0) {
                        tma_store_4d(&maps.out[t.py * 2 + t.px], o_stage, t.n_blk * BLOCK_N + (c0 - 32), t.tx * U_TW, t.ty * U_TH, t.img);
                        bulk_commit();
                    }
                    ++store_groups;
                }
            };
            uint32_t r0[32], r1[32];
            tmem_ld32(t_row, r0);
#pragma unroll 1
            for (int c0 = 0; c0 < BLOCK_N; c0 += 64) {
                tmem_ld_wait();
                tmem_ld32(t_row + (uint32_t)(c0 + 32), r1);
                group(c0, true, r0);
                tmem_ld_wait();
                if (c0 + 64 < BLOCK_N) tmem_ld32(t_row + (uint32_t)(c0 + 64), r0);
                group(c0 + 32, false, r1);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(acc ? tempty_sig1 : tempty_sig0);
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
        if (et == 0) bulk_wait<0>();
    }

    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    if (warp == 2) tmem_dealloc_pair(tmem_own, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------ two parity classes per tile
// The kernel above reads all four skip planes for every parity class: 78 KB of TMA writes per 64-channel chunk and class.  With a
// 256-wide accumulator that is hidden; with CO = 128 output channels (decoder level of 128 channels) the 256x128x16 UMMAs are too
// short (measured 2.36 ms against 2.03 for ConvTranspose + conv).  conv3x3_upm2_kernel therefore computes BOTH column-parity classes
// (py, 0) and (py, 1) of a tile in one 256-column accumulator [class 0 | class 1]:
//   * every A view that both classes read (the un-shifted column of a plane, the middle column of the low tile) is ONE N = 256 UMMA
//     whose B block stacks the two classes' weights; the views only one class reads are N = 128 UMMAs into that class's half;
//   * the planes are loaded once for two classes, so the TMA writes per UMMA cycle drop from 66 to 48 B/clk and the mean
//     operand read from 96 to 75 B/clk.
// B comes from two packed tensors: Bsh [2*CO][..] (shared views) and B1 [CO][..] (single views), blocks in consumption order.
// Measured at batch 64 x (257,1034): upconv3 (ConvTranspose 256 -> 128 + conv over 128 + 128 channels) 0.38 + 1.69 -> 1.50-1.53 ms.
// Without a low tensor (cl_chunks = 0) the kernel is a plain Conv3x3 + BN + ReLU with 128 output channels in the same formulation
// (upconv3.3: 0.89 -> 0.80 ms against conv_halo's CTA-pair kernel).  POOL adds the fused MaxPool2d(2) of DownSampleLayer
// (model.py:31): the pool window of a half-resolution pixel is exactly its four parity classes, so every epilogue thread keeps a
// running bf16x2 maximum over the two column classes of a tile and over the two row-parity tiles of a region -- a CTA pair walks its
// regions with py = 0, 1 back to back -- and stores the pooled tile after the second one (downconv2.3: 1.01 -> 0.94 ms).
constexpr int U2_LOW_BW = U_TW + 2;                               // the low tile spans columns x0-1 .. x0+8 for the two classes
constexpr int U2_A_STAGE = ((U_BH * U2_LOW_BW * 128) + 1023) & ~1023;   // 22 528 B (skip plane boxes are 17 x 9)
constexpr int U2_B_SLOT = 128 * 128;                              // one CTA's half of a shared block (a single block uses half of it)

struct Upm2Args {
    int c0_chunks, cl_chunks;
    int n_img, H, W, Hu, Wu;
    int tiles_y, ntx, pair_img, work_total;
    FastDiv div_ntx, div_ty;
    int a_stages, b_slots, a_stage_bytes;
    int pool;                      // fused MaxPool2d(2) of the output (plain conv only): (n, H/2, W/2, CO) through tmPool
    const float* scale;
    const float* shift;
    const float* wb;
};

template <int CO, bool POOL>
__global__ void __launch_bounds__(U_THREADS, 1)
conv3x3_upm2_kernel(const __grid_constant__ UpmMaps maps, const __grid_constant__ CUtensorMap tmLow,
                    const __grid_constant__ CUtensorMap tmBsh, const __grid_constant__ CUtensorMap tmB1,
                    const __grid_constant__ CUtensorMap tmPool, const Upm2Args a) {
    constexpr int BLOCK_N = 2 * CO;
    static_assert(BLOCK_N == 256, "two classes of 128 output channels");
    constexpr int TMEM_COLS = 2 * BLOCK_N;
    const uint32_t cta_rank = cluster_ctarank();
    const bool leader = (cta_rank == 0);
    extern __shared__ uint8_t smem_dyn[];
    const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_dyn + (smem_base - smem_u32(smem_dyn));

    const uint32_t a_base = smem_base;
    const uint32_t A_STAGE = (uint32_t)a.a_stage_bytes;
    const uint32_t b_base = a_base + (uint32_t)a.a_stages * A_STAGE;
    const uint32_t stage_off = (uint32_t)a.a_stages * A_STAGE + (uint32_t)a.b_slots * U2_B_SLOT;
    const uint32_t pool_off = stage_off + 2u * U_OUT_STAGE;              // POOL: two more [128 px][64 ch] staging tiles
    const uint32_t aux_off = pool_off + (POOL ? 2u * U_OUT_STAGE : 0u);
    float* s_scale = reinterpret_cast<float*>(smem_gen + aux_off);
    float* s_shift = s_scale + CO;
    const uint32_t aux_f32 = (uint32_t)(2 * CO) * 4;
    const uint32_t bar_base = smem_base + aux_off + aux_f32;
    auto full_a = [&](int s) { return bar_base + 8u * s; };
    auto empty_a = [&](int s) { return bar_base + 8u * (U_MAX_A + s); };
    auto full_b = [&](int s) { return bar_base + 8u * (2 * U_MAX_A + s); };
    auto empty_b = [&](int s) { return bar_base + 8u * (2 * U_MAX_A + U_MAX_B + s); };
    auto tfull = [&](int s) { return bar_base + 8u * (2 * U_MAX_A + 2 * U_MAX_B + s); };
    auto tempty = [&](int s) { return bar_base + 8u * (2 * U_MAX_A + 2 * U_MAX_B + 2 + s); };
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem_gen + aux_off + aux_f32 + (2 * U_MAX_A + 2 * U_MAX_B + 4) * 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        for (int p = 0; p < 4; ++p) { tma_prefetch_desc(&maps.skip[p]); tma_prefetch_desc(&maps.out[p]); }
        tma_prefetch_desc(&tmLow); tma_prefetch_desc(&tmBsh); tma_prefetch_desc(&tmB1);
        if (POOL) tma_prefetch_desc(&tmPool);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < U_MAX_A; ++s) { mbar_init(full_a(s), 1); mbar_init(empty_a(s), 1); }
        for (int s = 0; s < U_MAX_B; ++s) { mbar_init(full_b(s), 1); mbar_init(empty_b(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull(s), 1); mbar_init(tempty(s), 2 * U_EPI_THREADS / 32); }
        fence_barrier_init();
    }
    if (warp == 2) { tmem_alloc_pair(smem_u32(tmem_ptr_smem), TMEM_COLS); tmem_relinquish_pair(); }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_own = *tmem_ptr_smem;
    const uint32_t tmem_base = ld_shared_cluster_u32(mapa_shared(smem_u32(tmem_ptr_smem), 0));

    const int work_first = (int)(blockIdx.x >> 1), work_step = (int)(gridDim.x >> 1);
    const int Hc0 = (a.H + 1) >> 1, Hc1 = a.H >> 1, Wc0 = (a.W + 1) >> 1;

    // iteration it of this pair = (region work_first + (it >> 1) * work_step, row parity it & 1): the two row-parity tiles of a
    // region run back to back in ONE pair (the fused pool combines them; it also keeps the planes of the region in L2)
    const int regions = a.work_total >> 1;
    struct Tile { int py, ty, tx, img; bool live, done; };
    auto decode = [&](int it) {
        Tile t;
        t.py = it & 1;
        const int q = work_first + (it >> 1) * work_step;
        t.done = q >= regions;
        const int r = fast_div(q, a.div_ntx);
        const int txq = q - r * a.ntx;
        const int aa = fast_div(r, a.div_ty);
        t.ty = r - aa * a.tiles_y;
        t.tx = a.pair_img ? txq : 2 * txq + (int)cta_rank;
        t.img = a.pair_img ? 2 * aa + (int)cta_rank : aa;
        const int tx0 = a.pair_img ? txq : 2 * txq;
        t.live = !t.done && (t.ty * U_TH < (t.py ? Hc1 : Hc0)) && (tx0 * U_TW < Wc0);
        return t;
    };
    auto full_a_sig = [&](int s) { return mapa_shared(full_a(s), 0); };
    auto full_b_sig = [&](int s) { return mapa_shared(full_b(s), 0); };
    const int NS = a.c0_chunks * 8 + a.cl_chunks * 2;              // 64-wide K blocks per row parity in Bsh / B1
    const int N1 = a.c0_chunks * 8 + a.cl_chunks * 4;

    if (warp == 0) {
        // ===================================================================== A producer
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            auto load = [&](const CUtensorMap* map, uint32_t bytes, int c, int x, int y, int img) {
                mbar_wait(empty_a(stage), phase ^ 1u);
                if (leader) mbar_arrive_expect_tx(full_a(stage), 2 * bytes);
                tma_load_4d_pair(a_base + (uint32_t)stage * A_STAGE, map, full_a_sig(stage), c, x, y, img);
                if (++stage == a.a_stages) { stage = 0; phase ^= 1u; }
            };
            for (int it = 0;; ++it) {
                const Tile t = decode(it);
                if (t.done) break;
                if (!t.live) continue;
                const int x0 = t.tx * U_TW, y0 = t.ty * U_TH;
                for (int ch = 0; ch < a.c0_chunks; ++ch)
                    for (int plane = 0; plane < 4; ++plane) {
                        const int qy = plane >> 1, qx = plane & 1;
                        const int oy = (qy != t.py && t.py == 0) ? -1 : 0;
                        load(&maps.skip[plane], U_A_BYTES, ch * 64, x0 - qx, y0 + oy, t.img);      // odd plane: columns x0-1 .. x0+7
                    }
                for (int ch = 0; ch < a.cl_chunks; ++ch)
                    load(&tmLow, U_BH * U2_LOW_BW * 128, ch * 64, x0 - 1, y0 - (1 - t.py), t.img);
            }
        }
    } else if (warp == 2) {
        // ===================================================================== B producer
        if (lane == 0) {
            int slot = 0; uint32_t phase = 0;
            auto load_sh = [&](int kb) {
                mbar_wait(empty_b(slot), phase ^ 1u);
                if (leader) mbar_arrive_expect_tx(full_b(slot), 2 * 128 * 128);
                tma_load_2d_pair(b_base + (uint32_t)slot * U2_B_SLOT, &tmBsh, full_b_sig(slot), kb * 64, (int)cta_rank * 128);
                if (++slot == a.b_slots) { slot = 0; phase ^= 1u; }
            };
            auto load_1 = [&](int kb) {
                mbar_wait(empty_b(slot), phase ^ 1u);
                if (leader) mbar_arrive_expect_tx(full_b(slot), 2 * 64 * 128);
                tma_load_2d_pair(b_base + (uint32_t)slot * U2_B_SLOT, &tmB1, full_b_sig(slot), kb * 64, (int)cta_rank * 64);
                if (++slot == a.b_slots) { slot = 0; phase ^= 1u; }
            };
            for (int it = 0;; ++it) {
                const Tile t = decode(it);
                if (t.done) break;
                if (!t.live) continue;
                for (int ch = 0; ch < a.c0_chunks; ++ch)
                    for (int plane = 0; plane < 4; ++plane) {
                        const int nky = ((plane >> 1) != t.py) ? 2 : 1;
                        for (int iy = 0; iy < nky; ++iy) {
                            const int idx = (ch * 4 + plane) * 2 + iy;
                            load_sh(t.py * NS + idx);
                            load_1(t.py * N1 + idx);
                        }
                    }
                for (int ch = 0; ch < a.cl_chunks; ++ch)
                    for (int dy = 0; dy < 2; ++dy) {
                        load_sh(t.py * NS + a.c0_chunks * 8 + ch * 2 + dy);
                        load_1(t.py * N1 + a.c0_chunks * 8 + (ch * 2 + dy) * 2);
                        load_1(t.py * N1 + a.c0_chunks * 8 + (ch * 2 + dy) * 2 + 1);
                    }
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer
        if (leader) {
            constexpr uint32_t idesc_sh = make_idesc(BLOCK_N, 256), idesc_1 = make_idesc(CO, 256);
            constexpr uint64_t B_STEP = (uint64_t)(U2_B_SLOT >> 4);
            int sa = 0; uint32_t pa = 0;
            int sb = 0; uint32_t pb = 0;
            int acc = 0; uint32_t acc_phase = 0;
            const uint64_t db_base = make_sw128_desc(b_base);
            // one B slot: four k-steps of a UMMA with `idesc` into accumulator columns d, A view at `da`
            auto issue = [&](uint32_t d, uint64_t da, uint32_t idesc, uint32_t first_acc, bool release_a, int a_stage) {
                mbar_wait(full_b(sb), pb);
                tc_fence_after();
                const uint64_t db = db_base + (uint64_t)sb * B_STEP;
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16_pair(d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (k != 0) ? 1u : first_acc);
                    umma_commit_pair(empty_b(sb));
                    if (release_a) umma_commit_pair(empty_a(a_stage));
                }
                __syncwarp();
                if (++sb == a.b_slots) { sb = 0; pb ^= 1u; }
            };
            for (int it = 0;; ++it) {
                const Tile t = decode(it);
                if (t.done) break;
                if (!t.live) continue;
                mbar_wait(tempty(acc), acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
                uint32_t started = 0;
                for (int u = 0; u < a.c0_chunks * 4; ++u) {
                    const int plane = u & 3, qx = plane & 1;
                    const int nky = ((plane >> 1) != t.py) ? 2 : 1;
                    mbar_wait(full_a(sa), pa);
                    tc_fence_after();
                    const uint64_t da_stage = make_sw128_desc(a_base + (uint32_t)sa * A_STAGE, U_BW * 128);
                    for (int iy = 0; iy < nky; ++iy) {
                        // shared view: box column qx (both classes); single view: box column 1 - qx, class 1 - qx only
                        issue(d_tmem, da_stage + (uint64_t)((iy * U_BW + qx) * 8), idesc_sh, started, false, sa);
                        started = 1u;
                        issue(d_tmem + (uint32_t)((1 - qx) * CO), da_stage + (uint64_t)((iy * U_BW + (1 - qx)) * 8), idesc_1, 1u, iy == nky - 1, sa);
                    }
                    if (++sa == a.a_stages) { sa = 0; pa ^= 1u; }
                }
                for (int ch = 0; ch < a.cl_chunks; ++ch) {
                    mbar_wait(full_a(sa), pa);
                    tc_fence_after();
                    const uint64_t da_stage = make_sw128_desc(a_base + (uint32_t)sa * A_STAGE, U2_LOW_BW * 128);
                    for (int dy = 0; dy < 2; ++dy) {
                        issue(d_tmem, da_stage + (uint64_t)((dy * U2_LOW_BW + 1) * 8), idesc_sh, 1u, false, sa);
                        issue(d_tmem, da_stage + (uint64_t)((dy * U2_LOW_BW + 0) * 8), idesc_1, 1u, false, sa);
                        issue(d_tmem + (uint32_t)CO, da_stage + (uint64_t)((dy * U2_LOW_BW + 2) * 8), idesc_1, 1u, dy == 1, sa);
                    }
                    if (++sa == a.a_stages) { sa = 0; pa ^= 1u; }
                }
                if (elect_one()) umma_commit_pair(tfull(acc));
                __syncwarp();
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else {
        // ===================================================================== epilogue
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const int et = threadIdx.x - 96;
        const int lx = row & (U_TW - 1), ly = row >> 3;
        for (int c = et; c < CO; c += U_EPI_THREADS) { s_scale[c] = a.scale[c]; s_shift[c] = a.shift[c]; }
        named_bar_sync(1, U_EPI_THREADS);
        int acc = 0; uint32_t acc_phase = 0;
        uint32_t store_groups = 0, pool_groups = 0;
        uint32_t rmax[POOL ? CO / 2 : 1];                                // POOL: running 2x2 maximum of this plane pixel, bf16x2 per channel pair
        const uint32_t tempty_sig0 = mapa_shared(tempty(0), 0), tempty_sig1 = mapa_shared(tempty(1), 0);
        for (int it = 0;; ++it) {
            const Tile t = decode(it);
            if (t.done) break;
            if (!t.live) continue;
            const int Y = 2 * (t.ty * U_TH + ly) + t.py;
            uint32_t emask2[2] = {0u, 0u};
            if (a.cl_chunks > 0 && Y < a.H && t.img < a.n_img) {          // no low tensor (plain 3x3 conv in parity space): no bias terms
                uint32_t ry = 0;
#pragma unroll
                for (int k = 0; k < 3; ++k) if (Y + k - 1 < 0 || Y + k - 1 >= a.Hu) ry |= 1u << k;
#pragma unroll
                for (int px = 0; px < 2; ++px) {
                    const int X = 2 * (t.tx * U_TW + lx) + px;
                    if (X >= a.W) continue;
                    uint32_t rx = 0;
#pragma unroll
                    for (int k = 0; k < 3; ++k) if (X + k - 1 < 0 || X + k - 1 >= a.Wu) rx |= 1u << k;
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx)
                            if (((ry >> ky) | (rx >> kx)) & 1u) emask2[px] |= 1u << (ky * 3 + kx);
                }
            }
            mbar_wait(tfull(acc), acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BLOCK_N);
            auto group = [&](const int c0, const bool first_half, const uint32_t (&r)[32]) {
                const int px = c0 / CO, cc = c0 - px * CO;                     // class and channel offset of this 32-column group
                const uint32_t emask = px ? emask2[1] : emask2[0];
                const float4* sc4 = reinterpret_cast<const float4*>(s_scale + cc);
                const float4* sh4 = reinterpret_cast<const float4*>(s_shift + cc);
                float2 yv[16];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 sc = sc4[i], sh = sh4[i];
                    yv[2 * i] = pk_fma(make_float2(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1])), make_float2(sc.x, sc.y), make_float2(sh.x, sh.y));
                    yv[2 * i + 1] = pk_fma(make_float2(__uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3])), make_float2(sc.z, sc.w), make_float2(sh.z, sh.w));
                }
                if (emask) {
                    const float* wb = a.wb + cc;
                    for (int tap = 0; tap < 9; ++tap)
                        if ((emask >> tap) & 1u) {
                            const float2* w2 = reinterpret_cast<const float2*>(wb + tap * CO);
#pragma unroll
                            for (int i = 0; i < 16; ++i) { const float2 v = w2[i]; yv[i].x -= v.x; yv[i].y -= v.y; }
                        }
                }
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) pk[i] = pack_relu_bf16x2(yv[i].x, yv[i].y);
                const uint32_t cbase = first_half ? 0u : 4u;
                const uint32_t buf = store_groups & 1u;
                const uint32_t o_stage = smem_base + stage_off + buf * U_OUT_STAGE;
                if (POOL) {
                    // the 2x2 pool window of plane pixel (y, x) is its four parity classes: columns c (px = 0) and CO + c (px = 1) of
                    // the tiles py = 0 and py = 1.  Post-ReLU values: the first class initialises, the others take the maximum.
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        uint32_t& m = rmax[POOL ? cc / 2 + i : 0];
                        m = (t.py == 0 && px == 0) ? pk[i] : bf16x2_max(m, pk[i]);
                    }
                }
                if (first_half) {
                    if (et == 0) bulk_wait_read<1>();
                    named_bar_sync(1, U_EPI_THREADS);
                }
                const uint32_t rbase = o_stage + (uint32_t)row * 128u;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    st_shared_v4(rbase + (((cbase + i) ^ ((uint32_t)row & 7u)) << 4), pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
                if (!first_half) {
                    const bool pool_now = POOL && t.py == 1 && px == 1;   // the last class of this 64-channel block: its maximum is final
                    const uint32_t p_stage = smem_base + pool_off + (pool_groups & 1u) * U_OUT_STAGE;
                    if (pool_now) {
                        const uint32_t prow = p_stage + (uint32_t)row * 128u;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int j = POOL ? (cc - 32) / 2 + 4 * i : 0;
                            st_shared_v4(prow + (((uint32_t)i ^ ((uint32_t)row & 7u)) << 4), rmax[j], rmax[j + (POOL ? 1 : 0)], rmax[j + (POOL ? 2 : 0)],
                                         rmax[j + (POOL ? 3 : 0)]);
                        }
                    }
                    fence_proxy_async();
                    named_bar_sync(1, U_EPI_THREADS);
                    if (et == 0) {
                        tma_store_4d(&maps.out[t.py * 2 + px], o_stage, cc - 32, t.tx * U_TW, t.ty * U_TH, t.img);
                        if (pool_now) tma_store_4d(&tmPool, p_stage, cc - 32, t.tx * U_TW, t.ty * U_TH, t.img);
                        bulk_commit();
                    }
                    ++store_groups;
                    if (pool_now) ++pool_groups;
                }
            };
            uint32_t r0[32], r1[32];
            tmem_ld32(t_row, r0);
            if (POOL) {
#pragma unroll
                for (int c0 = 0; c0 < BLOCK_N; c0 += 64) {               // unrolled: the running-maximum registers are indexed by column
                    tmem_ld_wait();
                    tmem_ld32(t_row + (uint32_t)(c0 + 32), r1);
                    group(c0, true, r0);
                    tmem_ld_wait();
                    if (c0 + 64 < BLOCK_N) tmem_ld32(t_row + (uint32_t)(c0 + 64), r0);
                    group(c0 + 32, false, r1);
                }
            } else {
#pragma unroll 1
                for (int c0 = 0; c0 < BLOCK_N; c0 += 64) {
                    tmem_ld_wait();
                    tmem_ld32(t_row + (uint32_t)(c0 + 32), r1);
                    group(c0, true, r0);
                    tmem_ld_wait();
                    if (c0 + 64 < BLOCK_N) tmem_ld32(t_row + (uint32_t)(c0 + 64), r0);
                    group(c0 + 32, false, r1);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(acc ? tempty_sig1 : tempty_sig0);
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
        if (et == 0) bulk_wait<0>();
    }

    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    if (warp == 2) tmem_dealloc_pair(tmem_own, TMEM_COLS);
}

// Bsh / B1 for conv3x3_upm2_kernel, gathered from the merged tensor of the one-class kernel (wm: [co][9*c0 + 16*cl], see below).
// Block order per row parity py: skip (chunk, plane, iy in {0,1}), then low (chunk, dy) [B1: (chunk, dy, s)].  Unused iy slots are zero.
__global__ void upm2_repack_kernel(const __nv_bfloat16* __restrict__ wm, int co_n, int c0, int cl, __nv_bfloat16* __restrict__ bsh,
                                   __nv_bfloat16* __restrict__ b1) {
    const int c0c = c0 / 64, clc = cl / 64;
    const int NS = c0c * 8 + clc * 2, N1 = c0c * 8 + clc * 4;
    const int ktot = 9 * c0 + 16 * cl;
    const long long n_sh = 2ll * co_n * 2 * NS * 64, n_1 = (long long)co_n * 2 * N1 * 64;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_sh + n_1; i += (long long)gridDim.x * blockDim.x) {
        const bool sh = i < n_sh;
        const long long e = sh ? i : i - n_sh;
        const int kw = 2 * (sh ? NS : N1) * 64;                     // row length
        const int r = (int)(e / kw), k = (int)(e % kw);
        const int j = k & 63, kb = k >> 6;
        const int nb = sh ? NS : N1;
        const int py = kb / nb, b = kb % nb;
        const int co = sh ? r % co_n : r;
        int px = sh ? r / co_n : 0;
        float v = 0.f;
        bool zero = false;
        int src = 0;
        if (b < c0c * 8) {                                           // skip block
            const int iy = b & 1, plane = (b >> 1) & 3, ch = b >> 3;
            const int qy = plane >> 1, qx = plane & 1;
            const int nky = (qy != py) ? 2 : 1;
            if (iy >= nky) zero = true;
            const int ky = nky == 2 ? 2 * iy : 1;
            const int kx = sh ? ((px == qx) ? 1 : (qx == 1 ? 2 : 0)) : (qx == 1 ? 0 : 2);
            src = (ky * 3 + kx) * c0 + ch * 64 + j;
        } else {
            const int bl = b - c0c * 8;
            int ch, dy, dx;
            if (sh) { ch = bl >> 1; dy = bl & 1; dx = (px == 0) ? 1 : 0; }
            else { ch = bl >> 2; dy = (bl >> 1) & 1; px = bl & 1; dx = px; }
            src = 9 * c0 + (((py * 2 + px) * 4) + dy * 2 + dx) * cl + ch * 64 + j;
        }
        if (!zero) v = __bfloat162float(wm[(long long)co * ktot + src]);
        (sh ? bsh : b1)[e] = __float2bfloat16_rn(v);
    }
}

// ------------------------------------------------------------------------------------------------ weight merge (checkpoint load time)
// out[co][k]: k < 9*c0: skip weights [tap][c] = W3[co][c][tap];  then [cls = py*2+px][t = dy*2+dx][ci] =
//   sum_{ky -> dy, kx -> dx} sum_cu W3[co][c0 + cu][ky][kx] * Wt[ci][cu][(py+ky-1)&1][(px+kx-1)&1]       (fp32 sums, one bf16 rounding)
// grid (cl/32, c_out/32, 16), block (32, 32): a plain shared-memory-tiled fp32 product, a few ms per level.
__device__ __forceinline__ int upm_dof(int p, int k) { return ((p + k + 1) >> 1) - p; }   // low row index 0/1 hit by tap k of class parity p

__global__ void __launch_bounds__(1024) upm_merge_low_kernel(const float* __restrict__ w3, const float* __restrict__ wt, int c_out, int c0,
                                                             int cup, int cl, __nv_bfloat16* __restrict__ out) {
    __shared__ float sa[32][33], sb[32][33];
    const int cls = blockIdx.z >> 2, tp = blockIdx.z & 3;
    const int py = cls >> 1, px = cls & 1, dy = tp >> 1, dx = tp & 1;
    const int co0 = blockIdx.y * 32, ci0 = blockIdx.x * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int cin3 = c0 + cup;
    float acc = 0.f;
    for (int ky = 0; ky < 3; ++ky) {
        if (upm_dof(py, ky) != dy) continue;
        for (int kx = 0; kx < 3; ++kx) {
            if (upm_dof(px, kx) != dx) continue;
            const int q = (((py + ky + 1) & 1) << 1) | ((px + kx + 1) & 1);
            for (int cu0 = 0; cu0 < cup; cu0 += 32) {
                // sa[co][cu] = W3[co0+ty][c0+cu0+tx][ky][kx] ; sb[ci][cu] = Wt[ci0+ty][cu0+tx][q]
                sa[ty][tx] = (co0 + ty < c_out && cu0 + tx < cup) ? w3[((long long)(co0 + ty) * cin3 + c0 + cu0 + tx) * 9 + ky * 3 + kx] : 0.f;
                sb[ty][tx] = (ci0 + ty < cl && cu0 + tx < cup) ? wt[((long long)(ci0 + ty) * cup + cu0 + tx) * 4 + q] : 0.f;
                __syncthreads();
#pragma unroll
                for (int j = 0; j < 32; ++j) acc = fmaf(sa[ty][j], sb[tx][j], acc);
                __syncthreads();
            }
        }
    }
    if (co0 + ty < c_out && ci0 + tx < cl)
        out[(long long)(co0 + ty) * (9 * c0 + 16 * cl) + 9 * c0 + (cls * 4 + tp) * cl + ci0 + tx] = __float2bfloat16_rn(acc);
}

__global__ void upm_pack_skip_kernel(const float* __restrict__ w3, int c_out, int c0, int cup, int cl, __nv_bfloat16* __restrict__ out) {
    const long long total = (long long)c_out * 9 * c0;
    const int ktot = 9 * c0 + 16 * cl;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % c0);
        const int tap = (int)((i / c0) % 9);
        const int co = (int)(i / ((long long)c0 * 9));
        out[(long long)co * ktot + tap * c0 + c] = __float2bfloat16_rn(w3[((long long)co * (c0 + cup) + c) * 9 + tap]);
    }
}

// wb[tap][co] = scale[co] * sum_cu W3[co][c0+cu][tap] * bt[cu];  shift_m[co] = shift[co] + sum_tap wb[tap][co]
__global__ void upm_bias_kernel(const float* __restrict__ w3, const float* __restrict__ bt, const float* __restrict__ scale,
                                const float* __restrict__ shift, int c_out, int c0, int cup, float* __restrict__ shift_m, float* __restrict__ wb) {
    const int co = blockIdx.x * blockDim.x + threadIdx.x;
    if (co >= c_out) return;
    float tot = 0.f;
    for (int tap = 0; tap < 9; ++tap) {
        float s = 0.f;
        for (int cu = 0; cu < cup; ++cu) s = fmaf(w3[((long long)co * (c0 + cup) + c0 + cu) * 9 + tap], bt[cu], s);
        s *= scale[co];
        wb[tap * c_out + co] = s;
        tot += s;
    }
    shift_m[co] = shift[co] + tot;
}

// ------------------------------------------------------------------------------------------------ host side
// strided parity-plane view of an NHWC bf16 tensor: pixels (2r+qy, 2c+qx)
static int make_plane_map(CUtensorMap* map, const void* ptr, int n, int h, int w, int c, int qy, int qx, int bw, int bh) {
    PFN_tmapEncodeTiled enc = get_encode_fn();
    if (!enc) return ADN_ERR_DRIVER;
    const int hq = (h - qy + 1) / 2, wq = (w - qx + 1) / 2;
    if (hq < 1 || wq < 1) return ADN_ERR_ARG;
    const char* base = static_cast<const char*>(ptr) + ((long long)qy * w + qx) * c * 2;
    cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)wq, (cuuint64_t)hq, (cuuint64_t)n};
    cuuint64_t strides[3] = {(cuuint64_t)c * 4, (cuuint64_t)w * c * 4, (cuuint64_t)h * w * c * 2};
    cuuint32_t box[4] = {64u, (cuuint32_t)bw, (cuuint32_t)bh, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<char*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? ADN_OK : ADN_ERR_DRIVER;
}

template <int BLOCK_N>
static int launch_upm(const UpmMaps& maps, const CUtensorMap& mLow, const CUtensorMap& mB, UpmArgs& args, cudaStream_t stream) {
    constexpr int B_BLOCK = (BLOCK_N / 2) * 128;
    const int AUX = 2 * args.c_out * 4 + (2 * U_MAX_A + 2 * U_MAX_B + 4) * 8 + 16;
    constexpr int MAX_DYN = 232448;
    const int budget = MAX_DYN - 1024 - AUX - 2 * U_OUT_STAGE;
    args.a_stages = (BLOCK_N >= 256) ? 4 : 6;
    int sl = (budget - args.a_stages * U_A_STAGE) / B_BLOCK;
    args.b_slots = sl > U_MAX_B ? U_MAX_B : sl;
    if (args.b_slots < 4) return ADN_ERR_ARG;
    const int smem = 1024 + args.a_stages * U_A_STAGE + args.b_slots * B_BLOCK + 2 * U_OUT_STAGE + AUX;
    const int max_pairs = num_sms() / 2;
    const int grid = 2 * (args.work_total < max_pairs ? args.work_total : max_pairs);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid, 1, 1);
    cfg.blockDim = dim3(U_THREADS, 1, 1);
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    static unsigned char smem_set[64] = {0};
    ADN_CUDA_TRY(ensure_dyn_smem(conv3x3_upm_kernel<BLOCK_N>, MAX_DYN, smem_set));
    ADN_CUDA_TRY(cudaLaunchKernelEx(&cfg, conv3x3_upm_kernel<BLOCK_N>, maps, mLow, mB, args));
    ADN_LAUNCH_CHECK();
    return ADN_OK;
}

template <bool POOL>
static int launch_upm2(const UpmMaps& maps, const CUtensorMap& mLow, const CUtensorMap& mBsh, const CUtensorMap& mB1, const CUtensorMap& mPool,
                       Upm2Args& args, cudaStream_t stream) {
    const int AUX = 2 * 128 * 4 + (2 * U_MAX_A + 2 * U_MAX_B + 4) * 8 + 16;
    constexpr int MAX_DYN = 232448;
    const int staging = (POOL ? 4 : 2) * U_OUT_STAGE;
    const int budget = MAX_DYN - 1024 - AUX - staging;
    args.a_stage_bytes = args.cl_chunks > 0 ? U2_A_STAGE : U_A_STAGE;       // without a low tensor every box is a 17 x 9 skip plane
    args.a_stages = POOL ? 3 : 4;
    int sl = (budget - args.a_stages * args.a_stage_bytes) / U2_B_SLOT;
    args.b_slots = sl > U_MAX_B ? U_MAX_B : sl;
    if (args.b_slots < 4) return ADN_ERR_ARG;
    const int smem = 1024 + args.a_stages * args.a_stage_bytes + args.b_slots * U2_B_SLOT + staging + AUX;
    const int max_pairs = num_sms() / 2;
    const int grid = 2 * (args.work_total < max_pairs ? args.work_total : max_pairs);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid, 1, 1);
    cfg.blockDim = dim3(U_THREADS, 1, 1);
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    static unsigned char smem_set[64] = {0};
    ADN_CUDA_TRY(ensure_dyn_smem(conv3x3_upm2_kernel<128, POOL>, MAX_DYN, smem_set));
    ADN_CUDA_TRY(cudaLaunchKernelEx(&cfg, conv3x3_upm2_kernel<128, POOL>, maps, mLow, mBsh, mB1, mPool, args));
    ADN_LAUNCH_CHECK();
    return ADN_OK;
}

}  // namespace adn

using namespace adn;

extern "C" int64_t adn_upmerged_pair_weight_elems(int c_out, int c0, int cl, int which) {
    const int c0c = c0 / 64, clc = cl / 64;
    return which == 0 ? 2ll * c_out * 2 * (c0c * 8 + clc * 2) * 64 : (long long)c_out * 2 * (c0c * 8 + clc * 4) * 64;
}

extern "C" int adn_pack_upmerged_pair_weight_bf16(const void* w_merged, int c_out, int c0, int cl, void* bsh, void* b1, void* stream) {
    if (!w_merged || !bsh || !b1 || c_out <= 0 || c0 <= 0 || (c0 % 64) || cl < 0 || (cl % 64)) return ADN_ERR_ARG;
    int st = check_device(); if (st != ADN_OK) return st;
    upm2_repack_kernel<<<num_sms() * 8, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)w_merged, c_out, c0, cl, (__nv_bfloat16*)bsh,
                                                                       (__nv_bfloat16*)b1);
    ADN_LAUNCH_CHECK();
    return ADN_OK;
}

static int upm2_entry(const void* skip, int c0, const void* low, int cl, int hl, int wl, int n, int h, int w, const void* bsh, const void* b1,
                      int c_out, const float* scale, const float* shift_m, const float* wb, void* out, void* pool_out, void* stream) {
    if (pool_out && (cl > 0 || !aligned16(pool_out))) return ADN_ERR_ARG;
    if (!skip || !bsh || !b1 || !scale || !shift_m || !out) return ADN_ERR_ARG;
    if (n <= 0 || h < 2 || w < 2 || c_out != 128) return ADN_ERR_ARG;
    if (c0 <= 0 || (c0 % 64) || cl < 0 || (cl % 64)) return ADN_ERR_ARG;
    if (cl > 0) {
        if (!low || !wb || hl < 1 || wl < 1 || !aligned16(low)) return ADN_ERR_ARG;
        if (h - 2 * hl < 0 || h - 2 * hl > 1 || w - 2 * wl < 0 || w - 2 * wl > 1) return ADN_ERR_ARG;
    }
    if (!aligned16(skip) || !aligned16(bsh) || !aligned16(b1) || !aligned16(out)) return ADN_ERR_ARG;
    int st = check_device();
    if (st != ADN_OK) return st;
    Upm2Args args;
    args.c0_chunks = c0 / 64; args.cl_chunks = cl / 64;
    args.n_img = n; args.H = h; args.W = w; args.Hu = 2 * hl; args.Wu = 2 * wl;
    const int hc = (h + 1) / 2, wc = (w + 1) / 2;
    args.tiles_y = (hc + U_TH - 1) / U_TH;
    const int tiles_x = (wc + U_TW - 1) / U_TW;
    args.pair_img = (n % 2 == 0) ? 1 : 0;
    args.ntx = args.pair_img ? tiles_x : (tiles_x + 1) / 2;
    const int na = args.pair_img ? n / 2 : n;
    const long long work = (long long)na * args.tiles_y * args.ntx * 2;
    if (work > 0x7fffffffLL) return ADN_ERR_ARG;
    args.work_total = (int)work;
    args.div_ntx = make_fastdiv(args.ntx); args.div_ty = make_fastdiv(args.tiles_y);
    args.scale = scale; args.shift = shift_m; args.wb = wb;
    UpmMaps maps;
    for (int p = 0; p < 4; ++p) {
        st = make_plane_map(&maps.skip[p], skip, n, h, w, c0, p >> 1, p & 1, U_BW, U_BH);
        if (st != ADN_OK) return st;
        st = make_plane_map(&maps.out[p], out, n, h, w, c_out, p >> 1, p & 1, U_TW, U_TH);
        if (st != ADN_OK) return st;
    }
    CUtensorMap mLow, mBsh, mB1;
    if (cl > 0) st = make_act_map(&mLow, low, n, hl, wl, cl, U2_LOW_BW, U_BH); else mLow = maps.skip[0];
    if (st != ADN_OK) return st;
    const int c0c = c0 / 64, clc = cl / 64;
    st = make_weight_map(&mBsh, bsh, 2 * c_out, 2 * (c0c * 8 + clc * 2) * 64, 128);
    if (st != ADN_OK) return st;
    st = make_weight_map(&mB1, b1, c_out, 2 * (c0c * 8 + clc * 4) * 64, 64);
    if (st != ADN_OK) return st;
    args.pool = pool_out ? 1 : 0;
    CUtensorMap mPool = mLow;
    if (pool_out) {
        st = make_act_map(&mPool, pool_out, n, h / 2, w / 2, c_out, U_TW, U_TH);
        if (st != ADN_OK) return st;
        return launch_upm2<true>(maps, mLow, mBsh, mB1, mPool, args, (cudaStream_t)stream);
    }
    return launch_upm2<false>(maps, mLow, mBsh, mB1, mPool, args, (cudaStream_t)stream);
}

extern "C" int adn_conv3x3_upmerged_pair_bn_relu_bf16(const void* skip, int c0, const void* low, int cl, int hl, int wl, int n, int h, int w,
                                                      const void* bsh, const void* b1, int c_out, const float* scale, const float* shift_m,
                                                      const float* wb, void* out, void* stream) {
    return upm2_entry(skip, c0, low, cl, hl, wl, n, h, w, bsh, b1, c_out, scale, shift_m, wb, out, nullptr, stream);
}

extern "C" int adn_conv3x3_pair_bn_relu_pool_bf16(const void* src, int c_in, int n, int h, int w, const void* bsh, const void* b1, int c_out,
                                                  const float* scale, const float* shift, void* out, void* pool_out, void* stream) {
    if (!pool_out || h < 2 || w < 2) return ADN_ERR_ARG;
    return upm2_entry(src, c_in, nullptr, 0, 0, 0, n, h, w, bsh, b1, c_out, scale, shift, nullptr, out, pool_out, stream);
}

extern "C" int adn_conv3x3_upmerged_bn_relu_bf16(const void* skip, int c0, const void* low, int cl, int hl, int wl, int n, int h, int w,
                                                 const void* w_merged, int c_out, const float* scale, const float* shift_m,
                                                 const float* wb, void* out, void* stream) {
    if (!skip || !low || !w_merged || !scale || !shift_m || !wb || !out) return ADN_ERR_ARG;
    if (n <= 0 || h < 2 || w < 2 || hl < 1 || wl < 1) return ADN_ERR_ARG;
    if (c0 <= 0 || (c0 % 64) || cl <= 0 || (cl % 64) || c_out <= 0 || (c_out % 128)) return ADN_ERR_ARG;
    if (h - 2 * hl < 0 || h - 2 * hl > 1 || w - 2 * wl < 0 || w - 2 * wl > 1) return ADN_ERR_ARG;     // F.pad with diff in {0, 1}: no rows / columns in front
    if (!aligned16(skip) || !aligned16(low) || !aligned16(w_merged) || !aligned16(out)) return ADN_ERR_ARG;
    int st = check_device();
    if (st != ADN_OK) return st;

    const int block_n = (c_out % 256 == 0) ? 256 : 128;
    UpmArgs args;
    args.c0_chunks = c0 / 64; args.cl_chunks = cl / 64;
    args.n_img = n; args.H = h; args.W = w; args.Hu = 2 * hl; args.Wu = 2 * wl;
    const int hc = (h + 1) / 2, wc = (w + 1) / 2;
    args.tiles_y = (hc + U_TH - 1) / U_TH;
    const int tiles_x = (wc + U_TW - 1) / U_TW;
    args.pair_img = (n % 2 == 0) ? 1 : 0;
    args.ntx = args.pair_img ? tiles_x : (tiles_x + 1) / 2;
    const int na = args.pair_img ? n / 2 : n;
    args.c_out = c_out; args.n_blocks = c_out / block_n;
    const long long work = (long long)na * args.tiles_y * args.ntx * 4 * args.n_blocks;
    if (work > 0x7fffffffLL) return ADN_ERR_ARG;
    args.work_total = (int)work;
    args.div_nb = make_fastdiv(args.n_blocks); args.div_ntx = make_fastdiv(args.ntx); args.div_ty = make_fastdiv(args.tiles_y);
    args.scale = scale; args.shift = shift_m; args.wb = wb;

    UpmMaps maps;
    for (int p = 0; p < 4; ++p) {
        st = make_plane_map(&maps.skip[p], skip, n, h, w, c0, p >> 1, p & 1, U_BW, U_BH);
        if (st != ADN_OK) return st;
        st = make_plane_map(&maps.out[p], out, n, h, w, c_out, p >> 1, p & 1, U_TW, U_TH);
        if (st != ADN_OK) return st;
    }
    CUtensorMap mLow, mB;
    st = make_act_map(&mLow, low, n, hl, wl, cl, U_BW, U_BH);
    if (st != ADN_OK) return st;
    st = make_weight_map(&mB, w_merged, c_out, 9 * c0 + 16 * cl, block_n / 2);
    if (st != ADN_OK) return st;
    return block_n == 256 ? launch_upm<256>(maps, mLow, mB, args, (cudaStream_t)stream)
                          : launch_upm<128>(maps, mLow, mB, args, (cudaStream_t)stream);
}

extern "C" int adn_pack_upmerged_weight_bf16(const float* w3, const float* wt, const float* bt, const float* scale, const float* shift,
                                             int c_out, int c0, int cup, int cl, void* w_merged, float* shift_m, float* wb, void* stream) {
    if (!w3 || !wt || !bt || !scale || !shift || !w_merged || !shift_m || !wb) return ADN_ERR_ARG;
    if (c_out <= 0 || c0 <= 0 || cup <= 0 || cl <= 0) return ADN_ERR_ARG;
    int st = check_device(); if (st != ADN_OK) return st;
    cudaStream_t s = (cudaStream_t)stream;
    long long g = ((long long)c_out * 9 * c0 + 255) / 256;
    if (g > 4096) g = 4096;
    upm_pack_skip_kernel<<<(unsigned)g, 256, 0, s>>>(w3, c_out, c0, cup, cl, (__nv_bfloat16*)w_merged);
    ADN_LAUNCH_CHECK();
    upm_merge_low_kernel<<<dim3((cl + 31) / 32, (c_out + 31) / 32, 16), dim3(32, 32), 0, s>>>(w3, wt, c_out, c0, cup, cl, (__nv_bfloat16*)w_merged);
    ADN_LAUNCH_CHECK();
    upm_bias_kernel<<<(c_out + 127) / 128, 128, 0, s>>>(w3, bt, scale, shift, c_out, c0, cup, shift_m, wb);
    ADN_LAUNCH_CHECK();
    return ADN_OK;
}
