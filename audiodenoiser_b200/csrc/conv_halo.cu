// conv_halo.cu -- Conv2d 3x3 (pad 1) + folded BatchNorm + ReLU of the reference UNet (code/model.py:11-16), the channel
// concat + F.pad in front of the decoder convs (model.py:44-49) and the fused 1x1 head (model.py:68,93), as a persistent
// warp-specialised tcgen05/TMEM implicit GEMM whose A operand is a HALO TILE RESIDENT IN SHARED MEMORY.
//
//   GEMM view   D[M = 128 pixels (16 rows x 8 cols), N = BLOCK_N] += A_tap[M, 64] * B_tap[N, 64]^T over 9 taps x Cin/64 chunks.
//   A operand   one 4-D TMA box {64 ch, 16 px, 18 rows, 1 image} per 64-channel chunk = the tile plus its 1-pixel halo, rows
//               pitched at 16 pixels (2048 B) so that every 8-pixel row segment is one 1024-byte SWIZZLE_128B atom group.
//               The nine taps are nine UMMA descriptors INTO THE SAME TILE: start = base + (dy*16 + dx)*128 B, stride between
//               8-row groups = 2048 B (the swizzle follows absolute smem address bits, so no base offset is needed).  Each input
//               element is fetched from L2 once per tile instead of nine times -- the first version of this kernel
//               (one TMA box per tap) ran the 64-channel full-resolution layers at the L2->SM bandwidth limit.
//               Out-of-bounds rows / columns arrive as zeros = the conv's zero padding and the F.pad of the up-sampled map;
//               the K range may span two tensors (skip, up): torch.cat is never materialised.
//   B operand   packed weights [N][tap][Cin] bf16; per (tap, chunk) one 2-D TMA box {64, BLOCK_N}.  When the whole
//               9*Cin x BLOCK_N panel fits beside the A ring (<= 147 KB) it is loaded ONCE per CTA and stays resident for
//               every tile; otherwise it streams through its own ring.
//   roles       warp 0: A producer - warp 1: tcgen05.mma issuer - warp 2: B producer + TMEM allocator - warps 3..6: epilogue
//               (tcgen05.ld -> fp32 BN scale/shift -> ReLU -> bf16 NHWC store | fused 2x2 max-pool | fused 1x1 head).
//               Two TMEM accumulators: the epilogue of tile i overlaps the MMAs of tile i+1.
#include "tc_common.cuh"

namespace adn {

constexpr int H_TW = 8, H_TH = 16;                 // pixel tile: 16 rows x 8 columns = 128 GEMM rows
constexpr int H_ROWS = H_TH + 2;
// halo row pitch in pixels: 10 = the dense 18 x 10 halo tile (the swizzle follows absolute address bits, so 8-pixel row groups may
// start at any 128-byte offset and SBO = 1280 B works), 16 = rows padded to two 1024-B atoms (the first version)
constexpr int h_a_bytes(int pitch) { return H_ROWS * pitch * 128; }
constexpr int h_a_stage(int pitch) { return (h_a_bytes(pitch) + 1023) & ~1023; }
constexpr int H_THREADS = 224;
constexpr int H_EPI_THREADS = 128;
constexpr int H_MAX_A = 8, H_MAX_B = 16;
constexpr int H_OUT_STAGE = 128 * 128;             // one 128-pixel x 64-channel bf16 output tile
constexpr int H_POOL_STAGE = 32 * 128;             // its 2x2-pooled counterpart

enum { HEPI_NHWC = 0, HEPI_HEAD = 2 };


struct HaloArgs {
    int c0_chunks, c1_chunks;
    int n_img, H, W;
    int tiles_x, tiles_y;
    int c_out, n_blocks, num_tiles;
    FastDiv div_nb, div_tpi, div_tx;   // by n_blocks, tiles_x * tiles_y, tiles_x: the per-tile index split costs 2 instructions
    int epi;
    int a_stages, b_slots, b_resident;
    float relu_floor;              // 0 = ReLU, -inf = no activation (train-mode pre-BN output, dgrad)
    int tma_store;                 // epilogue stages bf16 tiles in smem and stores them with TMA (when the smem budget allows)
    int warp_store;                // each epilogue warp stores its own 32-pixel x 64-channel box (no CTA-wide barrier in the epilogue)
    const float* scale;
    const float* shift;
    __nv_bfloat16* out;
    __nv_bfloat16* pool_out;       // optional fused MaxPool2d(2): (n, H/2, W/2, c_out)
    const float* head_w;
    const float* head_b;
    float* head_out;
};

template <int BLOCK_N, bool B_RESIDENT, int NCTA, int H_PITCH>
__global__ void __launch_bounds__(H_THREADS, 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                    const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOut,
                    const __grid_constant__ CUtensorMap tmPool, const HaloArgs a) {
    // NCTA == 2: a CTA pair (cluster of 2, cta_group::2) computes TWO pixel tiles with ONE 256-row UMMA per k-step.  Each CTA
    // loads its own halo tile and HALF of the weight block, so the shared-memory operand traffic per CTA and MMA drops from
    // A + B to A + B/2 bytes -- the single-CTA kernel was measured at the smem->tensor-core operand bandwidth (~64 B/clk).
    constexpr bool PAIR = (NCTA == 2);
    constexpr int H_A_STAGE = h_a_stage(H_PITCH), H_A_BYTES = h_a_bytes(H_PITCH);
    constexpr int B_ROWS = BLOCK_N / NCTA;                         // weight rows held by this CTA
    constexpr int B_BLOCK = B_ROWS * 128;                          // bytes of one (tap, chunk) weight block in this CTA
    constexpr int TMEM_COLS = (2 * BLOCK_N) < 32 ? 32 : 2 * BLOCK_N;
    const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
    const bool leader = (cta_rank == 0);
    extern __shared__ uint8_t smem_dyn[];
    const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_dyn + (smem_base - smem_u32(smem_dyn));

    const uint32_t a_base = smem_base;
    const uint32_t b_base = a_base + (uint32_t)a.a_stages * H_A_STAGE;
    // epilogue staging (TMA-store mode): 2 x [128 px][64 ch] bf16 (+ 2 x [32 pooled px][64 ch]), SWIZZLE_128B, 1024-B aligned
    const uint32_t stage_off = (uint32_t)a.a_stages * H_A_STAGE + (uint32_t)a.b_slots * B_BLOCK;
    const uint32_t stage_bytes = a.tma_store ? (2u * H_OUT_STAGE + (a.pool_out ? 2u * H_POOL_STAGE : 0u)) : 0u;
    const uint32_t aux_off = stage_off + stage_bytes;
    float* s_scale = reinterpret_cast<float*>(smem_gen + aux_off);     // [c_out] all output channels, loaded once per CTA
    float* s_shift = s_scale + a.c_out;
    float* s_head = s_shift + a.c_out;
    const uint32_t aux_f32 = (uint32_t)(2 * a.c_out + 64) * 4;
    const uint32_t bar_base = smem_base + aux_off + aux_f32;
    auto full_a = [&](int s) { return bar_base + 8u * s; };
    auto empty_a = [&](int s) { return bar_base + 8u * (H_MAX_A + s); };
    auto full_b = [&](int s) { return bar_base + 8u * (2 * H_MAX_A + s); };
    auto empty_b = [&](int s) { return bar_base + 8u * (2 * H_MAX_A + H_MAX_B + s); };
    auto tfull = [&](int s) { return bar_base + 8u * (2 * H_MAX_A + 2 * H_MAX_B + s); };
    auto tempty = [&](int s) { return bar_base + 8u * (2 * H_MAX_A + 2 * H_MAX_B + 2 + s); };
    const uint32_t bres = bar_base + 8u * (2 * H_MAX_A + 2 * H_MAX_B + 4);
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem_gen + aux_off + aux_f32 + (2 * H_MAX_A + 2 * H_MAX_B + 5) * 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA0);
        tma_prefetch_desc(&tmA1);
        tma_prefetch_desc(&tmB);
        if (a.tma_store) { tma_prefetch_desc(&tmOut); if (a.pool_out) tma_prefetch_desc(&tmPool); }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < H_MAX_A; ++s) { mbar_init(full_a(s), 1); mbar_init(empty_a(s), 1); }
        for (int s = 0; s < H_MAX_B; ++s) { mbar_init(full_b(s), 1); mbar_init(empty_b(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull(s), 1); mbar_init(tempty(s), NCTA * H_EPI_THREADS / 32); }
        mbar_init(bres, 1);
        fence_barrier_init();
    }
    if (warp == 2) {
        if (PAIR) { tmem_alloc_pair(smem_u32(tmem_ptr_smem), TMEM_COLS); tmem_relinquish_pair(); }
        else { tmem_alloc(smem_u32(tmem_ptr_smem), TMEM_COLS); tmem_relinquish(); }
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();           // peers must see initialised barriers before any remote arrive
    tc_fence_after();
    // a pair works out of ONE TMEM address in both CTAs: the leader's UMMA writes D at the leader's allocation in each CTA, so
    // the peer takes the leader's base (its own cta_group::2 allocation is only freed at the end)
    const uint32_t tmem_own = *tmem_ptr_smem;
    const uint32_t tmem_base = PAIR ? ld_shared_cluster_u32(mapa_shared(smem_u32(tmem_ptr_smem), 0)) : tmem_own;

    const int chunks = a.c0_chunks + a.c1_chunks;
    const int tiles_per_img = a.tiles_x * a.tiles_y;
    // work items: single CTA -> (m tile, n block); pair -> (pair of m tiles, n block), this CTA taking m = 2*pair + rank
    const int num_m = a.n_img * tiles_per_img;
    const int work_total = PAIR ? ((num_m + 1) / 2) * a.n_blocks : a.num_tiles;
    const int work_first = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int work_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    auto tile_m = [&](int w) { return PAIR ? 2 * fast_div(w, a.div_nb) + (int)cta_rank : fast_div(w, a.div_nb); };
    auto tile_nblk = [&](int w) { return w - fast_div(w, a.div_nb) * a.n_blocks; };
    // full barriers live in the leader CTA; a peer's TMA completes its bytes there
    auto full_a_sig = [&](int s) { return PAIR ? mapa_shared(full_a(s), 0) : full_a(s); };
    auto full_b_sig = [&](int s) { return PAIR ? mapa_shared(full_b(s), 0) : full_b(s); };

    if (warp == 0) {
        // ===================================================================== A producer: one halo tile per (tile, chunk)
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int w = work_first; w < work_total; w += work_step) {
                const int m = tile_m(w);
                const int img = fast_div(m, a.div_tpi);           // m >= num_m (odd tail of a pair): img == n_img -> all zeros
                const int rem = m - img * tiles_per_img;
                const int ty = fast_div(rem, a.div_tx), tx = rem - ty * a.tiles_x;
                const int x0 = tx * H_TW - 1, y0 = ty * H_TH - 1;
                for (int ch = 0; ch < chunks; ++ch) {
                    mbar_wait(empty_a(stage), phase ^ 1u);
                    if (leader) mbar_arrive_expect_tx(full_a(stage), NCTA * H_A_BYTES);
                    const uint32_t dst = a_base + (uint32_t)stage * H_A_STAGE;
                    const CUtensorMap* map = (ch < a.c0_chunks) ? &tmA0 : &tmA1;
                    const int c = (ch < a.c0_chunks ? ch : ch - a.c0_chunks) * 64;
                    if (PAIR) tma_load_4d_pair(dst, map, full_a_sig(stage), c, x0, y0, img);
                    else tma_load_4d(dst, map, full_a(stage), c, x0, y0, img);
                    if (++stage == a.a_stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 2) {
        // ===================================================================== B producer (this CTA's B_ROWS of every block)
        if (lane == 0) {
            const int row0 = (int)cta_rank * B_ROWS;
            if (B_RESIDENT) {
                if (leader) mbar_arrive_expect_tx(bres, (uint32_t)(NCTA * 9 * chunks) * B_BLOCK);
                const uint32_t sig = PAIR ? mapa_shared(bres, 0) : bres;
                for (int ch = 0; ch < chunks; ++ch)
                    for (int tap = 0; tap < 9; ++tap) {
                        const uint32_t dst = b_base + (uint32_t)(ch * 9 + tap) * B_BLOCK;
                        if (PAIR) tma_load_2d_pair(dst, &tmB, sig, (tap * chunks + ch) * 64, row0);
                        else tma_load_2d(dst, &tmB, sig, (tap * chunks + ch) * 64, row0);
                    }
            } else {
                int slot = 0; uint32_t phase = 0;
                for (int w = work_first; w < work_total; w += work_step) {
                    const int n_blk = tile_nblk(w);
                    for (int ch = 0; ch < chunks; ++ch)
                        for (int tap = 0; tap < 9; ++tap) {
                            mbar_wait(empty_b(slot), phase ^ 1u);
                            if (leader) mbar_arrive_expect_tx(full_b(slot), NCTA * B_BLOCK);
                            const uint32_t dst = b_base + (uint32_t)slot * B_BLOCK;
                            if (PAIR) tma_load_2d_pair(dst, &tmB, full_b_sig(slot), (tap * chunks + ch) * 64, n_blk * BLOCK_N + row0);
                            else tma_load_2d(dst, &tmB, full_b(slot), (tap * chunks + ch) * 64, n_blk * BLOCK_N + row0);
                            if (++slot == a.b_slots) { slot = 0; phase ^= 1u; }
                        }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer (whole warp, one elected lane issues;
        // in a CTA pair only the leader's warp issues, for both CTAs)
        // Descriptor arithmetic is hoisted: per stage one base descriptor, per tap / k-step a compile-time constant is added
        // to the 14-bit start-address field (never carries: shared memory is < 256 KB).
        if (leader) {
            constexpr uint32_t idesc = make_idesc(BLOCK_N, 128 * NCTA);
            constexpr uint64_t B_STEP = (uint64_t)(B_BLOCK >> 4);
            auto mma = [&](uint32_t d, uint64_t da, uint64_t db, uint32_t accf) {
                if (PAIR) umma_bf16_pair(d, da, db, idesc, accf); else umma_bf16(d, da, db, idesc, accf);
            };
            auto commit = [&](uint32_t bar) { if (PAIR) umma_commit_pair(bar); else umma_commit(bar); };
            int sa = 0; uint32_t pa = 0;
            int sb = 0; uint32_t pb = 0;
            int acc = 0; uint32_t acc_phase = 0;
            if (B_RESIDENT) { mbar_wait(bres, 0); tc_fence_after(); }
            const uint64_t db_base = make_sw128_desc(b_base);
            for (int w = work_first; w < work_total; w += work_step) {
                mbar_wait(tempty(acc), acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
                for (int ch = 0; ch < chunks; ++ch) {
                    mbar_wait(full_a(sa), pa);
                    tc_fence_after();
                    // measured on B200: the UMMA swizzle is a function of the absolute shared-memory address bits (like TMA's),
                    // so a tap view that starts dx rows into a 1024-byte atom needs NO descriptor base offset
                    const uint64_t da_stage = make_sw128_desc(a_base + (uint32_t)sa * H_A_STAGE, H_PITCH * 128);
                    if (B_RESIDENT) {
                        const uint64_t db_chunk = db_base + (uint64_t)(ch * 9) * B_STEP;
                        if (elect_one()) {
#pragma unroll
                            for (int tap = 0; tap < 9; ++tap) {
                                const uint64_t da = da_stage + (uint64_t)(((tap / 3) * H_PITCH + (tap % 3)) * 8);
                                const uint64_t db = db_chunk + (uint64_t)tap * B_STEP;
#pragma unroll
                                for (int k = 0; k < 4; ++k)              // UMMA_K = 16: +32 bytes inside the swizzle row
                                    mma(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), (tap | k) != 0 ? 1u : (ch != 0 ? 1u : 0u));
                            }
                            commit(empty_a(sa));
                        }
                        __syncwarp();
                    } else {
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap) {
                            mbar_wait(full_b(sb), pb);
                            tc_fence_after();
                            const uint64_t da = da_stage + (uint64_t)(((tap / 3) * H_PITCH + (tap % 3)) * 8);
                            const uint64_t db = db_base + (uint64_t)sb * B_STEP;
                            if (elect_one()) {
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    mma(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), (tap | k) != 0 ? 1u : (ch != 0 ? 1u : 0u));
                                commit(empty_b(sb));
                                if (tap == 8) commit(empty_a(sa));
                            }
                            __syncwarp();
                            if (++sb == a.b_slots) { sb = 0; pb ^= 1u; }
                        }
                    }
                    if (++sa == a.a_stages) { sa = 0; pa ^= 1u; }
                }
                if (elect_one()) commit(tfull(acc));
                __syncwarp();
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else {
        // ===================================================================== epilogue (warps 3..6 = TMEM lane quadrants 3,0,1,2)
        const int quad = warp & 3;
        const int row = quad * 32 + lane;                            // accumulator row = pixel within the tile
        const int et = threadIdx.x - 96;                             // 0..127
        const int lx = row & (H_TW - 1), ly = row >> 3;
        if (a.epi == HEPI_HEAD && et < 64) s_head[et] = a.head_w[et];
        for (int c = et; c < a.c_out; c += H_EPI_THREADS) { s_scale[c] = a.scale[c]; s_shift[c] = a.shift[c]; }
        named_bar_sync(1, H_EPI_THREADS);
        const int Hp = a.H >> 1, Wp = a.W >> 1;
        int acc = 0; uint32_t acc_phase = 0;
        uint32_t store_groups = 0;
        const uint32_t tempty_sig0 = PAIR ? mapa_shared(tempty(0), 0) : tempty(0);
        const uint32_t tempty_sig1 = PAIR ? mapa_shared(tempty(1), 0) : tempty(1);
        for (int w = work_first; w < work_total; w += work_step) {
            const int n_blk = tile_nblk(w);
            const int m = tile_m(w);
            const int img = fast_div(m, a.div_tpi);                 // == n_img for the padding tile of an odd pair
            const int rem = m - img * tiles_per_img;
            const int ty = fast_div(rem, a.div_tx), tx = rem - ty * a.tiles_x;
            const int x = tx * H_TW + lx, y = ty * H_TH + ly;
            const bool valid = (x < a.W) && (y < a.H) && (m < num_m);
            // the pooled pixel this lane contributes to (its 2x2 window = lanes l, l^1, l^8, l^9 of the same warp) exists
            const bool pool_px = a.pool_out && (m < num_m) && (x >> 1) < Wp && (y >> 1) < Hp;

            const float* t_scale = s_scale + n_blk * BLOCK_N;
            const float* t_shift = s_shift + n_blk * BLOCK_N;

            mbar_wait(tfull(acc), acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BLOCK_N);
            float2 head_acc2 = make_float2(0.f, 0.f);                // even / odd channel partial sums of the 1x1 head
            // one 32-column group of the accumulator row: affine (+ ReLU) -> bf16 -> staging / global (+ pool | head)
            auto group = [&](const int c0, const bool first_half, const uint32_t (&r)[32]) {
                const float4* sc4 = reinterpret_cast<const float4*>(t_scale + c0);
                const float4* sh4 = reinterpret_cast<const float4*>(t_shift + c0);
                if (a.epi == HEPI_HEAD) {
                    const float4* hw4 = reinterpret_cast<const float4*>(s_head + c0);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 sc = sc4[i], sh = sh4[i], hw = hw4[i];
                        float2 y0 = pk_fma(make_float2(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1])), make_float2(sc.x, sc.y), make_float2(sh.x, sh.y));
                        float2 y1 = pk_fma(make_float2(__uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3])), make_float2(sc.z, sc.w), make_float2(sh.z, sh.w));
                        y0.x = fmaxf(y0.x, a.relu_floor); y0.y = fmaxf(y0.y, a.relu_floor);
                        y1.x = fmaxf(y1.x, a.relu_floor); y1.y = fmaxf(y1.y, a.relu_floor);
                        head_acc2 = pk_fma(y0, make_float2(hw.x, hw.y), head_acc2);
                        head_acc2 = pk_fma(y1, make_float2(hw.z, hw.w), head_acc2);
                    }
                    return;
                }
                // packed fp32x2 affine, ReLU folded into the bf16x2 conversion (relu_floor is 0 or -inf)
                uint32_t pk[16];
                if (a.relu_floor == 0.f) {                            // warp-uniform
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 sc = sc4[i], sh = sh4[i];
                        const float2 y0 = pk_fma(make_float2(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1])), make_float2(sc.x, sc.y), make_float2(sh.x, sh.y));
                        const float2 y1 = pk_fma(make_float2(__uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3])), make_float2(sc.z, sc.w), make_float2(sh.z, sh.w));
                        pk[2 * i] = pack_relu_bf16x2(y0.x, y0.y); pk[2 * i + 1] = pack_relu_bf16x2(y1.x, y1.y);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 sc = sc4[i], sh = sh4[i];
                        const float2 y0 = pk_fma(make_float2(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1])), make_float2(sc.x, sc.y), make_float2(sh.x, sh.y));
                        const float2 y1 = pk_fma(make_float2(__uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3])), make_float2(sc.z, sc.w), make_float2(sh.z, sh.w));
                        pk[2 * i] = pack_bf16x2(y0.x, y0.y); pk[2 * i + 1] = pack_bf16x2(y1.x, y1.y);
                    }
                }
                const int n = n_blk * BLOCK_N + c0;
                const uint32_t cbase = first_half ? 0u : 4u;
                uint32_t o_stage = 0, p_stage = 0;
                if (a.tma_store) {                                    // warp-uniform
                    const uint32_t buf = store_groups & 1u;
                    o_stage = smem_base + stage_off + buf * H_OUT_STAGE;
                    p_stage = smem_base + stage_off + 2u * H_OUT_STAGE + buf * H_POOL_STAGE;
                    if (first_half) {                                 // this buffer's previous TMA store must have read it
                        if (a.warp_store) { if (lane == 0) bulk_wait_read<1>(); __syncwarp(); }
                        else { if (et == 0) bulk_wait_read<1>(); named_bar_sync(1, H_EPI_THREADS); }
                    }
                    const uint32_t rbase = o_stage + (uint32_t)row * 128u;
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        st_shared_v4(rbase + (((cbase + i) ^ ((uint32_t)row & 7u)) << 4), pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
                } else if (valid) {
                    uint4* d4 = reinterpret_cast<uint4*>(a.out + (((long long)img * a.H + y) * a.W + x) * a.c_out + n);
#pragma unroll
                    for (int i = 0; i < 4; ++i) d4[i] = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
                }
                if (a.pool_out) {                                     // warp-uniform
                    // 2x2 max over the lane window {l, l^1, l^8, l^9} by halving exchanges: each lane sends the half of its
                    // registers the partner keeps, so 12 shuffles (not 32) leave every lane with ONE 16-byte piece (8 channels)
                    // of the pooled pixel: piece index = 2*(lx&1) + (ly&1).  Values are post-ReLU or plain maxima either way;
                    // windows hanging over the image edge only feed pooled pixels that are clipped / not written (floor pooling).
                    const bool ox = lx & 1, oy = ly & 1;
                    uint32_t m1[8], m2[4];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const uint32_t keep = ox ? pk[i + 8] : pk[i], send = ox ? pk[i] : pk[i + 8];
                        m1[i] = bf16x2_max(keep, __shfl_xor_sync(0xffffffffu, send, 1));
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint32_t keep = oy ? m1[i + 4] : m1[i], send = oy ? m1[i] : m1[i + 4];
                        m2[i] = bf16x2_max(keep, __shfl_xor_sync(0xffffffffu, send, 8));
                    }
                    const uint32_t piece = (ox ? 2u : 0u) + (oy ? 1u : 0u);
                    if (a.tma_store) {
                        const uint32_t pr = (uint32_t)((ly >> 1) * 4 + (lx >> 1));
                        st_shared_v4(p_stage + pr * 128u + (((cbase + piece) ^ (pr & 7u)) << 4), m2[0], m2[1], m2[2], m2[3]);
                    } else if (pool_px) {
                        *reinterpret_cast<uint4*>(a.pool_out + (((long long)img * Hp + (y >> 1)) * Wp + (x >> 1)) * a.c_out + n + piece * 8) =
                            make_uint4(m2[0], m2[1], m2[2], m2[3]);
                    }
                }
                if (a.tma_store && !first_half) {                     // a 64-channel group is staged: hand it to the TMA engine
                    fence_proxy_async();
                    const int ch0 = n_blk * BLOCK_N + (c0 - 32);
                    if (a.warp_store) {                               // this warp's 4 tile rows (32 pixels) / 2 pooled rows: its own boxes
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_4d(&tmOut, o_stage + (uint32_t)quad * 4096u, ch0, tx * H_TW, ty * H_TH + quad * 4, img);
                            if (a.pool_out) tma_store_4d(&tmPool, p_stage + (uint32_t)quad * 1024u, ch0, tx * (H_TW / 2), ty * (H_TH / 2) + quad * 2, img);
                            bulk_commit();
                        }
                    } else {
                        named_bar_sync(1, H_EPI_THREADS);
                        if (et == 0) {
                            tma_store_4d(&tmOut, o_stage, ch0, tx * H_TW, ty * H_TH, img);
                            if (a.pool_out) tma_store_4d(&tmPool, p_stage, ch0, tx * (H_TW / 2), ty * (H_TH / 2), img);
                            bulk_commit();
                        }
                    }
                    ++store_groups;
                }
            };
            // the TMEM load of the next 32 columns is in flight while the current group is processed
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
            if (a.epi == HEPI_HEAD && valid)
                a.head_out[((long long)img * a.H + y) * a.W + x] = (head_acc2.x + head_acc2.y) + a.head_b[0];
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {                                         // one arrival per epilogue warp (of both CTAs) frees the accumulator
                if (PAIR) mbar_arrive_cluster(acc ? tempty_sig1 : tempty_sig0); else mbar_arrive(tempty(acc));
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
        if (a.tma_store && (a.warp_store ? lane == 0 : et == 0)) bulk_wait<0>();     // smem must outlive the last bulk stores
    }

    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();           // the peer may still read this CTA's smem / write its TMEM
    tc_fence_after();
    if (warp == 2) { if (PAIR) tmem_dealloc_pair(tmem_own, TMEM_COLS); else tmem_dealloc(tmem_own, TMEM_COLS); }
}

// ------------------------------------------------------------------------------------------------ host side
static int g_dx_mode = 1;        // 0: never use the kx-in-N kernel of conv_dx.cu for the 64-output-channel layers
static int g_pair_mode = 3;      // 0: never use CTA pairs; 1: where the layer is shared-memory-operand bound and a pair pays off; 3: also the 256-wide layers
static int g_halo_pitch = 10;    // 10: dense halo tile; 16: padded rows
static int g_halo_stages = 0;    // > 0: A stages of the streaming-B configuration (tuning hook)
// tuning hook bits: 1 = CTA pairs also for the one-chunk 128-wide layer (measured: downconv2.0 0.63 -> 0.87 ms, off);
// 2 = stream the weights when keeping them resident would cost the TMA-store epilogue its staging buffers (measured at batch 64:
// downconv2.3 1.15 -> 1.02 ms, downconv2.0 0.63 -> 0.60, upconv3.3 0.92 -> 0.90; on);
// 4 = every epilogue warp stores its own 32-pixel box, no CTA-wide barrier in the epilogue (measured: downconv2.3 1.00 -> 0.96 ms,
// every other layer within noise, forward total unchanged: the epilogue barriers are not what holds the 128-wide layers back; off)
static int g_halo_tune = 2;

template <int BLOCK_N, int NCTA, int H_PITCH>
static int launch_halo(const CUtensorMap& mA0, const CUtensorMap& mA1, const CUtensorMap& mB, const CUtensorMap& mOut,
                       const CUtensorMap& mPool, HaloArgs& args, int chunks, cudaStream_t stream) {
    constexpr int B_BLOCK = (BLOCK_N / NCTA) * 128;              // bytes of one (tap, chunk) weight block in ONE CTA
    constexpr int H_A_STAGE = h_a_stage(H_PITCH);
    const int AUX = (2 * args.c_out + 64) * 4 + (2 * H_MAX_A + 2 * H_MAX_B + 5) * 8 + 16;
    constexpr int MAX_DYN = 232448;
    const int budget = MAX_DYN - 1024 - AUX;
    const int total_b = 9 * chunks * B_BLOCK;
    // TMA-store staging is worth its shared memory only where it does not starve the operand rings
    const int staging = (args.epi == HEPI_NHWC) ? 2 * H_OUT_STAGE + (args.pool_out ? 2 * H_POOL_STAGE : 0) : 0;
    args.tma_store = 0;
    bool resident = args.n_blocks == 1 && total_b + 2 * H_A_STAGE <= budget;
    if (resident && (g_halo_tune & 2) && staging && (budget - total_b) - staging < 3 * H_A_STAGE) resident = false;
    if (resident) {
        args.b_resident = 1;
        args.b_slots = 9 * chunks;
        int avail = budget - total_b;
        if (staging && avail - staging >= 3 * H_A_STAGE) { args.tma_store = 1; avail -= staging; }
        const int st = avail / H_A_STAGE;
        args.a_stages = st > H_MAX_A ? H_MAX_A : st;
    } else {
        args.b_resident = 0;
        args.a_stages = g_halo_stages > 0 ? g_halo_stages : (BLOCK_N >= 256) ? 2 : 3;
        int avail = budget - args.a_stages * H_A_STAGE;
        if (staging && (avail - staging) / B_BLOCK >= 4) { args.tma_store = 1; avail -= staging; }
        const int sl = avail / B_BLOCK;
        args.b_slots = sl > H_MAX_B ? H_MAX_B : sl;
        if (args.b_slots < 2) return ADN_ERR_ARG;
    }
    const int smem = 1024 + args.a_stages * H_A_STAGE + args.b_slots * B_BLOCK + (args.tma_store ? staging : 0) + AUX;
    const int sms = num_sms();
    int grid;
    if (NCTA == 2) {
        const int num_m = args.n_img * args.tiles_x * args.tiles_y;
        const int pairs = ((num_m + 1) / 2) * args.n_blocks;
        const int max_pairs = sms / 2;
        grid = 2 * (pairs < max_pairs ? pairs : max_pairs);
    } else {
        grid = args.num_tiles < sms ? args.num_tiles : sms;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid, 1, 1);
    cfg.blockDim = dim3(H_THREADS, 1, 1);
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = NCTA; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (NCTA == 2) ? 1 : 0;
    if (args.b_resident) {
        static unsigned char smem_set[64] = {0};
        ADN_CUDA_TRY(ensure_dyn_smem(conv3x3_halo_kernel<BLOCK_N, true, NCTA, H_PITCH>, MAX_DYN, smem_set));
        ADN_CUDA_TRY(cudaLaunchKernelEx(&cfg, conv3x3_halo_kernel<BLOCK_N, true, NCTA, H_PITCH>, mA0, mA1, mB, mOut, mPool, args));
    } else {
        static unsigned char smem_set[64] = {0};
        ADN_CUDA_TRY(ensure_dyn_smem(conv3x3_halo_kernel<BLOCK_N, false, NCTA, H_PITCH>, MAX_DYN, smem_set));
        ADN_CUDA_TRY(cudaLaunchKernelEx(&cfg, conv3x3_halo_kernel<BLOCK_N, false, NCTA, H_PITCH>, mA0, mA1, mB, mOut, mPool, args));
    }
    ADN_LAUNCH_CHECK();
    return ADN_OK;
}

static int conv3x3_halo(const void* src0, int c0, const void* src1, int c1, int h1, int w1, int n, int h, int w,
                        const void* w_packed, int c_out, const float* scale, const float* shift, int epi, void* out, void* pool_out,
                        const float* head_w, const float* head_b, float* head_out, cudaStream_t stream, int relu = 1) {
    if (!src0 || !w_packed || !scale || !shift || n <= 0 || h <= 0 || w <= 0) return ADN_ERR_ARG;
    if (c0 <= 0 || (c0 % 64) || c1 < 0 || (c1 % 64) || (c1 > 0 && !src1)) return ADN_ERR_ARG;
    if (c_out <= 0 || (c_out % 64)) return ADN_ERR_ARG;
    if (c1 > 0 && (h1 > h || w1 > w || h1 <= 0 || w1 <= 0)) return ADN_ERR_ARG;
    if (!aligned16(src0) || !aligned16(w_packed) || (src1 && !aligned16(src1))) return ADN_ERR_ARG;
    if (epi == HEPI_HEAD) { if (!head_w || !head_b || !head_out || c_out != 64) return ADN_ERR_ARG; }
    else if (!out || !aligned16(out) || (pool_out && !aligned16(pool_out))) return ADN_ERR_ARG;
    if (pool_out && (h < 2 || w < 2)) return ADN_ERR_ARG;
    int st = check_device();
    if (st != ADN_OK) return st;

    // 64 output channels over <= 128 input channels: the kx-in-N formulation (conv_dx.cu) reads a third of the A bytes per MAC
    if (g_dx_mode && conv3x3_dx_eligible(c0, c1, c_out))
        return conv3x3_dx(src0, c0, src1, c1, h1, w1, n, h, w, w_packed, scale, shift, relu ? 0.f : -INFINITY, out, pool_out,
                          head_w, head_b, epi == HEPI_HEAD ? head_out : nullptr, stream);

    const int block_n = (c_out % 256 == 0) ? 256 : (c_out % 128 == 0) ? 128 : 64;
    HaloArgs args;
    args.c0_chunks = c0 / 64; args.c1_chunks = c1 / 64;
    args.n_img = n; args.H = h; args.W = w;
    args.tiles_x = (w + H_TW - 1) / H_TW; args.tiles_y = (h + H_TH - 1) / H_TH;
    args.c_out = c_out; args.n_blocks = c_out / block_n;
    args.div_nb = make_fastdiv(args.n_blocks); args.div_tpi = make_fastdiv(args.tiles_x * args.tiles_y); args.div_tx = make_fastdiv(args.tiles_x);
    const long long tiles = (long long)n * args.tiles_x * args.tiles_y * args.n_blocks;
    if (tiles > 0x7fffffffLL) return ADN_ERR_ARG;
    args.num_tiles = (int)tiles;
    args.epi = epi;
    args.relu_floor = relu ? 0.f : -INFINITY;
    args.scale = scale; args.shift = shift;
    args.out = (__nv_bfloat16*)out; args.pool_out = (__nv_bfloat16*)pool_out;
    args.head_w = head_w; args.head_b = head_b; args.head_out = head_out;

    CUtensorMap mA0, mA1, mB;
    const int pitch = g_halo_pitch;
    st = make_act_map(&mA0, src0, n, h, w, c0, pitch, H_ROWS);
    if (st != ADN_OK) return st;
    if (c1 > 0) st = make_act_map(&mA1, src1, n, h1, w1, c1, pitch, H_ROWS); else mA1 = mA0;
    if (st != ADN_OK) return st;
    // CTA pairs (cta_group::2): two pixel tiles per 256-row UMMA, each CTA holding half of the weight rows.  Used where one CTA is
    // bound by the shared-memory operand reads (BLOCK_N 64 / 128: A + B bytes per MMA exceed 128 B/clk), and only when every
    // SM pair still gets work
    const int num_m = n * args.tiles_x * args.tiles_y;
    // measured (variant B, batch 64): 128-wide layers with >= 2 input chunks gain 8-32 % (upconv3.0: 1 171 -> 1 548 TFLOP/s);
    // 64-wide layers LOSE 30 % in pair mode (a 256x64 UMMA is too short to amortise), so they stay single-CTA; so does the 128-wide
    // layer with ONE input chunk (downconv2.0, K = 576: 0.65 -> 0.89 ms as a pair)
    const bool pair = g_pair_mode && num_m >= 2 * (num_sms() / 2) &&
                      ((block_n == 128 && ((c0 + c1) >= 128 || (g_halo_tune & 1)) && args.n_blocks == 1) || (block_n == 256 && (g_pair_mode & 2)));
    st = make_weight_map(&mB, w_packed, c_out, 9 * (c0 + c1), pair ? block_n / 2 : block_n);
    if (st != ADN_OK) return st;

    // output maps for the TMA-store epilogue: box {64 ch, 8 px, 16 rows} of the NHWC output, {64, 4, 8} of the pooled one
    CUtensorMap mOut = mA0, mPool = mA0;
    args.warp_store = (g_halo_tune & 4) ? 1 : 0;
    if (epi == HEPI_NHWC) {
        const int div = args.warp_store ? 4 : 1;                    // per-warp boxes: a quarter of the tile's rows
        st = make_act_map(&mOut, out, n, h, w, c_out, H_TW, H_TH / div);
        if (st != ADN_OK) return st;
        if (pool_out) st = make_act_map(&mPool, pool_out, n, h / 2, w / 2, c_out, H_TW / 2, H_TH / 2 / div);
        if (st != ADN_OK) return st;
    }

    const int chunks = args.c0_chunks + args.c1_chunks;
#define ADN_HALO_LAUNCH(BN, NC) (pitch == 10 ? launch_halo<BN, NC, 10>(mA0, mA1, mB, mOut, mPool, args, chunks, stream) \
                                             : launch_halo<BN, NC, 16>(mA0, mA1, mB, mOut, mPool, args, chunks, stream))
    switch (block_n) {
        case 256: return pair ? ADN_HALO_LAUNCH(256, 2) : ADN_HALO_LAUNCH(256, 1);
        case 128: return pair ? ADN_HALO_LAUNCH(128, 2) : ADN_HALO_LAUNCH(128, 1);
        default: return ADN_HALO_LAUNCH(64, 1);
    }
#undef ADN_HALO_LAUNCH
}

}  // namespace adn

// tuning / debugging hook (not in the public header): 0 disables the CTA-pair kernels
extern "C" void adn__conv_pair_mode(int mode) { adn::g_pair_mode = mode; }
extern "C" void adn__conv_halo_tune(int bits) { adn::g_halo_tune = bits; }
extern "C" void adn__conv_halo_stages(int stages) { adn::g_halo_stages = stages; }
extern "C" void adn__conv_halo_pitch(int pitch) { adn::g_halo_pitch = pitch == 16 ? 16 : 10; }
// 0: conv_halo kernels only; 1 (default state): conv_dx for the 64-output-channel layers, single CTA, epilogue warp sets chosen per
// layer; 2: conv_dx as CTA pairs; 3 / 4: single CTA with four / two epilogue warp sets forced
extern "C" void adn__conv_dx_mode(int mode) {
    adn::g_dx_mode = mode != 0; adn::g_dx_pair = mode == 2; adn::g_dx_sets = mode == 3 ? 4 : mode == 4 ? 2 : 0;
}

extern "C" int adn_conv3x3_bn_relu_bf16(const void* src0, int c0, const void* src1, int c1, int h1, int w1, int n, int h, int w,
                                        const void* w_packed, int c_out, const float* scale, const float* shift, void* out,
                                        void* pool_out, void* stream) {
    return adn::conv3x3_halo(src0, c0, src1, c1, h1, w1, n, h, w, w_packed, c_out, scale, shift, adn::HEPI_NHWC, out, pool_out,
                             nullptr, nullptr, nullptr, (cudaStream_t)stream);
}

// Same kernel without the activation: out = conv * scale + shift (train-mode pre-BatchNorm output with scale = 1, shift = conv
// bias; the data-gradient conv of the backward pass with flipped weights, scale = 1, shift = 0).
extern "C" int adn_conv3x3_affine_bf16(const void* src0, int c0, const void* src1, int c1, int h1, int w1, int n, int h, int w,
                                       const void* w_packed, int c_out, const float* scale, const float* shift, int relu, void* out,
                                       void* stream) {
    return adn::conv3x3_halo(src0, c0, src1, c1, h1, w1, n, h, w, w_packed, c_out, scale, shift, adn::HEPI_NHWC, out, nullptr,
                             nullptr, nullptr, nullptr, (cudaStream_t)stream, relu);
}

extern "C" int adn_conv3x3_bn_relu_head_f32(const void* src0, int c0, const void* src1, int c1, int h1, int w1, int n, int h, int w,
                                            const void* w_packed, int c_out, const float* scale, const float* shift,
                                            const float* head_w, const float* head_b, float* out_f32, void* stream) {
    if (c_out != 64) return ADN_ERR_ARG;
    return adn::conv3x3_halo(src0, c0, src1, c1, h1, w1, n, h, w, w_packed, c_out, scale, shift, adn::HEPI_HEAD, nullptr, nullptr,
                             head_w, head_b, out_f32, (cudaStream_t)stream);
}
