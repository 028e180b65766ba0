// conv_tc.cu -- tap-streaming tcgen05/TMEM implicit-GEMM kernel for sm_100a.  In the product it runs the reference UNet's
// ConvTranspose2d 2x2/2 (code/model.py:38,43: one tap, GEMM N = 4*Cout with a pixel-shuffle store); the 3x3 convolutions
// moved to conv_halo.cu, which keeps the halo tile resident in shared memory instead of re-fetching it per tap.
//
//   GEMM view     D[M = n*h*w pixels, N = c_out] = A[M, K] * B[N, K]^T,  K = taps * c_in (taps = 9 or 1), bf16 x bf16 -> fp32
//   A operand     never materialised: for every (tap, 64-channel chunk) one 4-D TMA box {64 ch, TW, TH, 1 image} of the
//                 NHWC activation, shifted by the tap offset; out-of-bounds rows/columns arrive as zeros, which is the
//                 conv's zero padding AND the F.pad of the up-sampled tensor (model.py:44-47).  The K range may be split
//                 over two tensors (skip, up) -- torch.cat([x2, x1]) of model.py:49 is never built.
//                 TH*TW = 128 pixels = the 128 TMEM lanes of one accumulator; the box lands as a K-major SWIZZLE_128B
//                 tile, exactly what the UMMA descriptor expects.
//   B operand     packed weights [N][K] bf16 (K-major), 2-D TMA box {64, BLOCK_N}.
//   roles         warp 0: TMA producer (1 elected lane) - warp 1: tcgen05.mma issuer (1 lane) - warps 2..5: epilogue
//                 (tcgen05.ld -> fp32 BN scale/shift -> ReLU -> bf16 -> 16-byte global stores).  smem ring of STAGES
//                 {A,B} buffers with full/empty mbarriers; two TMEM accumulators so the epilogue of tile i overlaps the
//                 MMAs of tile i+1.
//   epilogue      + bias, bf16, pixel-shuffle store: GEMM column n = q*Cout + co of input pixel (y,x) -> output (2y+dy, 2x+dx, co).
#include "tc_common.cuh"

namespace adn {

// ------------------------------------------------------------------------------------------------ kernel
constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;                 // bf16 elements = one 128-byte swizzle row
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int CONV_THREADS = 192;
constexpr int EPI_THREADS = 128;


struct ConvArgs {
    int c0_chunks, c1_chunks;      // 64-channel chunks taken from source 0 / source 1
    int taps;                      // 9 (3x3, pad 1) or 1
    int n_img, H, W;               // GEMM rows = n_img*H*W pixels of the A tensors' grid
    int tiles_x, tiles_y, tw_log2; // pixel tile = TH x TW, TW = 1 << tw_log2, TH = 128 >> tw_log2
    int n_total;                   // GEMM N (c_out, or 4*c_out for the transposed conv)
    int c_out;                     // channels of the output tensor
    int n_blocks;                  // n_total / BLOCK_N
    int num_tiles;                 // n_img*tiles_y*tiles_x*n_blocks
    const float* shift;            // bias[c_out]
    __nv_bfloat16* out;
};

template <int BLOCK_N>
struct ConvCfg {
    static constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int STAGES = (196608 / STAGE_BYTES) > 8 ? 8 : (196608 / STAGE_BYTES);
    static constexpr int TMEM_COLS = (2 * BLOCK_N) < 32 ? 32 : 2 * BLOCK_N;     // 128 / 256 / 512: powers of two
    static constexpr int AUX_BYTES = 2 * BLOCK_N * 4 + 64 * 4 + (2 * STAGES + 4) * 8 + 16;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + AUX_BYTES + 1024;  // +1024: manual 1024-byte alignment
};

template <int BLOCK_N>
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmB, const ConvArgs a) {
    using Cfg = ConvCfg<BLOCK_N>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_dyn[];
    const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_dyn + (smem_base - smem_u32(smem_dyn));

    // carve-up: [STAGES x {A, B}] [scale BLOCK_N] [shift BLOCK_N] [head_w 64] [full STAGES] [empty STAGES] [tfull 2] [tempty 2] [tmem ptr]
    float* s_scale = reinterpret_cast<float*>(smem_gen + STAGES * Cfg::STAGE_BYTES);
    float* s_shift = s_scale + BLOCK_N;
    const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES + (2 * BLOCK_N + 64) * 4;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem_gen + STAGES * Cfg::STAGE_BYTES + (2 * BLOCK_N + 64) * 4 + (2 * STAGES + 4) * 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA0);
        tma_prefetch_desc(&tmA1);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), EPI_THREADS / 32); }
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(smem_u32(tmem_ptr_smem), Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const int chunks = a.c0_chunks + a.c1_chunks;
    const int num_kb = a.taps * chunks;
    const int TW = 1 << a.tw_log2, TH = BLOCK_M >> a.tw_log2;
    const int tiles_per_img = a.tiles_x * a.tiles_y;

    if (warp == 0) {
        // ===================================================================== TMA producer
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
                const int n_blk = tile % a.n_blocks;
                const int m = tile / a.n_blocks;
                const int img = m / tiles_per_img;
                const int rem = m - img * tiles_per_img;
                const int ty = rem / a.tiles_x, tx = rem - ty * a.tiles_x;
                const int x0 = tx * TW, y0 = ty * TH;
                for (int tap = 0; tap < a.taps; ++tap) {
                    const int dy = (a.taps == 9) ? tap / 3 - 1 : 0;
                    const int dx = (a.taps == 9) ? tap % 3 - 1 : 0;
                    for (int ch = 0; ch < chunks; ++ch) {
                        mbar_wait(empty_bar(stage), phase ^ 1u);
                        const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
                        const uint32_t sb = sa + A_STAGE_BYTES;
                        mbar_arrive_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
                        if (ch < a.c0_chunks) tma_load_4d(sa, &tmA0, full_bar(stage), ch * BLOCK_K, x0 + dx, y0 + dy, img);
                        else tma_load_4d(sa, &tmA1, full_bar(stage), (ch - a.c0_chunks) * BLOCK_K, x0 + dx, y0 + dy, img);
                        tma_load_2d(sb, &tmB, full_bar(stage), (tap * chunks + ch) * BLOCK_K, n_blk * BLOCK_N);
                        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer (whole warp, one elected lane issues)
        constexpr uint32_t idesc = make_idesc(BLOCK_N);
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
            mbar_wait(tempty_bar(acc), acc_phase ^ 1u);          // epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(full_bar(stage), phase);               // TMA bytes have landed
                tc_fence_after();
                const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
                const uint64_t da = make_sw128_desc(sa);
                const uint64_t db = make_sw128_desc(sa + A_STAGE_BYTES);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < BLOCK_K / 16; ++k)       // UMMA_K = 16: +32 bytes inside the swizzle row
                        umma_bf16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
                    umma_commit(empty_bar(stage));               // frees the smem slot when these MMAs retire
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
            if (elect_one()) umma_commit(tfull_bar(acc));        // accumulator complete -> epilogue
            __syncwarp();
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
    } else {
        // ===================================================================== epilogue (4 warps = 128 TMEM lanes)
        const int quad = warp & 3;                                   // TMEM lane quadrant this warp may access
        const int row = quad * 32 + lane;                            // accumulator row = input pixel within the tile
        const int et = threadIdx.x - 64;                             // 0..127
        const int lx = row & (TW - 1), ly = row >> a.tw_log2;
        float* s_bias = s_scale;                                     // [c_out] <= 2*BLOCK_N floats, loaded once per CTA
        for (int c = et; c < a.c_out; c += EPI_THREADS) s_bias[c] = a.shift[c];
        named_bar_sync(1, EPI_THREADS);
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
            const int n_blk = tile % a.n_blocks;
            const int m = tile / a.n_blocks;
            const int img = m / tiles_per_img;
            const int rem = m - img * tiles_per_img;
            const int ty = rem / a.tiles_x, tx = rem - ty * a.tiles_x;
            const int x = tx * TW + lx, y = ty * TH + ly;
            const bool valid = (x < a.W) && (y < a.H);

            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BLOCK_N);
#pragma unroll 1
            for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(t_row + (uint32_t)c0, r);
                tmem_ld_wait();
                const int n = n_blk * BLOCK_N + c0;                  // first GEMM column of this 32-wide chunk
                const int q = n / a.c_out, co = n - q * a.c_out;     // q = dy*2 + dx ; a chunk never straddles a quadrant
                const float* bias = s_bias + co;
                if (valid) {
                    // pixel-shuffle store: input pixel (y, x) -> output pixel (2y + dy, 2x + dx), channels co .. co+31
                    const long long pix = ((long long)img * (2 * a.H) + (2 * y + (q >> 1))) * (2 * a.W) + (2 * x + (q & 1));
                    uint4* d4 = reinterpret_cast<uint4*>(a.out + pix * a.c_out + co);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint4 o;
                        __nv_bfloat162 p0 = __floats2bfloat162_rn(__uint_as_float(r[8 * i + 0]) + bias[8 * i + 0], __uint_as_float(r[8 * i + 1]) + bias[8 * i + 1]);
                        __nv_bfloat162 p1 = __floats2bfloat162_rn(__uint_as_float(r[8 * i + 2]) + bias[8 * i + 2], __uint_as_float(r[8 * i + 3]) + bias[8 * i + 3]);
                        __nv_bfloat162 p2 = __floats2bfloat162_rn(__uint_as_float(r[8 * i + 4]) + bias[8 * i + 4], __uint_as_float(r[8 * i + 5]) + bias[8 * i + 5]);
                        __nv_bfloat162 p3 = __floats2bfloat162_rn(__uint_as_float(r[8 * i + 6]) + bias[8 * i + 6], __uint_as_float(r[8 * i + 7]) + bias[8 * i + 7]);
                        o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1);
                        o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
                        d4[i] = o;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));             // 4 arrivals (one per epilogue warp) free the accumulator
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

template <int BLOCK_N>
static int launch_cfg(const CUtensorMap& mA0, const CUtensorMap& mA1, const CUtensorMap& mB, const ConvArgs& args, cudaStream_t stream) {
    using Cfg = ConvCfg<BLOCK_N>;
    auto kern = conv_gemm_kernel<BLOCK_N>;
    static unsigned char smem_set[64] = {0};
    ADN_CUDA_TRY(ensure_dyn_smem(kern, Cfg::SMEM_BYTES, smem_set));
    const int sms = num_sms();
    const int grid = args.num_tiles < sms ? args.num_tiles : sms;
    kern<<<grid, CONV_THREADS, Cfg::SMEM_BYTES, stream>>>(mA0, mA1, mB, args);
    ADN_LAUNCH_CHECK();
    return ADN_OK;
}

// common driver: A sources on an (n,h,w) pixel grid; src1 may have a smaller spatial extent (h1,w1)
// ConvTranspose2d(k=2, s=2) as a GEMM over the input pixel grid: M = n*h*w, K = c_in, N = 4*c_out (q-major).
static int convt_gemm(const void* src, int c_in, int n, int h, int w, const void* w_packed, int c_out, const float* bias, void* out,
                      cudaStream_t stream) {
    if (!src || !w_packed || !bias || !out || n <= 0 || h <= 0 || w <= 0) return ADN_ERR_ARG;
    if (c_in <= 0 || (c_in % 64) || c_out <= 0 || (c_out % 64)) return ADN_ERR_ARG;
    if (!aligned16(src) || !aligned16(w_packed) || !aligned16(out)) return ADN_ERR_ARG;
    int st = check_device();
    if (st != ADN_OK) return st;
    const int n_total = 4 * c_out;

    // pixel tile: 8x16 or 16x8 (TH x TW), whichever wastes fewer padded pixels
    auto padded = [&](int th, int tw) { return (long long)((h + th - 1) / th) * th * ((w + tw - 1) / tw) * tw; };
    const int tw_log2 = padded(8, 16) <= padded(16, 8) ? 4 : 3;
    const int tw = 1 << tw_log2, th = BLOCK_M >> tw_log2;

    const int block_n = 256;                                  // 4*c_out is always a multiple of 256
    if (c_out > 2 * block_n) return ADN_ERR_ARG;              // bias staging area holds 2*BLOCK_N floats
    ConvArgs args;
    args.c0_chunks = c_in / 64; args.c1_chunks = 0; args.taps = 1;
    args.n_img = n; args.H = h; args.W = w;
    args.tiles_x = (w + tw - 1) / tw; args.tiles_y = (h + th - 1) / th; args.tw_log2 = tw_log2;
    args.n_total = n_total; args.c_out = c_out; args.n_blocks = n_total / block_n;
    const long long tiles = (long long)n * args.tiles_x * args.tiles_y * args.n_blocks;
    if (tiles > 0x7fffffffLL) return ADN_ERR_ARG;
    args.num_tiles = (int)tiles;
    args.shift = bias;
    args.out = (__nv_bfloat16*)out;

    CUtensorMap mA0, mB;
    st = make_act_map(&mA0, src, n, h, w, c_in, tw, th);
    if (st != ADN_OK) return st;
    st = make_weight_map(&mB, w_packed, n_total, c_in, block_n);
    if (st != ADN_OK) return st;
    return launch_cfg<256>(mA0, mA0, mB, args, stream);
}

}  // namespace adn

extern "C" int adn_convt2x2_bf16(const void* src, int c_in, int n, int h, int w, const void* w_packed, int c_out, const float* bias,
                                 void* out, void* stream) {
    return adn::convt_gemm(src, c_in, n, h, w, w_packed, c_out, bias, out, (cudaStream_t)stream);
}
