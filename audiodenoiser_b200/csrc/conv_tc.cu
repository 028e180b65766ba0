// conv_tc.cu -- ConvTranspose2d(k=2, s=2) + bias of the reference UNet (code/model.py:38,43) as a persistent, warp-specialised
// tcgen05/TMEM GEMM with a pixel-shuffle TMA-store epilogue, for sm_100a.
//
//   GEMM view     D[M = n*h*w input pixels, N = 4*Cout] = A[M, K = Cin] * B[N, K]^T   (bf16 x bf16 -> fp32);
//                 column n = q*Cout + co with q = dy*2 + dx is output pixel (2y+dy, 2x+dx), channel co.
//   A operand     per 64-channel chunk one 4-D TMA box {64 ch, TW, TH, 1 image} of the NHWC input (TH*TW = 128 pixels = the 128
//                 TMEM lanes), landing as a K-major SWIZZLE_128B tile.
//   B operand     packed weights [q][Cout][Cin] bf16 (K-major), 2-D TMA box {64, 256}.
//   roles         warp 0: TMA producer - warp 1: tcgen05.mma issuer (warp-uniform loop, elect.sync around the issue) -
//                 warps 2..9: epilogue (two sets of four warps split the column groups of a tile).  Ring of STAGES {A,B} buffers with full/empty mbarriers; two TMEM accumulators so the
//                 epilogue of tile i overlaps the MMAs of tile i+1.
//   epilogue      tcgen05.ld -> + bias -> bf16 -> a [128 px][64 ch] SWIZZLE_128B staging tile in smem -> ONE TMA store per
//                 (quadrant, 64-channel group) through a tensor map of the strided output view {co, x (stride 2), y (stride 2), n}
//                 based at (dy, dx): the pixel shuffle costs no scattered 16-byte stores (the first version of this kernel
//                 spent ~2x its HBM time in the LSU), and out-of-image rows / columns are clipped by the TMA engine.
#include "tc_common.cuh"

namespace adn {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;                 // bf16 elements = one 128-byte swizzle row
constexpr int T_BLOCK_N = 256;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int B_STAGE_BYTES = T_BLOCK_N * BLOCK_K * 2;
constexpr int T_STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;      // 48 KB
constexpr int T_STAGES = 3;
constexpr int T_OUT_STAGE = 128 * 128;                             // [128 px][64 ch] bf16
constexpr int T_MAX_COUT = 512;
constexpr int EPI_THREADS = 128;                                   // one epilogue set = 4 warps = the 4 TMEM lane quadrants
constexpr int EPI_SETS = 2;                                        // two sets split the 64-column groups of a tile between them
constexpr int CONV_THREADS = 64 + EPI_SETS * EPI_THREADS;
constexpr int T_AUX_BYTES = T_MAX_COUT * 4 + (2 * T_STAGES + 4) * 8 + 16;
constexpr int T_SMEM_BYTES = T_STAGES * T_STAGE_BYTES + 2 * EPI_SETS * T_OUT_STAGE + T_AUX_BYTES + 1024;

struct ConvTArgs {
    int k_chunks;                  // Cin / 64
    int n_img, H, W;               // input pixel grid
    int tiles_x, tiles_y, tw_log2; // pixel tile = TH x TW, TW = 1 << tw_log2, TH = 128 >> tw_log2
    int c_out, n_blocks, num_tiles;
    const float* bias;             // [c_out] (forward) or NULL
    int dgrad;                     // 0: forward ConvTranspose2d (one A map, four pixel-shuffle output maps)
                                   // 1: its data gradient (four quadrant A maps over d_out, K = 4*Cout, one dense output map)
    int a_chunks_per_map;          // dgrad: Cout / 64 chunks per quadrant
    int n_valid;                   // dgrad: valid GEMM columns (= Cin); the 256-wide N block may overhang
};

__global__ void __launch_bounds__(CONV_THREADS, 1)
convt_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
                  const __grid_constant__ CUtensorMap tmA3, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmO0, const __grid_constant__ CUtensorMap tmO1,
                  const __grid_constant__ CUtensorMap tmO2, const __grid_constant__ CUtensorMap tmO3, const ConvTArgs a) {
    extern __shared__ uint8_t smem_dyn[];
    const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_dyn + (smem_base - smem_u32(smem_dyn));

    // carve-up: [STAGES x {A, B}] [2 x output staging] [bias] [full STAGES] [empty STAGES] [tfull 2] [tempty 2] [tmem ptr]
    const uint32_t out_stage_base = smem_base + T_STAGES * T_STAGE_BYTES;
    constexpr uint32_t AUX_OFF = T_STAGES * T_STAGE_BYTES + 2 * EPI_SETS * T_OUT_STAGE;
    float* s_bias = reinterpret_cast<float*>(smem_gen + AUX_OFF);
    const uint32_t bar_base = smem_base + AUX_OFF + T_MAX_COUT * 4;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (T_STAGES + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * T_STAGES + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * T_STAGES + 2 + s); };
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem_gen + AUX_OFF + T_MAX_COUT * 4 + (2 * T_STAGES + 4) * 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmA1); tma_prefetch_desc(&tmA2); tma_prefetch_desc(&tmA3);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmO0); tma_prefetch_desc(&tmO1); tma_prefetch_desc(&tmO2); tma_prefetch_desc(&tmO3);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < T_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), EPI_SETS * EPI_THREADS / 32); }
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(smem_u32(tmem_ptr_smem), 2 * T_BLOCK_N);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

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
                for (int ch = 0; ch < a.k_chunks; ++ch) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t sa = smem_base + stage * T_STAGE_BYTES;
                    mbar_arrive_expect_tx(full_bar(stage), T_STAGE_BYTES);
                    if (!a.dgrad) {
                        tma_load_4d(sa, &tmA, full_bar(stage), ch * BLOCK_K, tx * TW, ty * TH, img);
                    } else {                                         // K chunk = (quadrant q, 64-channel block) of d_out
                        const int q = ch / a.a_chunks_per_map, cc = ch - q * a.a_chunks_per_map;
                        const CUtensorMap* ma = (q == 0) ? &tmA : (q == 1) ? &tmA1 : (q == 2) ? &tmA2 : &tmA3;
                        tma_load_4d(sa, ma, full_bar(stage), cc * BLOCK_K, tx * TW, ty * TH, img);
                    }
                    tma_load_2d(sa + A_STAGE_BYTES, &tmB, full_bar(stage), ch * BLOCK_K, n_blk * T_BLOCK_N);
                    if (++stage == T_STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer (whole warp, one elected lane issues)
        constexpr uint32_t idesc = make_idesc(T_BLOCK_N);
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
            mbar_wait(tempty_bar(acc), acc_phase ^ 1u);          // epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * T_BLOCK_N);
            for (int kb = 0; kb < a.k_chunks; ++kb) {
                mbar_wait(full_bar(stage), phase);               // TMA bytes have landed
                tc_fence_after();
                const uint32_t sa = smem_base + stage * T_STAGE_BYTES;
                const uint64_t da = make_sw128_desc(sa);
                const uint64_t db = make_sw128_desc(sa + A_STAGE_BYTES);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < BLOCK_K / 16; ++k)       // UMMA_K = 16: +32 bytes inside the swizzle row
                        umma_bf16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
                    umma_commit(empty_bar(stage));               // frees the smem slot when these MMAs retire
                }
                __syncwarp();
                if (++stage == T_STAGES) { stage = 0; phase ^= 1u; }
            }
            if (elect_one()) umma_commit(tfull_bar(acc));        // accumulator complete -> epilogue
            __syncwarp();
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
    } else {
        // ===================================================================== epilogue: 2 sets x 4 warps.  A set covers the 128 TMEM
        // lanes (warp % 4 = lane quadrant); set e handles the 64-column groups g = e, e + 2 of every tile with its own pair of
        // staging buffers and its own named barrier, so the 64 KB of output per tile is converted and stored by 8 warps.
        const int eset = (warp - 2) >> 2;
        const int quad = warp & 3;                                   // TMEM lane quadrant this warp may access
        const int row = quad * 32 + lane;                            // accumulator row = input pixel within the tile
        const int et = (threadIdx.x - 64) & (EPI_THREADS - 1);       // 0..127 within the set
        const int bar_id = 1 + eset;
        if (a.bias) for (int c = threadIdx.x - 64; c < a.c_out; c += EPI_SETS * EPI_THREADS) s_bias[c] = a.bias[c];
        named_bar_sync(3, EPI_SETS * EPI_THREADS);
        int acc = 0; uint32_t acc_phase = 0;
        uint32_t store_groups = 0;
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
            const int n_blk = tile % a.n_blocks;
            const int m = tile / a.n_blocks;
            const int img = m / tiles_per_img;
            const int rem = m - img * tiles_per_img;
            const int ty = rem / a.tiles_x, tx = rem - ty * a.tiles_x;

            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * T_BLOCK_N);
#pragma unroll 1
            for (int g = eset; g < T_BLOCK_N / 64; g += EPI_SETS) {  // one 64-column group = one (quadrant, 64-channel) store
                const int n = n_blk * T_BLOCK_N + g * 64;
                if (a.dgrad && n >= a.n_valid) break;                // uniform: the N block overhangs Cin
                const int q = a.dgrad ? 0 : n / a.c_out;             // a group never straddles a quadrant (c_out % 64 == 0)
                const int co = n - q * a.c_out;
                const uint32_t o_stage = out_stage_base + (uint32_t)(eset * 2 + (store_groups & 1u)) * T_OUT_STAGE;
                if (et == 0) bulk_wait_read<1>();                    // the store that last used this buffer has read it
                named_bar_sync(bar_id, EPI_THREADS);
                const uint32_t rbase = o_stage + (uint32_t)row * 128u;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t r[32];
                    tmem_ld32(t_row + (uint32_t)(g * 64 + half * 32), r);
                    tmem_ld_wait();
                    const float* bias = s_bias + co + half * 32;
                    const bool has_bias = a.bias != nullptr;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint32_t w[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float b0 = has_bias ? bias[8 * i + 2 * j] : 0.f, b1 = has_bias ? bias[8 * i + 2 * j + 1] : 0.f;
                            __nv_bfloat162 p = __floats2bfloat162_rn(__uint_as_float(r[8 * i + 2 * j]) + b0,
                                                                      __uint_as_float(r[8 * i + 2 * j + 1]) + b1);
                            w[j] = *reinterpret_cast<uint32_t*>(&p);
                        }
                        st_shared_v4(rbase + ((((uint32_t)(half * 4 + i)) ^ ((uint32_t)row & 7u)) << 4), w[0], w[1], w[2], w[3]);
                    }
                }
                fence_proxy_async();
                named_bar_sync(bar_id, EPI_THREADS);
                if (et == 0) {
                    const CUtensorMap* mo = (q == 0) ? &tmO0 : (q == 1) ? &tmO1 : (q == 2) ? &tmO2 : &tmO3;
                    tma_store_4d(mo, o_stage, co, tx * TW, ty * TH, img);
                    bulk_commit();
                }
                ++store_groups;
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));             // 8 arrivals (one per epilogue warp) free the accumulator
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
        if (et == 0) bulk_wait<0>();                                 // smem must outlive the last bulk stores
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) tmem_dealloc(tmem_base, 2 * T_BLOCK_N);
}

// ------------------------------------------------------------------------------------------------ host side
// strided view of the (n, 2h, 2w, c_out) output for quadrant (dy, dx): element (co, x, y, img) = out[img][2y+dy][2x+dx][co]
static int make_shuffle_map(CUtensorMap* map, void* out, int n, int h, int w, int c_out, int dy, int dx, int tw, int th) {
    PFN_tmapEncodeTiled enc = get_encode_fn();
    if (!enc) return ADN_ERR_DRIVER;
    char* base = static_cast<char*>(out) + ((size_t)dy * (2 * w) + dx) * c_out * 2;
    cuuint64_t dims[4] = {(cuuint64_t)c_out, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
    cuuint64_t strides[3] = {(cuuint64_t)2 * c_out * 2, (cuuint64_t)2 * (2 * w) * c_out * 2, (cuuint64_t)(2 * h) * (2 * w) * c_out * 2};
    cuuint32_t box[4] = {64u, (cuuint32_t)tw, (cuuint32_t)th, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? ADN_OK : ADN_ERR_DRIVER;
}

static int convt_gemm(const void* src, int c_in, int n, int h, int w, const void* w_packed, int c_out, const float* bias, void* out,
                      cudaStream_t stream) {
    if (!src || !w_packed || !bias || !out || n <= 0 || h <= 0 || w <= 0) return ADN_ERR_ARG;
    if (c_in <= 0 || (c_in % 64) || c_out <= 0 || (c_out % 64) || c_out > T_MAX_COUT) return ADN_ERR_ARG;
    if (!aligned16(src) || !aligned16(w_packed) || !aligned16(out)) return ADN_ERR_ARG;
    int st = check_device();
    if (st != ADN_OK) return st;
    const int n_total = 4 * c_out;                            // always a multiple of 256

    // pixel tile: 8x16 or 16x8 (TH x TW), whichever wastes fewer padded pixels
    auto padded = [&](int th, int tw) { return (long long)((h + th - 1) / th) * th * ((w + tw - 1) / tw) * tw; };
    const int tw_log2 = padded(8, 16) <= padded(16, 8) ? 4 : 3;
    const int tw = 1 << tw_log2, th = BLOCK_M >> tw_log2;

    ConvTArgs args;
    args.k_chunks = c_in / 64;
    args.n_img = n; args.H = h; args.W = w;
    args.tiles_x = (w + tw - 1) / tw; args.tiles_y = (h + th - 1) / th; args.tw_log2 = tw_log2;
    args.c_out = c_out; args.n_blocks = n_total / T_BLOCK_N;
    const long long tiles = (long long)n * args.tiles_x * args.tiles_y * args.n_blocks;
    if (tiles > 0x7fffffffLL) return ADN_ERR_ARG;
    args.num_tiles = (int)tiles;
    args.bias = bias;
    args.dgrad = 0; args.a_chunks_per_map = args.k_chunks; args.n_valid = n_total;

    CUtensorMap mA, mB, mO[4];
    st = make_act_map(&mA, src, n, h, w, c_in, tw, th);
    if (st != ADN_OK) return st;
    st = make_weight_map(&mB, w_packed, n_total, c_in, T_BLOCK_N);
    if (st != ADN_OK) return st;
    for (int q = 0; q < 4; ++q) {
        st = make_shuffle_map(&mO[q], out, n, h, w, c_out, q >> 1, q & 1, tw, th);
        if (st != ADN_OK) return st;
    }

    static unsigned char smem_set[64] = {0};
    ADN_CUDA_TRY(ensure_dyn_smem(convt_gemm_kernel, T_SMEM_BYTES, smem_set));
    const int sms = num_sms();
    const int grid = args.num_tiles < sms ? args.num_tiles : sms;
    convt_gemm_kernel<<<grid, CONV_THREADS, T_SMEM_BYTES, stream>>>(mA, mA, mA, mA, mB, mO[0], mO[1], mO[2], mO[3], args);
    ADN_LAUNCH_CHECK();
    return ADN_OK;
}

// quadrant (dy, dx) of a channel slice [c_off, c_off + c) of an (n, 2h, 2w, ld) tensor, as an (n, h, w, c) view
static int make_quadrant_view(CUtensorMap* map, const void* ptr, int n, int h, int w, int c, int ld, int c_off, int dy, int dx, int tw, int th) {
    PFN_tmapEncodeTiled enc = get_encode_fn();
    if (!enc) return ADN_ERR_DRIVER;
    const char* base = static_cast<const char*>(ptr) + (((size_t)dy * (2 * w) + dx) * ld + c_off) * 2;
    cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
    cuuint64_t strides[3] = {(cuuint64_t)2 * ld * 2, (cuuint64_t)2 * (2 * w) * ld * 2, (cuuint64_t)(2 * h) * (2 * w) * ld * 2};
    cuuint32_t box[4] = {64u, (cuuint32_t)tw, (cuuint32_t)th, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<char*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? ADN_OK : ADN_ERR_DRIVER;
}

// data gradient of ConvTranspose2d(k=2,s=2): d_in[p][ci] = sum_{q,co} d_out[2p+q][co] * W[ci][co][q]
//   GEMM: M = n*h*w input pixels, K = 4*Cout (chunk = quadrant x 64 channels, A boxes from the quadrant views of d_out),
//   N = Cin, B = weights packed [ci][q*Cout + co] bf16 (adn_pack_convt2x2_dgrad_weight_bf16).
static int convt_dgrad(const void* d_out, int out_ld, int out_off, int c_out, int n, int h, int w, const void* w_packed, int c_in, void* d_in,
                       cudaStream_t stream) {
    if (!d_out || !w_packed || !d_in || n <= 0 || h <= 0 || w <= 0) return ADN_ERR_ARG;
    if (c_in <= 0 || (c_in % 64) || c_out <= 0 || (c_out % 64) || out_off < 0 || out_off + c_out > out_ld || (out_ld % 8) || (out_off % 8)) return ADN_ERR_ARG;
    if (!aligned16(d_out) || !aligned16(w_packed) || !aligned16(d_in)) return ADN_ERR_ARG;
    int st = check_device();
    if (st != ADN_OK) return st;
    auto padded = [&](int th, int tw) { return (long long)((h + th - 1) / th) * th * ((w + tw - 1) / tw) * tw; };
    const int tw_log2 = padded(8, 16) <= padded(16, 8) ? 4 : 3;
    const int tw = 1 << tw_log2, th = BLOCK_M >> tw_log2;
    ConvTArgs args;
    args.k_chunks = 4 * (c_out / 64);
    args.n_img = n; args.H = h; args.W = w;
    args.tiles_x = (w + tw - 1) / tw; args.tiles_y = (h + th - 1) / th; args.tw_log2 = tw_log2;
    args.c_out = c_in;                                          // epilogue column -> channel of the single output map
    args.n_blocks = (c_in + T_BLOCK_N - 1) / T_BLOCK_N;
    const long long tiles = (long long)n * args.tiles_x * args.tiles_y * args.n_blocks;
    if (tiles > 0x7fffffffLL) return ADN_ERR_ARG;
    args.num_tiles = (int)tiles;
    args.bias = nullptr;
    args.dgrad = 1; args.a_chunks_per_map = c_out / 64; args.n_valid = c_in;
    CUtensorMap mA[4], mB, mO;
    for (int q = 0; q < 4; ++q) {
        st = make_quadrant_view(&mA[q], d_out, n, h, w, c_out, out_ld, out_off, q >> 1, q & 1, tw, th);
        if (st != ADN_OK) return st;
    }
    st = make_weight_map(&mB, w_packed, c_in, 4 * c_out, T_BLOCK_N);
    if (st != ADN_OK) return st;
    st = make_act_map(&mO, d_in, n, h, w, c_in, tw, th);
    if (st != ADN_OK) return st;
    static unsigned char smem_set[64] = {0};
    ADN_CUDA_TRY(ensure_dyn_smem(convt_gemm_kernel, T_SMEM_BYTES, smem_set));
    const int sms = num_sms();
    const int grid = args.num_tiles < sms ? args.num_tiles : sms;
    convt_gemm_kernel<<<grid, CONV_THREADS, T_SMEM_BYTES, stream>>>(mA[0], mA[1], mA[2], mA[3], mB, mO, mO, mO, mO, args);
    ADN_LAUNCH_CHECK();
    return ADN_OK;
}

}  // namespace adn

extern "C" int adn_convt2x2_dgrad_bf16(const void* d_out, int out_ld, int out_off, int c_out, int n, int h, int w, const void* w_packed,
                                       int c_in, void* d_in, void* stream) {
    return adn::convt_dgrad(d_out, out_ld, out_off, c_out, n, h, w, w_packed, c_in, d_in, (cudaStream_t)stream);
}

extern "C" int adn_convt2x2_bf16(const void* src, int c_in, int n, int h, int w, const void* w_packed, int c_out, const float* bias,
                                 void* out, void* stream) {
    return adn::convt_gemm(src, c_in, n, h, w, w_packed, c_out, bias, out, (cudaStream_t)stream);
}
