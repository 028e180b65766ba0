// conv_tc.cu -- the reference UNet's dense contractions (code/model.py:11,14 Conv2d 3x3; :38 ConvTranspose2d 2x2/2;
// :68 the 1x1 head) as ONE warp-specialised, persistent tcgen05/TMEM implicit-GEMM kernel for sm_100a.
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
//   epilogues     0: NHWC bf16 store; 1: ConvTranspose pixel-shuffle store (+bias); 2: fused 1x1 head -> fp32 (n,1,h,w).
#include <cuda.h>
#include <mutex>
#include "adn_common.cuh"

namespace adn {

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}
// Waits are bounded: a barrier that does not flip within ~2 s of SM clocks is a protocol bug, and a trap (reported as a
// launch failure through the C ABI) is far better than a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]^T, kind::f16 (bf16 operands, fp32 accumulate), issued by ONE thread
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on an mbarrier once every tcgen05 op issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 format): 8-row groups are 1024 B apart
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);          // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                              // leading byte offset (unused for swizzled K-major), bits [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;                    // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                              // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                              // layout: SWIZZLE_128B
    return d;
}

// instruction descriptor, kind::f16: D = f32, A = B = bf16, both K-major, M = 128
__host__ __device__ constexpr uint32_t make_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------ kernel
constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;                 // bf16 elements = one 128-byte swizzle row
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int CONV_THREADS = 192;
constexpr int EPI_THREADS = 128;

enum { EPI_NHWC = 0, EPI_CONVT = 1, EPI_HEAD = 2 };

struct ConvArgs {
    int c0_chunks, c1_chunks;      // 64-channel chunks taken from source 0 / source 1
    int taps;                      // 9 (3x3, pad 1) or 1
    int n_img, H, W;               // GEMM rows = n_img*H*W pixels of the A tensors' grid
    int tiles_x, tiles_y, tw_log2; // pixel tile = TH x TW, TW = 1 << tw_log2, TH = 128 >> tw_log2
    int n_total;                   // GEMM N (c_out, or 4*c_out for the transposed conv)
    int c_out;                     // channels of the output tensor
    int n_blocks;                  // n_total / BLOCK_N
    int num_tiles;                 // n_img*tiles_y*tiles_x*n_blocks
    int relu;
    int epi;
    const float* scale;            // [n_total] or NULL (= 1)
    const float* shift;            // [n_total] (EPI_CONVT: bias[c_out], indexed by channel)
    __nv_bfloat16* out;
    const float* head_w;           // EPI_HEAD: [64]
    const float* head_b;           // EPI_HEAD: [1]
    float* head_out;               // EPI_HEAD: (n,1,H,W)
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
    float* s_head = s_shift + BLOCK_N;
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
        // ===================================================================== MMA issuer
        if (lane == 0) {
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
#pragma unroll
                    for (int k = 0; k < BLOCK_K / 16; ++k)           // UMMA_K = 16: +32 bytes inside the swizzle row
                        umma_bf16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
                    umma_commit(empty_bar(stage));                   // frees the smem slot when these MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
                umma_commit(tfull_bar(acc));                         // accumulator complete -> epilogue
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else {
        // ===================================================================== epilogue (4 warps = 128 TMEM lanes)
        const int quad = warp & 3;                                   // TMEM lane quadrant this warp may access
        const int row = quad * 32 + lane;                            // accumulator row = pixel within the tile
        const int et = threadIdx.x - 64;                             // 0..127
        const int lx = row & (TW - 1), ly = row >> a.tw_log2;
        if (a.epi == EPI_HEAD && et < 64) s_head[et] = a.head_w[et];
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
            const int n_blk = tile % a.n_blocks;
            const int m = tile / a.n_blocks;
            const int img = m / tiles_per_img;
            const int rem = m - img * tiles_per_img;
            const int ty = rem / a.tiles_x, tx = rem - ty * a.tiles_x;
            const int x = tx * TW + lx, y = ty * TH + ly;
            const bool valid = (x < a.W) && (y < a.H);

            named_bar_sync(1, EPI_THREADS);                          // previous tile's readers of s_scale/s_shift are done
            for (int c = et; c < BLOCK_N; c += EPI_THREADS) {
                const int n = n_blk * BLOCK_N + c;
                s_scale[c] = a.scale ? a.scale[n] : 1.0f;
                s_shift[c] = (a.epi == EPI_CONVT) ? a.shift[n % a.c_out] : a.shift[n];
            }
            named_bar_sync(1, EPI_THREADS);

            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BLOCK_N);
            float head_acc = 0.f;
#pragma unroll 1
            for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(t_row + (uint32_t)c0, r);
                tmem_ld_wait();
                float v[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float f = fmaf(__uint_as_float(r[i]), s_scale[c0 + i], s_shift[c0 + i]);
                    v[i] = a.relu ? fmaxf(f, 0.f) : f;
                }
                if (a.epi == EPI_HEAD) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) head_acc = fmaf(v[i], s_head[c0 + i], head_acc);
                } else if (valid) {
                    const int n = n_blk * BLOCK_N + c0;              // first GEMM column of this 32-wide chunk
                    __nv_bfloat16* dst;
                    if (a.epi == EPI_CONVT) {
                        const int q = n / a.c_out, co = n - q * a.c_out;         // q = dy*2 + dx
                        const long long pix = ((long long)img * (2 * a.H) + (2 * y + (q >> 1))) * (2 * a.W) + (2 * x + (q & 1));
                        dst = a.out + pix * a.c_out + co;
                    } else {
                        const long long pix = ((long long)img * a.H + y) * a.W + x;
                        dst = a.out + pix * a.c_out + n;
                    }
                    uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint4 o;
                        __nv_bfloat162 p0 = __floats2bfloat162_rn(v[8 * i + 0], v[8 * i + 1]);
                        __nv_bfloat162 p1 = __floats2bfloat162_rn(v[8 * i + 2], v[8 * i + 3]);
                        __nv_bfloat162 p2 = __floats2bfloat162_rn(v[8 * i + 4], v[8 * i + 5]);
                        __nv_bfloat162 p3 = __floats2bfloat162_rn(v[8 * i + 6], v[8 * i + 7]);
                        o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1);
                        o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
                        d4[i] = o;
                    }
                }
            }
            if (a.epi == EPI_HEAD && valid)
                a.head_out[((long long)img * a.H + y) * a.W + x] = head_acc + a.head_b[0];
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

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_tmapEncodeTiled get_encode_fn() {
    static PFN_tmapEncodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
    });
    return fn;
}

// NHWC bf16 activation (n,h,w,c): dims innermost-first {c, w, h, n}; box {64, TW, TH, 1}; zero OOB fill
static int make_act_map(CUtensorMap* map, const void* ptr, int n, int h, int w, int c, int tw, int th) {
    PFN_tmapEncodeTiled enc = get_encode_fn();
    if (!enc) return ADN_ERR_DRIVER;
    cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
    cuuint64_t strides[3] = {(cuuint64_t)c * 2, (cuuint64_t)w * c * 2, (cuuint64_t)h * w * c * 2};
    cuuint32_t box[4] = {(cuuint32_t)BLOCK_K, (cuuint32_t)tw, (cuuint32_t)th, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? ADN_OK : ADN_ERR_DRIVER;
}

// packed weights [n_total][k_total] bf16: dims {k_total, n_total}; box {64, block_n}
static int make_weight_map(CUtensorMap* map, const void* ptr, int n_total, int k_total, int block_n) {
    PFN_tmapEncodeTiled enc = get_encode_fn();
    if (!enc) return ADN_ERR_DRIVER;
    cuuint64_t dims[2] = {(cuuint64_t)k_total, (cuuint64_t)n_total};
    cuuint64_t strides[1] = {(cuuint64_t)k_total * 2};
    cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)block_n};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? ADN_OK : ADN_ERR_DRIVER;
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

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// common driver: A sources on an (n,h,w) pixel grid; src1 may have a smaller spatial extent (h1,w1)
static int conv_gemm(const void* src0, int c0, const void* src1, int c1, int h1, int w1, int n, int h, int w, int taps,
                     const void* w_packed, int n_total, int c_out, const float* scale, const float* shift, int relu, int epi,
                     void* out, const float* head_w, const float* head_b, float* head_out, cudaStream_t stream) {
    if (!src0 || !w_packed || !shift || n <= 0 || h <= 0 || w <= 0) return ADN_ERR_ARG;
    if (c0 <= 0 || (c0 % 64) || c1 < 0 || (c1 % 64) || (c1 > 0 && !src1)) return ADN_ERR_ARG;
    if (n_total <= 0 || (n_total % 64) || c_out <= 0 || (c_out % 64)) return ADN_ERR_ARG;
    if (c1 > 0 && (h1 > h || w1 > w || h1 <= 0 || w1 <= 0)) return ADN_ERR_ARG;
    if (!aligned16(src0) || !aligned16(w_packed) || (src1 && !aligned16(src1))) return ADN_ERR_ARG;
    if (epi == EPI_HEAD) { if (!head_w || !head_b || !head_out || n_total != 64) return ADN_ERR_ARG; }
    else if (!out || !aligned16(out)) return ADN_ERR_ARG;
    int st = check_device();
    if (st != ADN_OK) return st;

    // pixel tile: 8x16 or 16x8 (TH x TW), whichever wastes fewer padded pixels
    auto padded = [&](int th, int tw) { return (long long)((h + th - 1) / th) * th * ((w + tw - 1) / tw) * tw; };
    const int tw_log2 = padded(8, 16) <= padded(16, 8) ? 4 : 3;
    const int tw = 1 << tw_log2, th = BLOCK_M >> tw_log2;

    const int block_n = (n_total % 256 == 0) ? 256 : (n_total % 128 == 0) ? 128 : 64;
    ConvArgs args;
    args.c0_chunks = c0 / 64; args.c1_chunks = c1 / 64; args.taps = taps;
    args.n_img = n; args.H = h; args.W = w;
    args.tiles_x = (w + tw - 1) / tw; args.tiles_y = (h + th - 1) / th; args.tw_log2 = tw_log2;
    args.n_total = n_total; args.c_out = c_out; args.n_blocks = n_total / block_n;
    const long long tiles = (long long)n * args.tiles_x * args.tiles_y * args.n_blocks;
    if (tiles > 0x7fffffffLL) return ADN_ERR_ARG;
    args.num_tiles = (int)tiles;
    args.relu = relu; args.epi = epi; args.scale = scale; args.shift = shift;
    args.out = (__nv_bfloat16*)out; args.head_w = head_w; args.head_b = head_b; args.head_out = head_out;

    CUtensorMap mA0, mA1, mB;
    st = make_act_map(&mA0, src0, n, h, w, c0, tw, th);
    if (st != ADN_OK) return st;
    if (c1 > 0) st = make_act_map(&mA1, src1, n, h1, w1, c1, tw, th); else mA1 = mA0;
    if (st != ADN_OK) return st;
    st = make_weight_map(&mB, w_packed, n_total, taps * (c0 + c1), block_n);
    if (st != ADN_OK) return st;

    switch (block_n) {
        case 256: return launch_cfg<256>(mA0, mA1, mB, args, stream);
        case 128: return launch_cfg<128>(mA0, mA1, mB, args, stream);
        default: return launch_cfg<64>(mA0, mA1, mB, args, stream);
    }
}

}  // namespace adn

extern "C" int adn_conv3x3_bn_relu_bf16(const void* src0, int c0, const void* src1, int c1, int h1, int w1, int n, int h, int w,
                                        const void* w_packed, int c_out, const float* scale, const float* shift, void* out,
                                        void* pool_out, void* stream) {
    int st = adn::conv_gemm(src0, c0, src1, c1, h1, w1, n, h, w, 9, w_packed, c_out, c_out, scale, shift, 1, adn::EPI_NHWC, out,
                            nullptr, nullptr, nullptr, (cudaStream_t)stream);
    if (st != ADN_OK || !pool_out) return st;
    return adn_maxpool2x2_bf16(out, n, h, w, c_out, pool_out, stream);
}

extern "C" int adn_conv3x3_bn_relu_head_f32(const void* src0, int c0, const void* src1, int c1, int h1, int w1, int n, int h, int w,
                                            const void* w_packed, int c_out, const float* scale, const float* shift,
                                            const float* head_w, const float* head_b, float* out_f32, void* stream) {
    if (c_out != 64) return ADN_ERR_ARG;
    return adn::conv_gemm(src0, c0, src1, c1, h1, w1, n, h, w, 9, w_packed, c_out, c_out, scale, shift, 1, adn::EPI_HEAD, nullptr,
                          head_w, head_b, out_f32, (cudaStream_t)stream);
}

extern "C" int adn_convt2x2_bf16(const void* src, int c_in, int n, int h, int w, const void* w_packed, int c_out, const float* bias,
                                 void* out, void* stream) {
    return adn::conv_gemm(src, c_in, nullptr, 0, 0, 0, n, h, w, 1, w_packed, 4 * c_out, c_out, nullptr, bias, 0, adn::EPI_CONVT, out,
                          nullptr, nullptr, nullptr, (cudaStream_t)stream);
}
