// stft.cu -- fused framing + periodic-Hann window + 512-point real FFT + magnitude (or complex) for sm_100a.
//
// Replaces librosa.stft + librosa.magphase at code/create_train_dataset.py:167-173 (center=False) and
// code/create_test_dataset.py:39-40 (center=True, zero padding) of the reference.
//
// Mapping ("lane = frame").  One CTA = 8 warps owns a tile of TF = 32 consecutive frames of one clip; lane t of every
// warp works on frame t0 + t, so every shared-memory access of the transform is [point][frame] (conflict-free by
// construction) and every global store is a run of 32 consecutive frames of one bin row -- the reference .npy layout
// (257, T) with T contiguous -- straight from registers, with no output transpose.
//   1. the 35 hops (4 480 samples) the tile touches are staged once with cp.async into a [hop][130]-padded buffer
//      (frames overlap by 75 %: each sample is read once from HBM and 4x from shared memory; the pad makes the
//      lane-strided float2 reads conflict-free), zero fill for the centred edges; the NEXT tile's samples are in flight
//      (second buffer) while this tile is transformed;
//   2. the 512 real samples of a frame are packed as 256 complex points, n = 16 n1 + n2.  Pass 1: warp w transforms the
//      columns n2 = w, w + 8 (radix-16 in registers, window folded into the load, W256 twiddles) into a
//      [k1][n2][frame] work array.  Pass 2: warp w transforms the rows k1 = w and 16 - w (0 and 8 for warp 0), which
//      hold bins k and 256 - k of the real-FFT split in the SAME thread, so the split, the magnitude and the store need
//      no further exchange.
// Window / twiddle tables are indexed by warp-uniform values only (broadcast shared-memory loads): no registers are
// spent on them, so pass 2 can hold two rows (64 registers) per thread.
#define ADN_PACKED_FP32 1            // packed fp32x2 butterflies (see adn_common.cuh)
#include <cstdlib>
#include "tc_common.cuh"
#include "adn_tables.inc"

namespace adn {

constexpr int TF = 32;                                   // frames per tile = lanes per warp
constexpr int STFT_THREADS = 256;
constexpr int ST_ROWS = TF + 3;                          // hops touched by a tile
constexpr int ST_ROW_STRIDE = 130;                       // floats; 130 mod 32 = 2 -> lane-strided float2 reads hit 32 banks
constexpr int ST_SAMPLES_BYTES = (ST_ROWS * ST_ROW_STRIDE * 4 + 15) / 16 * 16;
constexpr int ST_WORK_BYTES = 256 * TF * 8;
constexpr int ST_TABLE_BYTES = 3 * 256 * 8;             // 0.5*hann as float2[256], W256^(n2 k1) [16][16], W512^k [256]
constexpr int ST_SMEM_BYTES = 2 * ST_SAMPLES_BYTES + ST_WORK_BYTES + ST_TABLE_BYTES;

__device__ __forceinline__ uint32_t st_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// base + k * stride_bytes as ONE IMAD.WIDE (k folds to an immediate after unrolling)
template <typename T>
__device__ __forceinline__ T* row_ptr(T* base, int stride_bytes, int k) {
    long long r;
    asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(r) : "r"(stride_bytes), "r"(k), "l"(reinterpret_cast<long long>(base)));
    return reinterpret_cast<T*>(r);
}
__device__ __forceinline__ float sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// samples [s0, s0 + 128 * rows) of one clip -> buf[row][0..127], zero outside [0, length); s0 is a multiple of 128.
// Interior full tiles (the common case) take a path whose addresses are all immediates: thread `tid` copies the pair at
// column 2 (tid & 63) of rows (tid >> 6) + 4 i.
__device__ __forceinline__ void stft_stage(float* buf, const float* __restrict__ src, const float* dummy, int s0,
                                           int length, int rows, int tid) {
    const bool al8 = (reinterpret_cast<uintptr_t>(src) & 7) == 0;
    if (al8 && rows == ST_ROWS && s0 >= 0 && s0 + ST_ROWS * ADN_HOP <= length) {
        const float* g = src + s0 + (tid >> 6) * ADN_HOP + ((tid & 63) << 1);
        const uint32_t dst = st_smem_u32(buf + (tid >> 6) * ST_ROW_STRIDE + ((tid & 63) << 1));
#pragma unroll
        for (int i = 0; i < 8; ++i) cp_async8(dst + i * 4 * ST_ROW_STRIDE * 4, g + i * 4 * ADN_HOP, 8);
        if (tid < 64 * (ST_ROWS - 32)) cp_async8(dst + 8 * 4 * ST_ROW_STRIDE * 4, g + 8 * 4 * ADN_HOP, 8);
        return;
    }
    if (length <= 0) src = dummy;                        // nothing is read (src-size 0), but keep the address valid
    const int g_last = (length - 1) & ~1;                // clamp target: even (8-byte aligned when src is) and in range
    if (al8) {
        // edge tiles (zero padding in front of the clip, ragged tail): same thread -> (row, column pair) map as the interior path,
        // with a per-row byte count.  s0 is a multiple of 128 and the pair index is even, so a pair never straddles sample 0;
        // at the tail it holds 2, 1 or 0 valid samples (cp.async zero-fills the rest).
        const int r0 = tid >> 6;
        const int g0 = s0 + r0 * ADN_HOP + ((tid & 63) << 1);
        const uint32_t dst = st_smem_u32(buf + r0 * ST_ROW_STRIDE + ((tid & 63) << 1));
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            if (r0 + 4 * i < rows) {
                const int g = g0 + 4 * i * ADN_HOP;
                const int nvalid = min(max(length - g, 0), 2);
                cp_async8(dst + i * 4 * ST_ROW_STRIDE * 4, src + min(max(g, 0), max(g_last, 0)), g < 0 ? 0 : 4 * nvalid);
            }
        }
        return;
    }
    for (int p = tid; p < rows * 64; p += STFT_THREADS) {
        const int row = p >> 6, c = (p & 63) << 1;
        const int g = s0 + row * ADN_HOP + c;
        const uint32_t dst = st_smem_u32(buf + row * ST_ROW_STRIDE + c);
        const bool ok0 = (unsigned)g < (unsigned)length, ok1 = (unsigned)(g + 1) < (unsigned)length;
        const float* sp = src + max(0, min(g, g_last));
        cp_async4(dst, sp, ok0 ? 4 : 0);
        cp_async4(dst + 4, ok1 ? sp + 1 : sp, ok1 ? 4 : 0);
    }
}

// CROP: additionally (or only, when out == nullptr) emit the training tensor of SpectrogramDataset (data_loader.py:41-72): the
// magnitudes rounded through float16 and cropped to (f_out, t_out), as a second output of the same pass.
template <bool COMPLEX_OUT, bool CROP>
__global__ void __launch_bounds__(STFT_THREADS, 2)
stft_kernel(const float* __restrict__ wave, long long n_clips, int length, long long clip_stride, int center,
            int n_frames, int tiles_per_clip, float* __restrict__ out, float* __restrict__ crop, int f_out, int t_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* const samples0 = reinterpret_cast<float*>(smem_raw);
    float* const samples1 = reinterpret_cast<float*>(smem_raw + ST_SAMPLES_BYTES);
    float2* const work = reinterpret_cast<float2*>(smem_raw + 2 * ST_SAMPLES_BYTES);   // [k1*16 + n2][frame]
    // tables in shared memory, read with warp-uniform addresses (broadcast, one wavefront each).  Indexed __constant__
    // loads were measured to miss the SM's small constant cache and saturate the GPC-level one (ncu gcc__ 57 %).
    float2* const s_hann = reinterpret_cast<float2*>(smem_raw + 2 * ST_SAMPLES_BYTES + ST_WORK_BYTES);
    float2* const s_tw256 = s_hann + 256;
    float2* const s_tw512 = s_hann + 512;

    const int tid = threadIdx.x, lane = tid & 31;
    const int w = __shfl_sync(0xffffffffu, tid >> 5, 0);          // warp index, provably warp-uniform for the compiler
    // layouts chosen so that the 16 values a warp needs are contiguous: two twiddles / window pairs per 128-bit broadcast load
    s_hann[tid] = adn_c_hann512_half[16 * (tid & 15) + (tid >> 4)];      // [n2][n1]: window at samples 2(16 n1 + n2), +1
    s_tw256[tid] = adn_c_tw256[tid >> 4][tid & 15];                      // [n2][k1]
    s_tw512[tid] = adn_c_tw512[(tid >> 4) + 16 * (tid & 15)];            // [a][k2]: W512^(a + 16 k2)
    const int total_tiles = (int)n_clips * tiles_per_clip;        // < 2^31 (host check)
    const int pad = center ? ADN_N_FFT / 2 : 0;

    auto stage = [&](int clip, int t0, float* buf) {
        const int nf = min(TF, n_frames - t0);
        stft_stage(buf, wave + clip * clip_stride, out, t0 * ADN_HOP - pad, length, nf + 3, tid);
    };

    int tile_id = blockIdx.x;
    int clip = tile_id / tiles_per_clip, t0 = (tile_id - clip * tiles_per_clip) * TF;
    if (tile_id < total_tiles) stage(clip, t0, samples0);
    cp_async_commit();
    int parity = 0;
    for (; tile_id < total_tiles; tile_id += gridDim.x, parity ^= 1) {
        const float* cur = parity ? samples1 : samples0;
        const int nf = min(TF, n_frames - t0);
        const int cur_clip = clip, cur_t0 = t0;
        {                                                         // decompose the next tile once; it becomes `cur` next round
            const int nxt = tile_id + gridDim.x;
            clip = nxt / tiles_per_clip;
            t0 = (nxt - clip * tiles_per_clip) * TF;
        }
        cp_async_wait_all();
        __syncthreads();             // samples of this tile visible; every warp is done with `work` of the previous tile

        // ---- pass 1: columns n2 = w, w + 8 of frame t0 + lane
        const float* fr = cur + lane * ST_ROW_STRIDE;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int n2 = w + 8 * c;
            float2 v[16], tw[16];
            const float4* hw = reinterpret_cast<const float4*>(s_hann + n2 * 16);
            const float4* tq = reinterpret_cast<const float4*>(s_tw256 + n2 * 16);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 q = hw[j];
                const float2 s0 = *reinterpret_cast<const float2*>(fr + ((2 * j) >> 2) * ST_ROW_STRIDE + 32 * ((2 * j) & 3) + 2 * n2);
                const float2 s1 = *reinterpret_cast<const float2*>(fr + ((2 * j + 1) >> 2) * ST_ROW_STRIDE + 32 * ((2 * j + 1) & 3) + 2 * n2);
                v[2 * j] = pk_mul(s0, make_float2(q.x, q.y));
                v[2 * j + 1] = pk_mul(s1, make_float2(q.z, q.w));
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {                                // in flight during the butterflies
                const float4 q = tq[j];
                tw[2 * j] = make_float2(q.x, q.y); tw[2 * j + 1] = make_float2(q.z, q.w);
            }
            dft16<false>(v);                                     // over n1: Y[k1]
#pragma unroll
            for (int k1 = 1; k1 < 16; ++k1) v[k1] = cmul(v[k1], tw[k1]);
#pragma unroll
            for (int k1 = 0; k1 < 16; ++k1) work[(k1 * 16 + n2) * TF + lane] = v[k1];
        }

        // next tile's samples stream into the other buffer while this tile finishes
        if (tile_id + gridDim.x < total_tiles) stage(clip, t0, parity ? samples0 : samples1);
        cp_async_commit();
        __syncthreads();

        // ---- pass 2 + real-FFT split.  Z = half-scale packed transform; for a pair (k, 256-k):
        //      E = Zk + conj(Zp), O = -i (Zk - conj(Zp)), X[k] = E + W512^k O, conj(X[256-k]) = E - W512^k O
        const bool live = lane < nf;
        const int a = w, b = (w == 0) ? 8 : 16 - w;
        float2 A[16], B[16];
#pragma unroll
        for (int n2 = 0; n2 < 16; ++n2) A[n2] = work[(a * 16 + n2) * TF + lane];
#pragma unroll
        for (int n2 = 0; n2 < 16; ++n2) B[n2] = work[(b * 16 + n2) * TF + lane];
        dft16<false>(A);                                         // A[k2] = Z[a + 16 k2]
        dft16<false>(B);                                         // B[k2] = Z[b + 16 k2]
        // split in place: zk <- X[k], zp <- conj(X[256-k])
        auto split = [&](float2 tw, float2& zk, float2& zp) {                        // tw = W512^k
            const float2 e = pk_add(zk, make_float2(zp.x, -zp.y));                    // Zk + conj(Zp)
            const float2 o = pk_add(rot_mi(zk), make_float2(zp.y, zp.x));             // -i (Zk - conj(Zp))
            const float2 wo = cmul(o, tw);
            zk = cadd(e, wo);
            zp = csub(e, wo);
        };
        const float4* twa = reinterpret_cast<const float4*>(s_tw512 + a * 16);        // W512^(a + 16 k2), two per load
        const float4* twb = reinterpret_cast<const float4*>(s_tw512 + 8 * 16);        // W512^(8 + 16 k2) (warp 0)
        if (!COMPLEX_OUT && !CROP && w != 0) {
            // Hot path (|STFT| only): every pair is stored as soon as it is split -- the 32 stores and square roots of a thread are
            // spread over the split arithmetic instead of queueing behind it (STG / MUFU were 26 % of pass 2's stall samples), and the
            // two row pointers advance by one 64-bit add per store (rows a + 16 k2 upwards, rows b + 16 k2 downwards from k2 = 15)
            // instead of a multiply-wide plus two adds.
            const long long row0 = ((long long)cur_clip * ADN_N_BINS) * n_frames + cur_t0 + lane;
            const long long step = 16LL * n_frames;
            float* pa = out + row0 + (long long)a * n_frames;
            float* pb = out + row0 + (long long)(b + 240) * n_frames;
#pragma unroll
            for (int j = 0; j < 8; ++j) {                                             // 256 - k = b + 16 (15 - k2)
                const float4 q = twa[j];
                split(make_float2(q.x, q.y), A[2 * j], B[15 - 2 * j]);
                split(make_float2(q.z, q.w), A[2 * j + 1], B[14 - 2 * j]);
                const float m0 = sqrt_approx(fmaf(A[2 * j].x, A[2 * j].x, A[2 * j].y * A[2 * j].y));
                const float m1 = sqrt_approx(fmaf(B[15 - 2 * j].x, B[15 - 2 * j].x, B[15 - 2 * j].y * B[15 - 2 * j].y));
                const float m2 = sqrt_approx(fmaf(A[2 * j + 1].x, A[2 * j + 1].x, A[2 * j + 1].y * A[2 * j + 1].y));
                const float m3 = sqrt_approx(fmaf(B[14 - 2 * j].x, B[14 - 2 * j].x, B[14 - 2 * j].y * B[14 - 2 * j].y));
                if (live) {
                    __stcs(pa, m0);
                    __stcs(pb, m1);
                    __stcs(pa + step, m2);
                    __stcs(pb - step, m3);
                }
                pa += 2 * step;
                pb -= 2 * step;
            }
            continue;                                            // next tile (the loop's first barrier orders `work`)
        }
        if (w != 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {                                             // 256 - k = b + 16 (15 - k2)
                const float4 q = twa[j];
                split(make_float2(q.x, q.y), A[2 * j], B[15 - 2 * j]);
                split(make_float2(q.z, q.w), A[2 * j + 1], B[14 - 2 * j]);
            }
        } else {
            {                                                    // bins 0 and 256 (real): X[0] = 2(Re + Im), X[256] = 2(Re - Im)
                const float2 z = A[0];
                A[0] = make_float2(2.f * (z.x + z.y), 2.f * (z.x - z.y));
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 q = twa[j];                                              // a = 0: W512^(16 k2)
                if (j > 0) split(make_float2(q.x, q.y), A[2 * j], A[16 - 2 * j]);
                split(make_float2(q.z, q.w), A[2 * j + 1], A[15 - 2 * j]);
            }
            A[8] = make_float2(2.f * A[8].x, -2.f * A[8].y);     // k = 128 pairs with itself: X[128] = 2 conj(Z[128])
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 q = twb[j];
                split(make_float2(q.x, q.y), B[2 * j], B[15 - 2 * j]);
                split(make_float2(q.z, q.w), B[2 * j + 1], B[14 - 2 * j]);
            }
        }
        // Where the values live now (row = bin):
        //   w != 0: A[k2] = X[a + 16 k2], B[j] = conj X[b + 16 j]
        //   w == 0: A[0] = (X[0], X[256]); A[k2<8] = X[16 k2]; A[8] = X[128]; A[k2>8] = conj X[16 k2];
        //           B[k2<8] = X[8 + 16 k2]; B[k2>=8] = conj X[8 + 16 k2]
        const long long row0 = ((long long)cur_clip * ADN_N_BINS) * n_frames + cur_t0 + lane;   // 257 * T < 2^31 (host check)
        const int s16 = 16 * n_frames * (COMPLEX_OUT ? 8 : 4);   // bytes between rows k and k + 16
        if (COMPLEX_OUT) {
            float2* const pa = reinterpret_cast<float2*>(out) + row0 + a * n_frames;
            float2* const pb = reinterpret_cast<float2*>(out) + row0 + b * n_frames;
            if (live) {
                if (w != 0) {
#pragma unroll
                    for (int k2 = 0; k2 < 16; ++k2) {
                        __stcs(row_ptr(pa, s16, k2), A[k2]);
                        __stcs(row_ptr(pb, s16, k2), make_float2(B[k2].x, -B[k2].y));
                    }
                } else {
                    __stcs(pa, make_float2(A[0].x, 0.f));
                    __stcs(row_ptr(pa, s16, 16), make_float2(A[0].y, 0.f));
#pragma unroll
                    for (int k2 = 1; k2 < 16; ++k2) __stcs(row_ptr(pa, s16, k2), make_float2(A[k2].x, k2 > 8 ? -A[k2].y : A[k2].y));
#pragma unroll
                    for (int k2 = 0; k2 < 16; ++k2) __stcs(row_ptr(pb, s16, k2), make_float2(B[k2].x, k2 >= 8 ? -B[k2].y : B[k2].y));
                }
            }
        } else {
            float* const pa = out + row0 + a * n_frames;
            float* const pb = out + row0 + b * n_frames;
            float ma[16], mb[16];
#pragma unroll
            for (int k2 = 0; k2 < 16; ++k2) {
                ma[k2] = sqrt_approx(fmaf(A[k2].x, A[k2].x, A[k2].y * A[k2].y));
                mb[k2] = sqrt_approx(fmaf(B[k2].x, B[k2].x, B[k2].y * B[k2].y));
            }
            if (w == 0) ma[0] = fabsf(A[0].x);
            if (live && (!CROP || out != nullptr)) {
                if (w == 0) __stcs(row_ptr(pa, s16, 16), fabsf(A[0].y));
#pragma unroll
                for (int k2 = 0; k2 < 16; ++k2) {
                    __stcs(row_ptr(pa, s16, k2), ma[k2]);
                    __stcs(row_ptr(pb, s16, k2), mb[k2]);
                }
            }
            if (CROP) {
                const int tf = cur_t0 + lane;
                if (live && tf < t_out) {
                    float* const cp = crop + (long long)cur_clip * f_out * t_out + tf;
                    if (w == 0 && 256 < f_out) cp[256 * t_out] = __half2float(__float2half_rn(fabsf(A[0].y)));
#pragma unroll
                    for (int k2 = 0; k2 < 16; ++k2) {
                        if (a + 16 * k2 < f_out) cp[(a + 16 * k2) * t_out] = __half2float(__float2half_rn(ma[k2]));
                        if (b + 16 * k2 < f_out) cp[(b + 16 * k2) * t_out] = __half2float(__float2half_rn(mb[k2]));
                    }
                }
            }
        }
    }
    cp_async_wait_all();
}


// ------------------------------------------------------------------------------------------------ warp-specialised variant
// stft_kernel's two CTAs per SM run the same barrier-separated phases and stall together: after each __syncthreads all 16
// warps of the SM issue their shared-memory loads at once and nobody has arithmetic to issue until the data arrive (LDS was
// 26 % / 17 % of the stall samples of pass 1 / pass 2; the kernel issues on 67 % of the cycles although it is issue-bound:
// a packed fp32x2 instruction occupies the issue port for two cycles, scripts/microbench/issue_mix.cu).
// stft_ws_kernel removes the CTA-wide barriers.  One CTA of 16 warps per SM: warps 0-7 run pass 1 (window, radix-16 over n1,
// twiddles -> exchange array), warps 8-15 run pass 2 (radix-16 over n2, real-FFT split, magnitudes, stores), one tile apart,
// over TWO exchange arrays and THREE sample buffers.  All hand-offs are mbarriers with one arrival per warp, so a warp only
// ever waits for data it needs: pass-1 warps drift apart (their load bursts meet other warps' butterflies), and the pass-2
// warps work on tile t while pass 1 already transforms tile t + 1.
constexpr int WS_THREADS = 512;
constexpr int WS_SAMPLE_SLOTS = 3;
constexpr int WS_SMEM_BYTES = WS_SAMPLE_SLOTS * ST_SAMPLES_BYTES + 2 * ST_WORK_BYTES + ST_TABLE_BYTES + 128;

__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

// TFIX: the frame count T as a compile-time constant (0 = run-time).  For the three shapes the reference scripts produce -- 188
// (create_test_dataset.py, 3 s @ 8 kHz), 122 (create_train_dataset.py chunks), 1034 (3 s @ 44.1 kHz) -- the 32 row addresses of a
// thread's stores are then `base + immediate`: no address instruction at all (the pointer chains of the generic path cost 64 of
// ~1 350 issue slots per warp-tile).
template <int TFIX>
__global__ void __launch_bounds__(WS_THREADS, 1)
stft_ws_kernel(const float* __restrict__ wave, long long n_clips, int length, long long clip_stride, int center,
               int n_frames_rt, int tiles_per_clip, float* __restrict__ out) {
    const int n_frames = TFIX ? TFIX : n_frames_rt;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* const samples = reinterpret_cast<float*>(smem_raw);                                     // [slot][hop][130]
    float2* const work0 = reinterpret_cast<float2*>(smem_raw + WS_SAMPLE_SLOTS * ST_SAMPLES_BYTES);   // [buf][k1*16 + n2][frame]
    float2* const s_hann = reinterpret_cast<float2*>(smem_raw + WS_SAMPLE_SLOTS * ST_SAMPLES_BYTES + 2 * ST_WORK_BYTES);
    float2* const s_tw256 = s_hann + 256;
    float2* const s_tw512 = s_hann + 512;
    const uint32_t bars = smem_u32(smem_raw + WS_SAMPLE_SLOTS * ST_SAMPLES_BYTES + 2 * ST_WORK_BYTES + ST_TABLE_BYTES);
    // mbarriers (8 bytes each): samples_ready[3] (256 copy arrivals), slot_free[3] (8 warps), full[2] (8), empty[2] (8)
    const uint32_t bar_ready = bars, bar_free = bars + 24, bar_full = bars + 48, bar_empty = bars + 64;

    const int tid = threadIdx.x, lane = tid & 31;
    const int role = __shfl_sync(0xffffffffu, tid >> 8, 0);       // 0: pass 1, 1: pass 2
    const int gt = tid & 255;
    const int w = __shfl_sync(0xffffffffu, gt >> 5, 0);           // warp index inside the role
    if (tid < 256) {
        s_hann[tid] = adn_c_hann512_half[16 * (tid & 15) + (tid >> 4)];      // [n2][n1]
        s_tw256[tid] = adn_c_tw256[tid >> 4][tid & 15];                      // [n2][k1]
        s_tw512[tid] = adn_c_tw512[(tid >> 4) + 16 * (tid & 15)];            // [a][k2]: W512^(a + 16 k2)
    }
    if (tid == 0) {
        for (int i = 0; i < 3; ++i) { mbar_init(bar_ready + 8 * i, 256); mbar_init(bar_free + 8 * i, 8); }
        for (int i = 0; i < 2; ++i) { mbar_init(bar_full + 8 * i, 8); mbar_init(bar_empty + 8 * i, 8); }
        fence_barrier_init();
    }
    __syncthreads();

    const int total_tiles = (int)n_clips * tiles_per_clip;        // < 2^31 (host check)
    const int pad = center ? ADN_N_FFT / 2 : 0;
    const int n_local = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;      // tiles of this CTA (>= 1)

    if (role == 0) {
        // =========================================================== pass-1 warps
        TileWalk st;                                              // the tile being STAGED (two ahead of the one being transformed)
        st.init((int)blockIdx.x, (int)gridDim.x, tiles_per_clip);
        int st_slot = 0;
        auto stage = [&]() {                                      // samples of the next tile -> slot st_slot, completion on bar_ready
            const int t0 = st.tin * TF;
            const int nf = min(TF, n_frames - t0);
            stft_stage(samples + st_slot * (ST_SAMPLES_BYTES / 4), wave + st.clip * clip_stride, out, t0 * ADN_HOP - pad, length, nf + 3, gt);
            cp_async_mbar_arrive_noinc(bar_ready + 8 * st_slot);
            st.next();
            st_slot = st_slot == WS_SAMPLE_SLOTS - 1 ? 0 : st_slot + 1;
        };
        if (0 < n_local) stage();
        if (1 < n_local) stage();
        int slot = 0;                                             // sample slot of tile i = i % 3
        uint32_t ready_par = 0;                                   // parity of bar_ready[slot] for tile i: (i / 3) & 1
        uint32_t free_bits = 0;                                   // bit s: parity of the NEXT wait on bar_free[s]
        for (int i = 0; i < n_local; ++i) {
            const int buf = i & 1;
            mbar_wait_sleep(bar_ready + 8 * slot, ready_par);
            const float* fr = samples + slot * (ST_SAMPLES_BYTES / 4) + lane * ST_ROW_STRIDE;
            float2 v0[16], v1[16];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int n2 = w + 8 * c;
                const float4* hw = reinterpret_cast<const float4*>(s_hann + n2 * 16);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 hq = hw[j];
                    const float2 s0 = *reinterpret_cast<const float2*>(fr + ((2 * j) >> 2) * ST_ROW_STRIDE + 32 * ((2 * j) & 3) + 2 * n2);
                    const float2 s1 = *reinterpret_cast<const float2*>(fr + ((2 * j + 1) >> 2) * ST_ROW_STRIDE + 32 * ((2 * j + 1) & 3) + 2 * n2);
                    (c ? v1 : v0)[2 * j] = pk_mul(s0, make_float2(hq.x, hq.y));
                    (c ? v1 : v0)[2 * j + 1] = pk_mul(s1, make_float2(hq.z, hq.w));
                }
            }
            // this warp is done with the sample slot (the multiplies above consumed every load)
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_free + 8 * slot);
            // refill: tile i + 2 goes into the slot tile i - 1 used (= st_slot); wait until all eight warps have released it
            if (i + 2 < n_local) {
                if (i >= 1) {
                    mbar_wait_sleep(bar_free + 8 * st_slot, (free_bits >> st_slot) & 1u);
                    free_bits ^= 1u << st_slot;
                }
                stage();
            }
            if (slot == WS_SAMPLE_SLOTS - 1) { slot = 0; ready_par ^= 1u; } else ++slot;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int n2 = w + 8 * c;
                const float4* tq = reinterpret_cast<const float4*>(s_tw256 + n2 * 16);
                float2 tw[16];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 t4 = tq[j];
                    tw[2 * j] = make_float2(t4.x, t4.y); tw[2 * j + 1] = make_float2(t4.z, t4.w);
                }
                if (c == 0) {
                    dft16<false>(v0);
#pragma unroll
                    for (int k1 = 1; k1 < 16; ++k1) v0[k1] = cmul(v0[k1], tw[k1]);
                } else {
                    dft16<false>(v1);
#pragma unroll
                    for (int k1 = 1; k1 < 16; ++k1) v1[k1] = cmul(v1[k1], tw[k1]);
                }
            }
            // exchange array `buf` was last read by pass 2 of tile i - 2
            if (i >= 2) mbar_wait_sleep(bar_empty + 8 * buf, ((i - 2) >> 1) & 1);
            float2* const work = work0 + buf * (ST_WORK_BYTES / 8);
#pragma unroll
            for (int k1 = 0; k1 < 16; ++k1) {
                work[(k1 * 16 + w) * TF + lane] = v0[k1];
                work[(k1 * 16 + w + 8) * TF + lane] = v1[k1];
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_full + 8 * buf);       // release: the warp's stores are ordered before the arrival
        }
        cp_async_wait_all();
    } else {
        // =========================================================== pass-2 warps
        const int a = w, b = (w == 0) ? 8 : 16 - w;
        TileWalk tw_;
        tw_.init((int)blockIdx.x, (int)gridDim.x, tiles_per_clip);
        for (int i = 0; i < n_local; ++i, tw_.next()) {
            const int buf = i & 1;
            const int cur_clip = tw_.clip, cur_t0 = tw_.tin * TF;
            const int nf = min(TF, n_frames - cur_t0);
            mbar_wait_sleep(bar_full + 8 * buf, (i >> 1) & 1);
            const float2* const work = work0 + buf * (ST_WORK_BYTES / 8);
            float2 A[16], B[16];
#pragma unroll
            for (int n2 = 0; n2 < 16; ++n2) A[n2] = work[(a * 16 + n2) * TF + lane];
#pragma unroll
            for (int n2 = 0; n2 < 16; ++n2) B[n2] = work[(b * 16 + n2) * TF + lane];
            dft16<false>(A);                                         // A[k2] = Z[a + 16 k2]
            dft16<false>(B);                                         // B[k2] = Z[b + 16 k2]
            __syncwarp();                                            // both rows are consumed: pass 1 may overwrite the array
            if (lane == 0) mbar_arrive(bar_empty + 8 * buf);
            auto split = [&](float2 tw, float2& zk, float2& zp) {                        // tw = W512^k
                const float2 e = pk_add(zk, make_float2(zp.x, -zp.y));                    // Zk + conj(Zp)
                const float2 o = pk_add(rot_mi(zk), make_float2(zp.y, zp.x));             // -i (Zk - conj(Zp))
                const float2 wo = cmul(o, tw);
                zk = cadd(e, wo);
                zp = csub(e, wo);
            };
            auto mag = [](float2 z) { return sqrt_approx(fmaf(z.x, z.x, z.y * z.y)); };
            const bool live = lane < nf;
            const long long row0 = ((long long)cur_clip * ADN_N_BINS) * n_frames + cur_t0 + lane;
            const long long step = 16LL * n_frames;
            const float4* twa = reinterpret_cast<const float4*>(s_tw512 + a * 16);        // W512^(a + 16 k2), two per load
            if (w != 0) {
                float* pa = out + row0 + (long long)a * n_frames;                         // rows a + 16 k2, upwards
                float* pb = out + row0 + (long long)b * n_frames;                         // rows b + 16 k2 (TFIX) / from k2 = 15 downwards
                if (!TFIX) pb += 15 * step;
#pragma unroll
                for (int j = 0; j < 8; ++j) {                                             // 256 - k = b + 16 (15 - k2)
                    const float4 t4 = twa[j];
                    split(make_float2(t4.x, t4.y), A[2 * j], B[15 - 2 * j]);
                    split(make_float2(t4.z, t4.w), A[2 * j + 1], B[14 - 2 * j]);
                    const float m0 = mag(A[2 * j]), m1 = mag(B[15 - 2 * j]), m2 = mag(A[2 * j + 1]), m3 = mag(B[14 - 2 * j]);
                    if (TFIX) {                                                           // immediate offsets from two base pointers
                        constexpr long long ST = 16LL * TFIX;
                        if (live) {
                            __stcs(pa + (2 * j) * ST, m0);
                            __stcs(pb + (15 - 2 * j) * ST, m1);
                            __stcs(pa + (2 * j + 1) * ST, m2);
                            __stcs(pb + (14 - 2 * j) * ST, m3);
                        }
                    } else {
                        if (live) {
                            __stcs(pa, m0);
                            __stcs(pb, m1);
                            __stcs(pa + step, m2);
                            __stcs(pb - step, m3);
                        }
                        pa += 2 * step;
                        pb -= 2 * step;
                    }
                }
            } else {
                const float4* twb = reinterpret_cast<const float4*>(s_tw512 + 8 * 16);    // W512^(8 + 16 k2)
                const float x0 = fabsf(2.f * (A[0].x + A[0].y)), x256 = fabsf(2.f * (A[0].x - A[0].y));     // bins 0 and 256 are real
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 t4 = twa[j];                                             // a = 0: W512^(16 k2)
                    if (j > 0) split(make_float2(t4.x, t4.y), A[2 * j], A[16 - 2 * j]);
                    split(make_float2(t4.z, t4.w), A[2 * j + 1], A[15 - 2 * j]);
                }
                A[8] = make_float2(2.f * A[8].x, -2.f * A[8].y);     // k = 128 pairs with itself: X[128] = 2 conj(Z[128])
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 t4 = twb[j];
                    split(make_float2(t4.x, t4.y), B[2 * j], B[15 - 2 * j]);
                    split(make_float2(t4.z, t4.w), B[2 * j + 1], B[14 - 2 * j]);
                }
                if (live) {
                    float* p0 = out + row0;                                               // rows 16 k2 (A) and 8 + 16 k2 (B)
                    __stcs(p0, x0);
                    __stcs(p0 + 16 * step, x256);
#pragma unroll
                    for (int k2 = 0; k2 < 16; ++k2) {
                        if (k2 > 0) __stcs(p0 + k2 * step, mag(A[k2]));
                        __stcs(p0 + 8LL * n_frames + k2 * step, mag(B[k2]));
                    }
                }
            }
        }
    }
}

template <bool COMPLEX_OUT, bool CROP>
static int launch_stft(const float* wave, int64_t n_clips, int64_t length, int64_t clip_stride, int center, float* out,
                       float* crop, int f_out, int t_out, cudaStream_t stream) {
    if (n_clips < 0 || length < 0 || clip_stride < length) return ADN_ERR_ARG;
    const int64_t T = adn_stft_num_frames(length, center);
    if (T < 0) return ADN_ERR_SHORT;
    if (n_clips == 0) return ADN_OK;
    if ((!wave && length > 0) || (!out && !CROP)) return ADN_ERR_ARG;      // an empty centred clip still yields one zero frame
    if (CROP && (!crop || f_out <= 0 || t_out <= 0)) return ADN_ERR_ARG;
    if (length >= ((int64_t)1 << 30) || T * ADN_N_BINS >= ((int64_t)1 << 31)) return ADN_ERR_ARG;   // 32-bit per-clip offsets
    int st = check_device();
    if (st != ADN_OK) return st;
    // without the full-magnitude output only the frames that survive the crop are transformed
    const int64_t frames_needed = (CROP && !out) ? (T < t_out ? T : (int64_t)t_out) : T;
    const int tiles_per_clip = (int)((frames_needed + TF - 1) / TF);
    const long long total = (long long)n_clips * tiles_per_clip;
    if (total >= ((int64_t)1 << 31) - 4096) return ADN_ERR_ARG;
    if (CROP && (T < t_out || ADN_N_BINS < f_out))                         // zero padding of data_loader.py:54-72
        ADN_CUDA_TRY(cudaMemsetAsync(crop, 0, (size_t)n_clips * f_out * t_out * sizeof(float), stream));
    static const int impl = getenv("ADN_STFT_IMPL") ? atoi(getenv("ADN_STFT_IMPL")) : 2;        // 1: barrier-phased CTAs, 2: warp-specialised
    if (impl == 2 && !COMPLEX_OUT && !CROP) {
        const long long sms = num_sms();
        const int grid = (int)(total < sms ? total : sms);          // persistent: one 16-warp CTA per SM
#define ADN_WS_LAUNCH(TF_)                                                                                                          \
        do {                                                                                                                        \
            static unsigned char smem_set[64] = {0};                                                                                \
            ADN_CUDA_TRY(ensure_dyn_smem(stft_ws_kernel<TF_>, WS_SMEM_BYTES, smem_set));                                            \
            stft_ws_kernel<TF_><<<grid, WS_THREADS, WS_SMEM_BYTES, stream>>>(wave, n_clips, (int)length, clip_stride, center, (int)T, \
                                                                            tiles_per_clip, out);                                  \
        } while (0)
        if (T == 188) ADN_WS_LAUNCH(188);
        else if (T == 1034) ADN_WS_LAUNCH(1034);
        else if (T == 122) ADN_WS_LAUNCH(122);
        else ADN_WS_LAUNCH(0);
#undef ADN_WS_LAUNCH
        ADN_LAUNCH_CHECK();
        return ADN_OK;
    }
    const size_t smem = ST_SMEM_BYTES;
    auto kern = stft_kernel<COMPLEX_OUT, CROP>;
    static unsigned char smem_set[64] = {0};
    ADN_CUDA_TRY(ensure_dyn_smem(kern, (int)smem, smem_set));
    const long long max_grid = (long long)num_sms() * 2;         // persistent: 2 resident CTAs per SM loop over the tiles
    const int grid = (int)(total < max_grid ? total : max_grid);
    kern<<<grid, STFT_THREADS, smem, stream>>>(wave, n_clips, (int)length, clip_stride, center, (int)T, tiles_per_clip, out, crop, f_out, t_out);
    ADN_LAUNCH_CHECK();
    return ADN_OK;
}

}  // namespace adn

extern "C" int64_t adn_stft_num_frames(int64_t length, int center) {
    if (length < 0) return -1;
    const int64_t padded = length + (center ? ADN_N_FFT : 0);
    if (padded < ADN_N_FFT) return -1;
    return 1 + (padded - ADN_N_FFT) / ADN_HOP;
}

extern "C" int adn_stft_mag_f32(const float* wave, int64_t n_clips, int64_t length, int64_t clip_stride, int center,
                                float* mag, void* stream) {
    return adn::launch_stft<false, false>(wave, n_clips, length, clip_stride, center, mag, nullptr, 0, 0, (cudaStream_t)stream);
}

extern "C" int adn_stft_complex_f32(const float* wave, int64_t n_clips, int64_t length, int64_t clip_stride, int center,
                                    float* spec_c64, void* stream) {
    return adn::launch_stft<true, false>(wave, n_clips, length, clip_stride, center, spec_c64, nullptr, 0, 0, (cudaStream_t)stream);
}

extern "C" int adn_stft_mag_crop_f16_f32(const float* wave, int64_t n_clips, int64_t length, int64_t clip_stride, int center, float* mag,
                                         float* crop, int f_out, int t_out, void* stream) {
    return adn::launch_stft<false, true>(wave, n_clips, length, clip_stride, center, mag, crop, f_out, t_out, (cudaStream_t)stream);
}
