// istft.cu -- fused (magnitude x phasor) + 512-point inverse real FFT + Hann window + gather-form overlap-add
// + division by the window sum-of-squares + centre trim, for sm_100a.
//
// Replaces `magnitude_spectrogram * angles` and librosa.istft(hop_length=128) in the reference's
// griffin_lim_reconstruction (code/test.py:36-37,40,48).
//
// Mapping ("lane = frame", the mirror image of stft.cu).  One CTA = 8 warps owns S = 29 output hops (3 712 samples) of
// one clip and inverse-transforms the S + 3 = 32 frames that touch them; lane t of every warp works on frame tA + t:
//   1. spectrum rows are read straight from global memory into registers -- 32 consecutive frames of one bin are one
//      coalesced 128-byte (magnitude) / 256-byte (phasor) request, which is exactly the reference (257, T) layout;
//      warp w holds the bin columns k = 16 i1 + w and 16 i1 + (16 - w) (0 and 8 for warp 0), i.e. bins k and 256 - k
//      of the Hermitian pre-pass sit in the SAME thread;
//   2. pass 1 (radix-16 over i1, conjugate W256 twiddles) -> [o1][i2][frame] work array -> pass 2 (radix-16 over i2)
//      -> window / 512 -> the frame's 512 samples are parked in a [frame][514]-padded buffer that aliases the work array;
//   3. every output sample gathers its <= 4 contributions in ascending frame order (deterministic, no atomics),
//      divides by sum(w^2) over the same frames; a warp stores 256 contiguous bytes.
// Tables (window, twiddles, w^2) are shared-memory arrays read with warp-uniform (broadcast) addresses.
#include <cstdlib>
#include "tc_common.cuh"
#include "adn_tables.inc"

namespace adn {

constexpr int IS_FRAMES = 32;                 // frames per tile = lanes per warp
constexpr int IS_HOPS = IS_FRAMES - 3;        // 29 output hops per tile
constexpr int IS_THREADS = 256;
constexpr int IS_YSTRIDE = 514;               // floats per frame in the sample buffer: 514 mod 32 = 2 -> lane-strided float2 stores hit 32 banks
constexpr int IS_WORK_BYTES = IS_FRAMES * IS_YSTRIDE * 4;      // 65 792 >= 256 * 32 * 8 (work array)
constexpr int IS_TABLE_BYTES = 3 * 256 * 8 + 512 * 4 + 128 * 4;
constexpr int IS_SMEM_BYTES = IS_WORK_BYTES + IS_TABLE_BYTES;

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
// per-clip key of the counter-based phase generator (test.py:36 draws the phase from the unseeded numpy RNG)
__device__ __forceinline__ uint32_t phase_key(unsigned long long seed, long long clip) {
    return mix32((uint32_t)seed ^ mix32((uint32_t)(seed >> 32) * 0x9E3779B9u + (uint32_t)clip) ^ (uint32_t)((unsigned long long)clip >> 32));
}
// 16 random phase bits of bin `row`, frame t of a clip.  One 32-bit hash serves two bins 16 rows apart (rows c + 16 i1 with
// i1 = 2j, 2j+1), which the iSTFT kernel holds in the same thread.
__device__ __forceinline__ uint32_t phase_bits(uint32_t key, uint32_t c, uint32_t i1, uint32_t n_frames, uint32_t t) {   // row = c + 16 i1
    const uint32_t h = mix32(((c + 16u * (i1 >> 1)) * n_frames + t) ^ key);
    return (i1 & 1u) ? (h >> 16) : (h & 0xffffu);
}
__device__ __forceinline__ float2 phasor_from_bits(uint32_t bits16) {
    const float ang = (float)((int)bits16 - 32768) * (6.283185307179586f / 65536.0f);   // [-pi, pi)
    float s, c;
    __sincosf(ang, &s, &c);
    return make_float2(c, s);
}

template <typename T>
__device__ __forceinline__ const T* is_row_ptr(const T* base, int stride_bytes, int k) {
    long long r;
    asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(r) : "r"(stride_bytes), "r"(k), "l"(reinterpret_cast<long long>(base)));
    return reinterpret_cast<const T*>(r);
}

// MODE 0: spec = mag * phasor ; MODE 1: spec = phasor array itself (complex spectrogram) ; MODE 2: mag * random phasor ;
// MODE 3: spec = mag * C / |C| for an arbitrary complex array C (the phase of another spectrogram, e.g. the noisy input's:
//         SURVEY 8f row 4, an opt-in that is NOT what the reference computes)
template <int MODE>
__global__ void __launch_bounds__(IS_THREADS, 2)
istft_kernel(const float* __restrict__ mag, const float2* __restrict__ ph, unsigned long long seed,
             const unsigned long long* __restrict__ seed_counter, long long n_clips, int n_frames, int tiles_per_clip,
             float* __restrict__ audio) {
    // a device-side call counter added to the seed: a CUDA-graph replay bakes `seed` in, the counter (advanced by
    // adn_u64_add inside the same graph) still gives every replay a fresh phase, as test.py:36 draws one per call
    if (MODE == 2 && seed_counter != nullptr) seed += *seed_counter;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* const work = reinterpret_cast<float2*>(smem_raw);            // [o1*16 + i2][frame]
    float* const ybuf = reinterpret_cast<float*>(smem_raw);              // [frame][514], aliases `work`
    float2* const s_win = reinterpret_cast<float2*>(smem_raw + IS_WORK_BYTES);   // hann[2n], hann[2n+1], each / 512
    float2* const s_tw256 = s_win + 256;                                 // conj W256^(i2 o1), [i2][o1]
    float2* const s_tw512 = s_win + 512;                                 // W512^k, k < 256
    float* const s_w2 = reinterpret_cast<float*>(s_win + 768);           // hann^2 [512]
    float* const s_wss = s_w2 + 512;                                     // 1 / (interior sum of the four w^2 terms) [128]

    const int tid = threadIdx.x, lane = tid & 31;
    const int w = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int out_len = ADN_HOP * (n_frames - 1);

    s_win[tid] = make_float2(adn_c_hann512[2 * tid] * (1.0f / 512.0f), adn_c_hann512[2 * tid + 1] * (1.0f / 512.0f));
    {
        const float2 t = adn_c_tw256[tid >> 4][tid & 15];
        s_tw256[tid] = make_float2(t.x, -t.y);
    }
    s_tw512[tid] = adn_c_tw512[tid];
    for (int i = tid; i < ADN_N_FFT; i += IS_THREADS) { const float h = adn_c_hann512[i]; s_w2[i] = h * h; }
    if (tid < 128) {                                                     // ascending frame order = descending offset
        float acc = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) { const float h = adn_c_hann512[384 - 128 * q + tid]; acc += h * h; }
        s_wss[tid] = 1.0f / acc;
    }

    const int a = w, b = (w == 0) ? 8 : 16 - w;
    const int total_tiles = (int)n_clips * tiles_per_clip;               // < 2^31 (host check)

    // Raw spectrum columns a and b of this lane's frame, fetched one tile AHEAD (the loads are issued before the gather
    // phase of the previous tile and consumed after it, so their latency hides behind it).  Dead lanes (frame outside
    // [0, T)) load a clamped, in-range address and are zeroed by `vf`: no branches around the loads.
    float rm[(MODE == 1) ? 1 : 33];
    float2 rp[(MODE == 2) ? 1 : 33];
    auto prefetch = [&](int tile) {
        if (tile >= total_tiles) return;
        const int clip = tile / tiles_per_clip;
        const int t = (tile - clip * tiles_per_clip) * IS_HOPS - 1 + lane;
        const int tc = (t >= 0 && t < n_frames) ? t : 0;
        const long long base = (long long)clip * ADN_N_BINS * n_frames + tc;   // 257 * T < 2^30 (host check)
#pragma unroll
        for (int i1 = 0; i1 < 16; ++i1) {
            if (MODE != 1) {
                rm[i1] = __ldcs(is_row_ptr(mag + base + a * n_frames, 16 * n_frames * 4, i1));
                rm[16 + i1] = __ldcs(is_row_ptr(mag + base + b * n_frames, 16 * n_frames * 4, i1));
            }
            if (MODE != 2) {
                rp[i1] = __ldcs(is_row_ptr(ph + base + a * n_frames, 16 * n_frames * 8, i1));
                rp[16 + i1] = __ldcs(is_row_ptr(ph + base + b * n_frames, 16 * n_frames * 8, i1));
            }
        }
        if (w == 0) {
            if (MODE != 1) rm[32] = __ldcs(is_row_ptr(mag + base, 16 * n_frames * 4, 16));
            if (MODE != 2) rp[32] = __ldcs(is_row_ptr(ph + base, 16 * n_frames * 8, 16));
        }
    };

    // Only the magnitude-only mode (33 registers) is fetched ahead; with an explicit phasor the raw tile is 99 registers,
    // which cannot stay live across the gather without spilling, so modes 0 / 1 load at the top of their own tile.
    constexpr bool AHEAD = (MODE == 2);
    if (AHEAD) prefetch(blockIdx.x);
    for (int tile_id = blockIdx.x; tile_id < total_tiles; tile_id += gridDim.x) {
        if (!AHEAD) prefetch(tile_id);
        const int clip = tile_id / tiles_per_clip;
        const int h0 = (tile_id - clip * tiles_per_clip) * IS_HOPS;      // first output hop
        const int tA = h0 - 1;                                           // first frame of the tile (may be -1)
        const int t = tA + lane;
        const bool valid = (t >= 0 && t < n_frames);
        const float vf = valid ? 1.f : 0.f;
        const int tc = valid ? t : 0;
        const bool interior = (tA >= 0) && (tA + IS_FRAMES <= n_frames);   // uniform: every frame of the tile exists

        // ---- 1. spectrum columns a and b of frame t: X = mag * phasor
        float2 A[16], B[16];
        float x256 = 0.f;                                                // Re X[256] (warp 0)
        {
            uint32_t key = 0;
            if (MODE == 2) key = phase_key(seed, clip);
            auto spec = [&](int j, int c, int i1) -> float2 {            // raw slot j = element (c + 16 i1) * T + t
                if (MODE == 1) return rp[j];
                if (MODE == 3) {                                          // phase of rp[j]; a zero bin keeps phase 0
                    const float2 q = rp[j];
                    const float n2 = fmaf(q.x, q.x, q.y * q.y);
                    const float s = n2 > 0.f ? rm[j] * rsqrtf(n2) : 0.f;
                    return n2 > 0.f ? make_float2(q.x * s, q.y * s) : make_float2(rm[j], 0.f);
                }
                const float2 p = (MODE == 0) ? rp[j] : phasor_from_bits(phase_bits(key, (uint32_t)c, (uint32_t)i1, (uint32_t)n_frames, (uint32_t)tc));
                return make_float2(rm[j] * p.x, rm[j] * p.y);
            };
#pragma unroll
            for (int i1 = 0; i1 < 16; ++i1) { A[i1] = spec(i1, a, i1); B[i1] = spec(16 + i1, b, i1); }
            if (w == 0) x256 = spec(32, 0, 16).x;
            if (!interior) {                                             // dead lanes (frames outside [0, T)) contribute zeros
#pragma unroll
                for (int i1 = 0; i1 < 16; ++i1) { A[i1].x *= vf; A[i1].y *= vf; B[i1].x *= vf; B[i1].y *= vf; }
                x256 *= vf;
            }
        }

        // ---- Hermitian pre-pass in place (unscaled: the 1/2 lives in the window table).  For a pair (k, 256-k):
        //      E = X[k] + conj X[256-k], D = X[k] - conj X[256-k], O = D conj(W512^k);
        //      Z[k] = E + iO, Z[256-k] = conj(E) + i conj(O)
        auto pre = [&](int k, float2& xk, float2& xp) {
            const float2 e = make_float2(xk.x + xp.x, xk.y - xp.y);
            const float2 d = make_float2(xk.x - xp.x, xk.y + xp.y);
            const float2 o = cmul_conj(d, s_tw512[k]);
            xk = make_float2(e.x - o.y, e.y + o.x);
            xp = make_float2(e.x + o.y, o.x - e.y);
        };
        if (w != 0) {
#pragma unroll
            for (int i1 = 0; i1 < 16; ++i1) pre(16 * i1 + a, A[i1], B[15 - i1]);      // 256 - k = 16 (15 - i1) + b
        } else {
            A[0] = make_float2(A[0].x + x256, A[0].x - x256);            // k = 0 with 256: c2r ignores Im(DC), Im(Nyquist)
#pragma unroll
            for (int i1 = 1; i1 < 8; ++i1) pre(16 * i1, A[i1], A[16 - i1]);
            A[8] = make_float2(2.f * A[8].x, -2.f * A[8].y);             // k = 128 pairs with itself: Z = 2 conj(X) (unscaled)
#pragma unroll
            for (int i1 = 0; i1 < 8; ++i1) pre(16 * i1 + 8, B[i1], B[15 - i1]);
        }

        // ---- pass 1: inverse radix-16 over i1 for columns i2 = a, b; conjugate W256 twiddles; -> work[o1][i2][frame]
        dft16<true>(A);
        dft16<true>(B);
#pragma unroll
        for (int o1 = 1; o1 < 16; ++o1) {
            A[o1] = cmul(A[o1], s_tw256[a * 16 + o1]);
            B[o1] = cmul(B[o1], s_tw256[b * 16 + o1]);
        }
        __syncthreads();                                                 // the previous tile's gather has finished with ybuf
#pragma unroll
        for (int o1 = 0; o1 < 16; ++o1) {
            work[(o1 * 16 + a) * IS_FRAMES + lane] = A[o1];
            work[(o1 * 16 + b) * IS_FRAMES + lane] = B[o1];
        }
        __syncthreads();

        // ---- pass 2: rows o1 = w, w + 8 -> z[o1 + 16 o2] = x[2n] + i x[2n+1]; window
#pragma unroll
        for (int i2 = 0; i2 < 16; ++i2) {
            A[i2] = work[(w * 16 + i2) * IS_FRAMES + lane];
            B[i2] = work[((w + 8) * 16 + i2) * IS_FRAMES + lane];
        }
        dft16<true>(A);
        dft16<true>(B);
        __syncthreads();                                                 // every warp has read its rows: `work` may become `ybuf`
        {
            float* yrow = ybuf + lane * IS_YSTRIDE;
#pragma unroll
            for (int o2 = 0; o2 < 16; ++o2) {
                const int n0 = w + 16 * o2, n1 = n0 + 8;
                const float2 w0 = s_win[n0], w1 = s_win[n1];
                *reinterpret_cast<float2*>(yrow + 2 * n0) = make_float2(A[o2].x * w0.x, A[o2].y * w0.y);
                *reinterpret_cast<float2*>(yrow + 2 * n1) = make_float2(B[o2].x * w1.x, B[o2].y * w1.y);
            }
        }
        __syncthreads();
        if (AHEAD) prefetch(tile_id + gridDim.x);                        // in flight during the gather below (A, B are dead here)

        // ---- 3. gather overlap-add: sample n = 128*h + r, padded position p = n + 256 = 128*(h+2) + r,
        //         contributions from frames t = h-1 .. h+2 at offsets 384+r, 256+r, 128+r, r (dead frames hold zeros)
        const int hops = min(IS_HOPS, (n_frames - 1) - h0);
        float* __restrict__ dst = audio + (long long)clip * out_len + h0 * ADN_HOP;
        {
            const int hq = tid >> 6, r = (tid & 63) * 2;                 // this thread: hops hq, hq+4, ... ; offset r (even)
            const float* yb = ybuf + hq * IS_YSTRIDE + 384 + r;
            float* d = dst + hq * ADN_HOP + r;
            if (interior) {
                const float2 inv = *reinterpret_cast<const float2*>(&s_wss[r]);     // 1 / sum of the four w^2 terms
                for (int hl = hq; hl < hops; hl += 4, yb += 4 * IS_YSTRIDE, d += 4 * ADN_HOP) {
                    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {                        // frame slot hl + q  <->  t = h0 + hl - 1 + q
                        const float2 y = *reinterpret_cast<const float2*>(yb + q * (IS_YSTRIDE - 128));
                        acc.x += y.x; acc.y += y.y;
                    }
                    __stcs(reinterpret_cast<float2*>(d), make_float2(acc.x * inv.x, acc.y * inv.y));
                }
            } else {
                for (int hl = hq; hl < hops; hl += 4, yb += 4 * IS_YSTRIDE, d += 4 * ADN_HOP) {
                    float2 acc = make_float2(0.f, 0.f), wss = make_float2(0.f, 0.f);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float2 y = *reinterpret_cast<const float2*>(yb + q * (IS_YSTRIDE - 128));
                        acc.x += y.x; acc.y += y.y;
                        const int tq = tA + hl + q;
                        if (tq >= 0 && tq < n_frames) {
                            const float2 ww = *reinterpret_cast<const float2*>(&s_w2[384 - 128 * q + r]);
                            wss.x += ww.x; wss.y += ww.y;
                        }
                    }
                    const float tiny = 1.17549435e-38f;
                    __stcs(reinterpret_cast<float2*>(d), make_float2(wss.x > tiny ? acc.x / wss.x : acc.x, wss.y > tiny ? acc.y / wss.y : acc.y));
                }
            }
        }
        // the next tile's first barrier (before its work[] stores) orders this gather against them
    }
}

// ------------------------------------------------------------------------------------------------ warp-specialised variant
// istft_kernel's two CTAs per SM run the same phases in step (spectrum loads -> pre-pass / pass 1 -> barrier -> pass 2 -> barrier ->
// gather): in the explicit-phasor modes every warp of the SM first waits for its 64 uncached loads and then computes, so neither the
// issue port nor HBM is busy for more than half of the time (0.745 ms where HBM alone needs 0.43 and the issue port 0.37).
// istft_ws_kernel is one CTA of 16 warps per SM with no CTA-wide barrier inside the tile loop:
//   warps 0-7  ("pass 1"): spectrum columns a, b of frame lane -> X = mag * phasor -> Hermitian pre-pass -> radix-16 over i1 ->
//              exchange array [o1][i2][frame] (two arrays, mbarrier full / empty, one arrival per warp).  The magnitudes of the NEXT
//              tile are already in flight: every warp copies its own 33 rows with cp.async into a private staging area (each lane
//              reads back only what it copied itself: no synchronisation, zero fill for frames outside [0, T)).
//   warps 8-15 ("pass 2"): rows o1 = w, w + 8 -> conjugate W256 twiddles -> radix-16 over i2 -> window -> overlap-add IN REGISTERS:
//              lane = frame, and the four segments that meet in one output hop sit in four neighbouring lanes at the same register
//              index (sample 2 (o1 + 16 o2) + {0,1} has offset r = 2 (o1 + 16 (o2 mod 4)) in segment o2 / 4), so the gather is three
//              shuffles and three adds per value, in ascending frame order like the gather of istft_kernel -- the [frame][514]
//              sample buffer and its two CTA-wide barriers disappear, and the exchange array is released right after it was read.
//              The 29 x 128 output samples of a tile are transposed through a small double-buffered staging tile so that every
//              global store is a 256-byte run; a warp drains tile i - 1 after it has delivered tile i, so no warp waits for another.
constexpr int IW_THREADS = 512;
constexpr int IW_WORK_BYTES = 256 * IS_FRAMES * 8;
constexpr int IW_OSTRIDE = 130;                                       // floats per hop row: lane-strided float2 stores hit 32 banks
constexpr int IW_OBUF_BYTES = (IS_HOPS * IW_OSTRIDE * 4 + 15) / 16 * 16;
constexpr int IW_STAGE_FLOATS = 33 * 32;                              // per pass-1 warp: 33 magnitude rows x 32 frames
constexpr int IW_STAGE_BYTES = 8 * IW_STAGE_FLOATS * 4;
constexpr int IW_TABLE_BYTES = 3 * 256 * 8 + 4 * 128 * 4;
// Shared-memory plan per mode.  The seeded-phase mode (2) has no phasor input: two exchange arrays, magnitude staging only.  The
// modes that read a complex array (0, 1, 3) stage it too (8 B per bin: 66 KB per tile) and pay for that with ONE exchange array --
// pass 1 holds tile i + 1 in registers until pass 2 has read tile i, which costs nothing when the two roles take equally long.
__host__ __device__ constexpr int iw_nbuf(int mode) { return mode == 2 ? 2 : 1; }
__host__ __device__ constexpr int iw_mag_bytes(int mode) { return mode == 1 ? 0 : IW_STAGE_BYTES; }
__host__ __device__ constexpr int iw_ph_bytes(int mode) { return mode == 2 ? 0 : 2 * IW_STAGE_BYTES; }
__host__ __device__ constexpr int iw_smem_bytes(int mode) {
    return iw_nbuf(mode) * IW_WORK_BYTES + 2 * IW_OBUF_BYTES + iw_mag_bytes(mode) + iw_ph_bytes(mode) + IW_TABLE_BYTES + 64;
}

// mbarrier wait with a short sleep between polls: a polling warp issues SYNCS + BRA pairs that compete for the issue port with the
// warps it is waiting for (measured: 15 % of all issued instructions of the first version were polls).  Bounded like mbar_wait_sleep.
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity) {
#pragma unroll 1
    for (int spin = 0; spin < (1 << 22); ++spin) {
        uint32_t ok;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
        __nanosleep(64);
    }
    __trap();
}

__device__ __forceinline__ void is_cp_async4(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void is_cp_async8(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}

template <int MODE, int TFIX>
__global__ void __launch_bounds__(IW_THREADS, 1)
istft_ws_kernel(const float* __restrict__ mag, const float2* __restrict__ ph, unsigned long long seed,
                const unsigned long long* __restrict__ seed_counter, long long n_clips, int n_frames_rt, int tiles_per_clip,
                float* __restrict__ audio) {
    const int n_frames = TFIX ? TFIX : n_frames_rt;
    if (MODE == 2 && seed_counter != nullptr) seed += *seed_counter;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NBUF = iw_nbuf(MODE);
    constexpr int OFF_OBUF = NBUF * IW_WORK_BYTES, OFF_MAG = OFF_OBUF + 2 * IW_OBUF_BYTES, OFF_PH = OFF_MAG + iw_mag_bytes(MODE),
                  OFF_TAB = OFF_PH + iw_ph_bytes(MODE);
    float2* const work0 = reinterpret_cast<float2*>(smem_raw);                                  // [buf][o1*16 + i2][frame]
    float* const obuf0 = reinterpret_cast<float*>(smem_raw + OFF_OBUF);                         // [buf][hop][130]
    float* const stage = reinterpret_cast<float*>(smem_raw + OFF_MAG);                          // [warp][row slot][frame] magnitudes
    float2* const stage_ph = reinterpret_cast<float2*>(smem_raw + OFF_PH);                      // [warp][row slot][frame] complex input
    float2* const s_win = reinterpret_cast<float2*>(smem_raw + OFF_TAB);
    float2* const s_tw256 = s_win + 256;                                 // conj W256^(o1 i2), [o1][i2]
    float2* const s_tw512 = s_win + 512;                                 // W512^(a + 16 i1), [a][i1]
    float* const s_inv = reinterpret_cast<float*>(s_win + 768);          // [sel][r]: 1 / sum of the w^2 terms that exist
    const uint32_t bars = smem_u32(s_inv + 4 * 128);
    const uint32_t bar_full = bars, bar_empty = bars + 16, bar_ofull = bars + 32, bar_oempty = bars + 48;

    const int tid = threadIdx.x, lane = tid & 31;
    const int role = __shfl_sync(0xffffffffu, tid >> 8, 0);              // 0: pass 1, 1: pass 2
    const int w = __shfl_sync(0xffffffffu, (tid & 255) >> 5, 0);         // warp index inside the role
    if (tid < 256) {
        const int o1 = tid >> 4, o2 = tid & 15, n = o1 + 16 * o2;        // [o1][o2]: window at samples 2n, 2n + 1, each / 512
        s_win[tid] = make_float2(adn_c_hann512[2 * n] * (1.0f / 512.0f), adn_c_hann512[2 * n + 1] * (1.0f / 512.0f));
        const float2 t = adn_c_tw256[tid >> 4][tid & 15];
        s_tw256[tid] = make_float2(t.x, -t.y);
        s_tw512[tid] = adn_c_tw512[16 * (tid & 15) + (tid >> 4)];
    } else if (tid < 384) {
        // sel bit 0: the hop is the clip's first (frame h - 1 does not exist); bit 1: its last (frame h + 2 does not exist).
        // Ascending frame order = descending window offset, as in the gather of istft_kernel.
        const int r = tid - 256;
#pragma unroll
        for (int sel = 0; sel < 4; ++sel) {
            float acc = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if ((q == 0 && (sel & 1)) || (q == 3 && (sel & 2))) continue;
                const float h = adn_c_hann512[384 - 128 * q + r];
                acc += h * h;
            }
            s_inv[sel * 128 + r] = 1.0f / acc;
        }
    }
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_full + 8 * i, 8); mbar_init(bar_empty + 8 * i, 8);
            mbar_init(bar_ofull + 8 * i, 8); mbar_init(bar_oempty + 8 * i, 8);
        }
        fence_barrier_init();
    }
    __syncthreads();

    const int total_tiles = (int)n_clips * tiles_per_clip;               // < 2^31 (host check)
    const int n_local = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;     // tiles of this CTA (>= 1)
    TileWalk tw;
    tw.init((int)blockIdx.x, (int)gridDim.x, tiles_per_clip);

    if (role == 0) {
        // =========================================================== pass-1 warps
        const int a = w, b = (w == 0) ? 8 : 16 - w;
        const long long rs = 16LL * n_frames;                            // elements between rows k and k + 16
        float* const stg = stage + w * IW_STAGE_FLOATS + lane;           // slot j of this lane: stg[32 j]
        const uint32_t stg_u32 = smem_u32(stg);
        float2* const stp = stage_ph + w * IW_STAGE_FLOATS + lane;
        const uint32_t stp_u32 = smem_u32(stp);
        // columns a, b (+ row 256) of frame `lane` of a tile -> this warp's staging slots; frames outside [0, T) are zero-filled
        auto stage_in = [&](int clip, int tin) {
            const int t = tin * IS_HOPS - 1 + lane;
            const bool ok = t >= 0 && t < n_frames;
            const long long e0 = ((long long)clip * ADN_N_BINS + a) * n_frames + (ok ? t : 0);
            if (MODE != 1) {
                const int nb = ok ? 4 : 0;
                const float* pa = mag + e0;
                const float* pb = pa + (long long)(b - a) * n_frames;
#pragma unroll
                for (int i1 = 0; i1 < 16; ++i1) {
                    is_cp_async4(stg_u32 + i1 * 128, pa + i1 * rs, nb);
                    is_cp_async4(stg_u32 + (16 + i1) * 128, pb + i1 * rs, nb);
                }
                if (w == 0) is_cp_async4(stg_u32 + 32 * 128, pa + 16 * rs, nb);
            }
            if (MODE != 2) {
                const int nb = ok ? 8 : 0;
                const float2* qa = ph + e0;
                const float2* qb = qa + (long long)(b - a) * n_frames;
#pragma unroll
                for (int i1 = 0; i1 < 16; ++i1) {
                    is_cp_async8(stp_u32 + i1 * 256, qa + i1 * rs, nb);
                    is_cp_async8(stp_u32 + (16 + i1) * 256, qb + i1 * rs, nb);
                }
                if (w == 0) is_cp_async8(stp_u32 + 32 * 256, qa + 16 * rs, nb);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        stage_in(tw.clip, tw.tin);
        for (int i = 0; i < n_local; ++i) {
            const int buf = (NBUF == 2) ? (i & 1) : 0;
            const int clip = tw.clip, h0 = tw.tin * IS_HOPS;
            tw.next();                                                   // now the tile after this one
            const int tA = h0 - 1, t = tA + lane;
            const bool valid = (t >= 0 && t < n_frames);
            const int tc = valid ? t : 0;

            // ---- spectrum columns a and b of frame t: X = mag * phasor
            float2 A[16], B[16];
            float x256 = 0.f;
            {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                float2 rp[(MODE == 2) ? 1 : 33];
                if (MODE != 2) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) rp[j] = stp[32 * j];
                    if (w == 0) rp[32] = stp[32 * 32];
                }
                float rm[(MODE == 1) ? 1 : 33];
                if (MODE != 1) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) rm[j] = stg[32 * j];
                    if (w == 0) rm[32] = stg[32 * 32];
                }
                uint32_t key = 0;
                if (MODE == 2) key = phase_key(seed, clip);
                auto spec = [&](int j, int c, int i1) -> float2 {        // raw slot j = element (c + 16 i1) * T + t
                    if (MODE == 1) return rp[j];
                    if (MODE == 3) {                                      // phase of rp[j]; a zero bin keeps phase 0
                        const float2 q = rp[j];
                        const float n2 = fmaf(q.x, q.x, q.y * q.y);
                        const float s = n2 > 0.f ? rm[j] * rsqrtf(n2) : 0.f;
                        return n2 > 0.f ? make_float2(q.x * s, q.y * s) : make_float2(rm[j], 0.f);
                    }
                    const float2 p = (MODE == 0) ? rp[j] : phasor_from_bits(phase_bits(key, (uint32_t)c, (uint32_t)i1, (uint32_t)n_frames, (uint32_t)tc));
                    return make_float2(__fmul_rn(rm[j], p.x), __fmul_rn(rm[j], p.y));    // never contracted: modes 0 and 2 agree bit for bit
                };
#pragma unroll
                for (int i1 = 0; i1 < 16; ++i1) { A[i1] = spec(i1, a, i1); B[i1] = spec(16 + i1, b, i1); }
                if (w == 0) x256 = spec(32, 0, 16).x;
            }
            // the staging slots have been read back: the next tile's input streams in behind the arithmetic below
            if (i + 1 < n_local) stage_in(tw.clip, tw.tin);

            // ---- Hermitian pre-pass in place (see istft_kernel)
            const float4* t5a = reinterpret_cast<const float4*>(s_tw512 + a * 16);
            auto pre = [&](float2 twk, float2& xk, float2& xp) {
                const float2 e = make_float2(xk.x + xp.x, xk.y - xp.y);
                const float2 d = make_float2(xk.x - xp.x, xk.y + xp.y);
                const float2 o = cmul_conj(d, twk);
                xk = make_float2(e.x - o.y, e.y + o.x);
                xp = make_float2(e.x + o.y, o.x - e.y);
            };
            if (w != 0) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {                            // 256 - k = 16 (15 - i1) + b
                    const float4 q = t5a[j];
                    pre(make_float2(q.x, q.y), A[2 * j], B[15 - 2 * j]);
                    pre(make_float2(q.z, q.w), A[2 * j + 1], B[14 - 2 * j]);
                }
            } else {
                const float4* t5b = reinterpret_cast<const float4*>(s_tw512 + 8 * 16);
                A[0] = make_float2(A[0].x + x256, A[0].x - x256);        // k = 0 with 256: c2r ignores Im(DC), Im(Nyquist)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 q = t5a[j];                             // a = 0: W512^(16 i1)
                    if (j > 0) pre(make_float2(q.x, q.y), A[2 * j], A[16 - 2 * j]);
                    pre(make_float2(q.z, q.w), A[2 * j + 1], A[15 - 2 * j]);
                }
                A[8] = make_float2(2.f * A[8].x, -2.f * A[8].y);         // k = 128 pairs with itself: Z = 2 conj(X) (unscaled)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 q = t5b[j];
                    pre(make_float2(q.x, q.y), B[2 * j], B[15 - 2 * j]);
                    pre(make_float2(q.z, q.w), B[2 * j + 1], B[14 - 2 * j]);
                }
            }

            // ---- pass 1: inverse radix-16 over i1 for columns i2 = a, b (the W256 twiddles are applied by the pass-2 warps)
            dft16<true>(A);
            dft16<true>(B);
            if (i >= NBUF) mbar_wait_backoff(bar_empty + 8 * buf, ((i - NBUF) / NBUF) & 1);   // pass 2 of tile i - NBUF has read the array
            float2* const wk = work0 + buf * (IW_WORK_BYTES / 8) + lane;
#pragma unroll
            for (int o1 = 0; o1 < 16; ++o1) {
                wk[(o1 * 16 + a) * IS_FRAMES] = A[o1];
                wk[(o1 * 16 + b) * IS_FRAMES] = B[o1];
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_full + 8 * buf);              // release: the warp's stores are ordered before the arrival
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else {
        // =========================================================== pass-2 warps
        const int out_len = ADN_HOP * (n_frames - 1);
        const int last_hop = n_frames - 2;
        // 1 / (w^2 sum) of an interior hop for this thread's eight output pairs: r = 2 (w + 16 m) and 2 (w + 8 + 16 m)
        float2 invA[4], invB[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            invA[m] = *reinterpret_cast<const float2*>(&s_inv[2 * (w + 16 * m)]);
            invB[m] = *reinterpret_cast<const float2*>(&s_inv[2 * (w + 8 + 16 * m)]);
        }
        int p_clip = 0, p_h0 = 0;                                        // the tile delivered last (drained one iteration later)
        auto drain = [&](int j) {                                        // staging tile of local tile j -> 256-byte runs of global stores
            const int ob = j & 1;
            mbar_wait_backoff(bar_ofull + 8 * ob, (j >> 1) & 1);
            const int hops = min(IS_HOPS, (n_frames - 1) - p_h0);
            const float* src = obuf0 + ob * (IW_OBUF_BYTES / 4) + 2 * lane;
            float* dst = audio + (long long)p_clip * out_len + p_h0 * ADN_HOP + 2 * lane;
            for (int idx = w; idx < 2 * hops; idx += 8) {                // idx = 2 hop + half
                const int row = idx >> 1, half = idx & 1;
                const float2 v = *reinterpret_cast<const float2*>(src + row * IW_OSTRIDE + half * 64);
                __stcs(reinterpret_cast<float2*>(dst + row * ADN_HOP + half * 64), v);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_oempty + 8 * ob);
        };
        for (int i = 0; i < n_local; ++i, tw.next()) {
            const int buf = (NBUF == 2) ? (i & 1) : 0, obf = i & 1;
            const int clip = tw.clip, h0 = tw.tin * IS_HOPS;
            mbar_wait_backoff(bar_full + 8 * buf, (i / NBUF) & 1);
            const float2* const wk = work0 + buf * (IW_WORK_BYTES / 8) + lane;
            float2 A[16], B[16];
#pragma unroll
            for (int i2 = 0; i2 < 16; ++i2) A[i2] = wk[(w * 16 + i2) * IS_FRAMES];
#pragma unroll
            for (int i2 = 0; i2 < 16; ++i2) B[i2] = wk[((w + 8) * 16 + i2) * IS_FRAMES];
            {                                                            // conjugate W256^(o1 i2) twiddles, two per broadcast load
                const float4* ta = reinterpret_cast<const float4*>(s_tw256 + w * 16);
                const float4* tb = reinterpret_cast<const float4*>(s_tw256 + (w + 8) * 16);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 qa = ta[j], qb = tb[j];
                    if (j > 0) { A[2 * j] = cmul(A[2 * j], make_float2(qa.x, qa.y)); B[2 * j] = cmul(B[2 * j], make_float2(qb.x, qb.y)); }
                    A[2 * j + 1] = cmul(A[2 * j + 1], make_float2(qa.z, qa.w));
                    B[2 * j + 1] = cmul(B[2 * j + 1], make_float2(qb.z, qb.w));
                }
            }
            __syncwarp();                                                // both rows are in registers: pass 1 may overwrite the array
            if (lane == 0) mbar_arrive(bar_empty + 8 * buf);
            dft16<true>(A);                                              // A[o2] = z[w + 16 o2] = x[2n] + i x[2n+1]
            dft16<true>(B);
            {
                const float4* wa = reinterpret_cast<const float4*>(s_win + w * 16);
                const float4* wb = reinterpret_cast<const float4*>(s_win + (w + 8) * 16);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 qa = wa[j], qb = wb[j];
                    A[2 * j] = make_float2(A[2 * j].x * qa.x, A[2 * j].y * qa.y);
                    A[2 * j + 1] = make_float2(A[2 * j + 1].x * qa.z, A[2 * j + 1].y * qa.w);
                    B[2 * j] = make_float2(B[2 * j].x * qb.x, B[2 * j].y * qb.y);
                    B[2 * j + 1] = make_float2(B[2 * j + 1].x * qb.z, B[2 * j + 1].y * qb.w);
                }
            }
            // ---- overlap-add across lanes: hop h0 + lane = frame slots lane .. lane + 3 with segments 3 .. 0 (ascending frame order)
            const bool interior = (h0 >= 1) && (h0 + IS_FRAMES - 1 <= n_frames);
            const int hop = h0 + lane;
            const int sel = (hop == 0 ? 1 : 0) + (hop == last_hop ? 2 : 0);
            float2 oA[4], oB[4];
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                float2 sa = A[12 + m], sb = B[12 + m];
#pragma unroll
                for (int q = 1; q < 4; ++q) {
                    sa.x += __shfl_down_sync(0xffffffffu, A[12 - 4 * q + m].x, q);
                    sa.y += __shfl_down_sync(0xffffffffu, A[12 - 4 * q + m].y, q);
                    sb.x += __shfl_down_sync(0xffffffffu, B[12 - 4 * q + m].x, q);
                    sb.y += __shfl_down_sync(0xffffffffu, B[12 - 4 * q + m].y, q);
                }
                float2 ia = invA[m], ib = invB[m];
                if (!interior) {
                    ia = *reinterpret_cast<const float2*>(&s_inv[sel * 128 + 2 * (w + 16 * m)]);
                    ib = *reinterpret_cast<const float2*>(&s_inv[sel * 128 + 2 * (w + 8 + 16 * m)]);
                }
                oA[m] = make_float2(sa.x * ia.x, sa.y * ia.y);
                oB[m] = make_float2(sb.x * ib.x, sb.y * ib.y);
            }
            if (i >= 2) mbar_wait_backoff(bar_oempty + 8 * obf, ((i - 2) >> 1) & 1);   // every warp has drained tile i - 2
            if (lane < IS_HOPS) {                                        // rows past the clip's last hop are never drained
                float* const ob = obuf0 + obf * (IW_OBUF_BYTES / 4) + lane * IW_OSTRIDE + 2 * w;
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    *reinterpret_cast<float2*>(ob + 32 * m) = oA[m];
                    *reinterpret_cast<float2*>(ob + 32 * m + 16) = oB[m];
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_ofull + 8 * obf);
            if (i >= 1) drain(i - 1);
            p_clip = clip; p_h0 = h0;
        }
        drain(n_local - 1);
    }
}

__global__ void random_phasor_kernel(unsigned long long seed, long long n_clips, int n_frames, float2* __restrict__ out) {
    const int per_clip = ADN_N_BINS * n_frames;
    const long long total = n_clips * per_clip;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long clip = i / per_clip;
        const int e = (int)(i - clip * per_clip), row = e / n_frames, t = e - row * n_frames;
        out[i] = phasor_from_bits(phase_bits(phase_key(seed, clip), (uint32_t)(row & 15), (uint32_t)(row >> 4), (uint32_t)n_frames, (uint32_t)t));
    }
}

__global__ void u64_add_kernel(unsigned long long* counter, unsigned long long inc) { *counter += inc; }

static int launch_istft(const float* mag, const float* phasor, int spec_is_complex, uint64_t seed, const uint64_t* seed_counter,
                        int64_t n_clips, int64_t n_frames, float* audio, cudaStream_t stream) {
    if (n_clips < 0 || n_frames < 1 || n_frames * ADN_N_BINS >= ((int64_t)1 << 31) / 2) return ADN_ERR_ARG;   // 32-bit per-clip byte offsets
    if (n_clips == 0 || n_frames == 1) return ADN_OK;          // hop*(T-1) = 0 samples
    if (!audio) return ADN_ERR_ARG;
    if (spec_is_complex == 2 ? (!phasor || !mag) : spec_is_complex ? !phasor : !mag) return ADN_ERR_ARG;
    if ((reinterpret_cast<uintptr_t>(audio) & 15) != 0) return ADN_ERR_ARG;
    int st = check_device();
    if (st != ADN_OK) return st;
    const int hops_total = (int)n_frames - 1;
    const int tiles_per_clip = (hops_total + IS_HOPS - 1) / IS_HOPS;
    const long long total = (long long)n_clips * tiles_per_clip;
    if (total >= ((int64_t)1 << 31) - 4096) return ADN_ERR_ARG;
    const float2* ph = reinterpret_cast<const float2*>(phasor);
    const int mode = spec_is_complex == 2 ? 3 : spec_is_complex ? 1 : phasor ? 0 : 2;
    // 1: barrier-phased CTAs (istft_kernel); 2: warp-specialised (istft_ws_kernel); 3 (default): warp-specialised for the seeded-phase
    // mode, barrier-phased for the modes that read a complex array (their 66 KB of input per tile can only be staged with 8-byte
    // cp.async, which costs the load/store unit more than the direct loads of istft_kernel: measured 0.82 vs 0.745 ms at config 2)
    static const int impl = getenv("ADN_ISTFT_IMPL") ? atoi(getenv("ADN_ISTFT_IMPL")) : 3;
    if (impl == 2 || (impl == 3 && mode == 2)) {
        const long long sms = num_sms();
        const int grid = (int)(total < sms ? total : sms);     // persistent: one 16-warp CTA per SM
#define ADN_IW_LAUNCH(MODE, TF_)                                                                                              \
        do {                                                                                                                  \
            static unsigned char smem_set[64] = {0};                                                                          \
            ADN_CUDA_TRY(ensure_dyn_smem(istft_ws_kernel<MODE, TF_>, iw_smem_bytes(MODE), smem_set));                         \
            istft_ws_kernel<MODE, TF_><<<grid, IW_THREADS, iw_smem_bytes(MODE), stream>>>(                                    \
                mag, ph, (unsigned long long)seed, reinterpret_cast<const unsigned long long*>(seed_counter), n_clips,        \
                (int)n_frames, tiles_per_clip, audio);                                                                        \
        } while (0)
#define ADN_IW_MODE(MODE)                                                                                                     \
        do {                                                                                                                  \
            if (n_frames == 188) ADN_IW_LAUNCH(MODE, 188);                                                                    \
            else if (n_frames == 1034) ADN_IW_LAUNCH(MODE, 1034);                                                             \
            else ADN_IW_LAUNCH(MODE, 0);                                                                                      \
        } while (0)
        if (mode == 0) ADN_IW_MODE(0);
        else if (mode == 1) ADN_IW_MODE(1);
        else if (mode == 2) ADN_IW_MODE(2);
        else ADN_IW_MODE(3);
#undef ADN_IW_MODE
#undef ADN_IW_LAUNCH
        ADN_LAUNCH_CHECK();
        return ADN_OK;
    }
    const size_t smem = IS_SMEM_BYTES;
    const long long max_grid = (long long)num_sms() * 2;       // persistent: 2 resident CTAs per SM loop over the tiles
    const int grid = (int)(total < max_grid ? total : max_grid);
#define ADN_ISTFT_LAUNCH(MODE)                                                                                     \
    do {                                                                                                           \
        static unsigned char smem_set[64] = {0};                                                                   \
        ADN_CUDA_TRY(ensure_dyn_smem(istft_kernel<MODE>, (int)smem, smem_set));                                    \
        istft_kernel<MODE><<<grid, IS_THREADS, smem, stream>>>(mag, ph, (unsigned long long)seed,                  \
                                                               reinterpret_cast<const unsigned long long*>(seed_counter), n_clips,  \
                                                               (int)n_frames, tiles_per_clip, audio);              \
    } while (0)
    if (spec_is_complex == 2) ADN_ISTFT_LAUNCH(3);
    else if (spec_is_complex) ADN_ISTFT_LAUNCH(1);
    else if (phasor) ADN_ISTFT_LAUNCH(0);
    else ADN_ISTFT_LAUNCH(2);
#undef ADN_ISTFT_LAUNCH
    ADN_LAUNCH_CHECK();
    return ADN_OK;
}

}  // namespace adn

extern "C" int adn_istft_ola_f32(const float* mag, const float* phasor_c64, int spec_is_complex, uint64_t seed,
                                 int64_t n_clips, int64_t n_frames, float* audio, void* stream) {
    return adn::launch_istft(mag, phasor_c64, spec_is_complex, seed, nullptr, n_clips, n_frames, audio, (cudaStream_t)stream);
}

extern "C" int adn_istft_ola_counter_f32(const float* mag, uint64_t seed, const uint64_t* seed_counter_dev, int64_t n_clips,
                                         int64_t n_frames, float* audio, void* stream) {
    if (!seed_counter_dev) return ADN_ERR_ARG;
    return adn::launch_istft(mag, nullptr, 0, seed, seed_counter_dev, n_clips, n_frames, audio, (cudaStream_t)stream);
}

extern "C" int adn_u64_add(uint64_t* counter_dev, uint64_t inc, void* stream) {
    if (!counter_dev) return ADN_ERR_ARG;
    int st = adn::check_device();
    if (st != ADN_OK) return st;
    adn::u64_add_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(reinterpret_cast<unsigned long long*>(counter_dev), (unsigned long long)inc);
    ADN_LAUNCH_CHECK();
    return ADN_OK;
}

extern "C" int adn_random_phasor_c64(uint64_t seed, int64_t n_clips, int64_t n_frames, float* phasor_c64, void* stream) {
    if (n_clips < 0 || n_frames < 1 || n_frames * ADN_N_BINS >= ((int64_t)1 << 31) / 2) return ADN_ERR_ARG;
    if (n_clips == 0) return ADN_OK;
    if (!phasor_c64) return ADN_ERR_ARG;
    int st = adn::check_device();
    if (st != ADN_OK) return st;
    const long long total = n_clips * n_frames * ADN_N_BINS;
    const long long want = (total + 255) / 256, cap = (long long)adn::num_sms() * 16;
    adn::random_phasor_kernel<<<(int)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(
        (unsigned long long)seed, n_clips, (int)n_frames, reinterpret_cast<float2*>(phasor_c64));
    ADN_LAUNCH_CHECK();
    return ADN_OK;
}
