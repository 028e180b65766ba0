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
#include "adn_common.cuh"
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
    const size_t smem = IS_SMEM_BYTES;
    const long long max_grid = (long long)num_sms() * 2;       // persistent: 2 resident CTAs per SM loop over the tiles
    const int grid = (int)(total < max_grid ? total : max_grid);
    const float2* ph = reinterpret_cast<const float2*>(phasor);
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
