// stft.cu -- fused framing + periodic-Hann window + 512-point real FFT + magnitude (or complex) for sm_100a.
//
// Replaces librosa.stft + librosa.magphase at code/create_train_dataset.py:167-173 (center=False) and
// code/create_test_dataset.py:39-40 (center=True, zero padding) of the reference.
//
// Mapping.  One CTA = 256 threads = 16 half-warps owns a tile of TF = 32 consecutive frames of one clip.
//   1. the (TF-1)*128+512 contiguous samples the tile touches are staged once in shared memory (frames overlap
//      by 75 %, so each sample is reused 4x from smem, read once from HBM), 128-bit loads when aligned, zero
//      fill for the centred edges;
//   2. a half-warp transforms one frame: the 512 real samples are packed as 256 complex points, 16 per lane,
//      register-resident radix-16 x radix-16 with one shared-memory exchange between the passes, then the
//      real-FFT split produces bins k and 256-k together;
//   3. magnitudes are parked in a [257][TF] shared tile and written out as T-contiguous row segments, which is
//      the reference .npy layout (257, T).
// Window and twiddles are compile-time float32 tables (adn_tables.inc), held in registers across tiles.
#include "adn_common.cuh"
#include "adn_tables.inc"

namespace adn {

constexpr int TF = 32;                                   // frames per tile
constexpr int STFT_THREADS = 256;
constexpr int HALF_WARPS = STFT_THREADS / 16;
constexpr int TILE_SAMPLES = (TF - 1) * ADN_HOP + ADN_N_FFT;   // 4480
constexpr int XCH_STRIDE = 17;                           // float2 row stride of the 16x16 exchange (conflict-free)
constexpr int XCH_FLOAT2 = 16 * XCH_STRIDE;              // 272 >= 256 (also holds Z[0..255] for the split)
constexpr int MAG_STRIDE = TF + 2;                       // 34: (2k + t) mod 32 distinct across a warp

template <bool COMPLEX_OUT>
struct StftSmem {
    float samples[TILE_SAMPLES];
    float2 xch[HALF_WARPS][XCH_FLOAT2];
    float tile[COMPLEX_OUT ? 2 : 1][ADN_N_BINS * MAG_STRIDE];   // re (and im) planes, [bin][frame]
};

template <bool COMPLEX_OUT>
__global__ void __launch_bounds__(STFT_THREADS, 2)
stft_kernel(const float* __restrict__ wave, long long n_clips, long long length, long long clip_stride, int center,
            int n_frames, int tiles_per_clip, float* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    StftSmem<COMPLEX_OUT>& sm = *reinterpret_cast<StftSmem<COMPLEX_OUT>*>(smem_raw);

    const int tid = threadIdx.x;
    const int hw = tid >> 4;        // half-warp index 0..15
    const int j = tid & 15;         // lane within the half-warp
    const int warp = tid >> 5, lane = tid & 31;
    const unsigned hmask = 0xFFFFu << (16 * (hw & 1));   // the two half-warps of a warp run independent frame loops

    // thread-constant tables: window at the samples this lane packs, pass-1 twiddles W256^(j*k1)
    float2 win[16];
    float2 tw[16];
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
        win[n1] = make_float2(adn_hann512[32 * n1 + 2 * j], adn_hann512[32 * n1 + 2 * j + 1]);
        tw[n1] = adn_tw256[j][n1];
    }

    const long long total_tiles = n_clips * (long long)tiles_per_clip;
    for (long long tile_id = blockIdx.x; tile_id < total_tiles; tile_id += gridDim.x) {
        const long long clip = tile_id / tiles_per_clip;
        const int t0 = (int)(tile_id % tiles_per_clip) * TF;
        const int nf = min(TF, n_frames - t0);
        const float* __restrict__ src = wave + clip * clip_stride;
        const long long s0 = (long long)t0 * ADN_HOP - (center ? ADN_N_FFT / 2 : 0);
        const int need = (nf - 1) * ADN_HOP + ADN_N_FFT;

        // ---- 1. stage samples [s0, s0 + need) with zero fill outside [0, length)
        const bool vec_ok = ((reinterpret_cast<uintptr_t>(src) & 15) == 0);   // s0 is a multiple of 128
        if (vec_ok) {
            for (int i = tid * 4; i < need; i += STFT_THREADS * 4) {
                const long long g = s0 + i;
                float4 v;
                if (g >= 0 && g + 3 < length) {
                    v = __ldg(reinterpret_cast<const float4*>(src + g));
                } else {
                    v.x = (g >= 0 && g < length) ? src[g] : 0.f;
                    v.y = (g + 1 >= 0 && g + 1 < length) ? src[g + 1] : 0.f;
                    v.z = (g + 2 >= 0 && g + 2 < length) ? src[g + 2] : 0.f;
                    v.w = (g + 3 >= 0 && g + 3 < length) ? src[g + 3] : 0.f;
                }
                *reinterpret_cast<float4*>(&sm.samples[i]) = v;
            }
        } else {
            for (int i = tid; i < need; i += STFT_THREADS) {
                const long long g = s0 + i;
                sm.samples[i] = (g >= 0 && g < length) ? src[g] : 0.f;
            }
        }
        __syncthreads();

        // ---- 2. transforms: half-warp hw takes frames hw, hw+16
        float2* xch = sm.xch[hw];
#pragma unroll 1
        for (int fl = hw; fl < nf; fl += HALF_WARPS) {
            const float* fr = &sm.samples[fl * ADN_HOP];
            float2 v[16];
#pragma unroll
            for (int n1 = 0; n1 < 16; ++n1) {
                const float2 s = *reinterpret_cast<const float2*>(fr + 32 * n1 + 2 * j);
                v[n1] = make_float2(s.x * win[n1].x, s.y * win[n1].y);
            }
            dft16<false>(v);                       // over n1: Y[k1][n2 = j]
#pragma unroll
            for (int k1 = 1; k1 < 16; ++k1) v[k1] = cmul(v[k1], tw[k1]);
#pragma unroll
            for (int k1 = 0; k1 < 16; ++k1) xch[k1 * XCH_STRIDE + j] = v[k1];
            __syncwarp(hmask);
#pragma unroll
            for (int n2 = 0; n2 < 16; ++n2) v[n2] = xch[j * XCH_STRIDE + n2];   // lane j now plays k1 = j
            __syncwarp(hmask);
            dft16<false>(v);                       // over n2: Z[j + 16*k2]
#pragma unroll
            for (int k2 = 0; k2 < 16; ++k2) xch[j + 16 * k2] = v[k2];
            __syncwarp(hmask);
            // real-FFT split: X[k] = E + W512^k O ; conj(X[256-k]) = E - W512^k O
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const int k = j + 16 * m;          // 0..127
                const float2 a = xch[k];
                const float2 b = xch[(256 - k) & 255];
                const float2 e = make_float2(0.5f * (a.x + b.x), 0.5f * (a.y - b.y));
                const float2 o = make_float2(0.5f * (a.y + b.y), 0.5f * (b.x - a.x));   // -i/2 (a - conj b)
                const float2 wo = cmul(o, adn_tw512[k]);
                const float2 xk = cadd(e, wo);
                const float2 xn = csub(e, wo);     // conj of X[256-k]
                if (COMPLEX_OUT) {
                    sm.tile[0][k * MAG_STRIDE + fl] = xk.x;
                    sm.tile[COMPLEX_OUT ? 1 : 0][k * MAG_STRIDE + fl] = xk.y;
                    sm.tile[0][(256 - k) * MAG_STRIDE + fl] = xn.x;
                    sm.tile[COMPLEX_OUT ? 1 : 0][(256 - k) * MAG_STRIDE + fl] = -xn.y;
                } else {
                    sm.tile[0][k * MAG_STRIDE + fl] = sqrtf(fmaf(xk.x, xk.x, xk.y * xk.y));
                    sm.tile[0][(256 - k) * MAG_STRIDE + fl] = sqrtf(fmaf(xn.x, xn.x, xn.y * xn.y));
                }
            }
            if (j == 0) {                          // k = 128 pairs with itself: X[128] = conj(Z[128])
                const float2 a = xch[128];
                if (COMPLEX_OUT) {
                    sm.tile[0][128 * MAG_STRIDE + fl] = a.x;
                    sm.tile[COMPLEX_OUT ? 1 : 0][128 * MAG_STRIDE + fl] = -a.y;
                } else {
                    sm.tile[0][128 * MAG_STRIDE + fl] = sqrtf(fmaf(a.x, a.x, a.y * a.y));
                }
            }
            __syncwarp(hmask);
        }
        __syncthreads();

        // ---- 3. write the tile: one warp per bin row, lanes along frames (T-contiguous output)
        if (COMPLEX_OUT) {
            float2* __restrict__ dst = reinterpret_cast<float2*>(out) + (clip * ADN_N_BINS) * (long long)n_frames + t0;
            for (int f = warp; f < ADN_N_BINS; f += STFT_THREADS / 32)
                if (lane < nf)
                    __stcs(dst + (long long)f * n_frames + lane,
                           make_float2(sm.tile[0][f * MAG_STRIDE + lane], sm.tile[COMPLEX_OUT ? 1 : 0][f * MAG_STRIDE + lane]));
        } else {
            float* __restrict__ dst = out + (clip * ADN_N_BINS) * (long long)n_frames + t0;
            for (int f = warp; f < ADN_N_BINS; f += STFT_THREADS / 32)
                if (lane < nf) __stcs(dst + (long long)f * n_frames + lane, sm.tile[0][f * MAG_STRIDE + lane]);
        }
        __syncthreads();
    }
}

template <bool COMPLEX_OUT>
static int launch_stft(const float* wave, int64_t n_clips, int64_t length, int64_t clip_stride, int center, float* out,
                       cudaStream_t stream) {
    if (n_clips < 0 || length < 0 || clip_stride < length) return ADN_ERR_ARG;
    const int64_t T = adn_stft_num_frames(length, center);
    if (T < 0) return ADN_ERR_SHORT;
    if (n_clips == 0) return ADN_OK;
    if ((!wave && length > 0) || !out) return ADN_ERR_ARG;      // an empty centred clip still yields one zero frame
    if (T > (int64_t)1 << 30) return ADN_ERR_ARG;
    int st = check_device();
    if (st != ADN_OK) return st;
    const int tiles_per_clip = (int)((T + TF - 1) / TF);
    const long long total = (long long)n_clips * tiles_per_clip;
    const size_t smem = sizeof(StftSmem<COMPLEX_OUT>);
    auto kern = stft_kernel<COMPLEX_OUT>;
    static unsigned char smem_set[64] = {0};
    ADN_CUDA_TRY(ensure_dyn_smem(kern, (int)smem, smem_set));
    const long long max_grid = (long long)num_sms() * 2 * 8;     // 2 resident CTAs per SM, 8 waves before looping
    const int grid = (int)(total < max_grid ? total : max_grid);
    kern<<<grid, STFT_THREADS, smem, stream>>>(wave, n_clips, length, clip_stride, center, (int)T, tiles_per_clip, out);
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
    return adn::launch_stft<false>(wave, n_clips, length, clip_stride, center, mag, (cudaStream_t)stream);
}

extern "C" int adn_stft_complex_f32(const float* wave, int64_t n_clips, int64_t length, int64_t clip_stride, int center,
                                    float* spec_c64, void* stream) {
    return adn::launch_stft<true>(wave, n_clips, length, clip_stride, center, spec_c64, (cudaStream_t)stream);
}
