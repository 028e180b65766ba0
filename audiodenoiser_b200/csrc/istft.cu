// istft.cu -- fused (magnitude x phasor) + 512-point inverse real FFT + Hann window + gather-form overlap-add
// + division by the window sum-of-squares + centre trim, for sm_100a.
//
// Replaces `magnitude_spectrogram * angles` and librosa.istft(hop_length=128) in the reference's
// griffin_lim_reconstruction (code/test.py:36-37,40,48).
//
// Mapping.  One CTA = 256 threads owns S = 29 output hops (3712 samples) of one clip and inverse-transforms the
// S + 3 = 32 frames that touch them:
//   1. rows of 32 consecutive frames are read T-contiguously (coalesced) for all 257 bins, multiplied by the
//      phasor, and transposed into per-frame slots in shared memory;
//   2. a half-warp inverse-transforms one frame (Hermitian pre-pass -> radix-16 x radix-16 -> window / 256),
//      writing the 512 windowed samples back over the frame's own slot;
//   3. every output sample gathers its <= 4 contributions in ascending frame order (deterministic, no atomics),
//      divides by sum(w^2) over the same frames and is stored with 128-bit writes.
#include "adn_common.cuh"
#include "adn_tables.inc"

namespace adn {

constexpr int IS_FRAMES = 32;                 // frames per tile
constexpr int IS_HOPS = IS_FRAMES - 3;        // 29 output hops per tile
constexpr int IS_THREADS = 256;
constexpr int IS_HALF_WARPS = IS_THREADS / 16;
constexpr int SLOT = 257;                     // float2 per frame slot: 514 words == 2 (mod 32) -> conflict-free transpose
constexpr int IXCH_STRIDE = 17;
constexpr int IXCH_FLOAT2 = 16 * IXCH_STRIDE;

struct IstftSmem {
    float2 slot[IS_FRAMES][SLOT];             // spectrum in, then 512 windowed samples out (2056 B >= 2048 B)
    float2 xch[IS_HALF_WARPS][IXCH_FLOAT2];
    float w2[ADN_N_FFT];
};

__device__ __forceinline__ float2 random_phasor(unsigned long long seed, unsigned long long idx) {
    // splitmix64 finaliser as a counter-based generator; 24 random bits -> phase in [0, 1) turns
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (idx + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    const float u = (float)(unsigned)(z >> 40) * (1.0f / 16777216.0f);
    float s, c;
    sincospif(2.0f * u, &s, &c);
    return make_float2(c, s);
}

// MODE 0: spec = mag * phasor ; MODE 1: spec = phasor array itself (complex spectrogram) ; MODE 2: mag * random phasor
template <int MODE>
__global__ void __launch_bounds__(IS_THREADS, 2)
istft_kernel(const float* __restrict__ mag, const float2* __restrict__ ph, unsigned long long seed,
             long long n_clips, int n_frames, int tiles_per_clip, float* __restrict__ audio) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    IstftSmem& sm = *reinterpret_cast<IstftSmem*>(smem_raw);
    const int tid = threadIdx.x, hw = tid >> 4, j = tid & 15, warp = tid >> 5, lane = tid & 31;
    const int out_len = ADN_HOP * (n_frames - 1);
    const unsigned hmask = 0xFFFFu << (16 * (hw & 1));   // half-warps skip invalid frames independently

    for (int i = tid; i < ADN_N_FFT; i += IS_THREADS) { const float w = adn_hann512[i]; sm.w2[i] = w * w; }

    float2 win[16], tw[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        // output samples 2*(j+16m), +1 ; the irfft 1/512 normalisation = 1/256 on the packed transform
        win[m] = make_float2(adn_hann512[32 * m + 2 * j] * (1.0f / 256.0f), adn_hann512[32 * m + 2 * j + 1] * (1.0f / 256.0f));
        const float2 t = adn_tw256[j][m];
        tw[m] = make_float2(t.x, -t.y);       // inverse transform: conjugate twiddles
    }

    const long long total_tiles = n_clips * (long long)tiles_per_clip;
    for (long long tile_id = blockIdx.x; tile_id < total_tiles; tile_id += gridDim.x) {
        const long long clip = tile_id / tiles_per_clip;
        const int h0 = (int)(tile_id % tiles_per_clip) * IS_HOPS;      // first output hop
        const int tA = h0 - 1;                                         // first frame of the tile (may be -1)
        const long long base = clip * (long long)ADN_N_BINS * n_frames;

        // ---- 1. load + transpose: warp per bin row, lanes along frames
        {
            const int t = tA + lane;
            const bool valid = (t >= 0 && t < n_frames);
            for (int f = warp; f < ADN_N_BINS; f += IS_THREADS / 32) {
                float2 x = make_float2(0.f, 0.f);
                if (valid) {
                    const long long g = base + (long long)f * n_frames + t;
                    if (MODE == 1) {
                        x = __ldcs(ph + g);
                    } else {
                        const float m = __ldcs(mag + g);
                        const float2 p = (MODE == 0) ? __ldcs(ph + g) : random_phasor(seed, (unsigned long long)g);
                        x = make_float2(m * p.x, m * p.y);
                    }
                    if (f == 0 || f == 256) x.y = 0.f;                 // c2r: Im(DC), Im(Nyquist) are ignored
                }
                sm.slot[lane][f] = x;
            }
        }
        __syncthreads();

        // ---- 2. inverse transforms: half-warp hw takes frame slots hw, hw+16
        float2* xch = sm.xch[hw];
#pragma unroll 1
        for (int fl = hw; fl < IS_FRAMES; fl += IS_HALF_WARPS) {
            const int t = tA + fl;
            if (t < 0 || t >= n_frames) continue;                      // uniform across the half-warp
            float2* X = sm.slot[fl];
            float2 v[16];
            // Hermitian pre-pass: Z[k] = E + iO with E = (X[k] + conj X[256-k])/2, O = conj(W512^k) (X[k] - conj X[256-k])/2
#pragma unroll
            for (int n1 = 0; n1 < 16; ++n1) {
                const int k = 16 * n1 + j;
                const float2 a = X[k];
                const float2 b = X[256 - k];
                const float2 e = make_float2(0.5f * (a.x + b.x), 0.5f * (a.y - b.y));
                const float2 d = make_float2(0.5f * (a.x - b.x), 0.5f * (a.y + b.y));
                const float2 o = cmul_conj(d, adn_tw512[k]);
                v[n1] = make_float2(e.x - o.y, e.y + o.x);             // E + i*O
            }
            __syncwarp(hmask);                                              // all lanes have read X before it is overwritten
            dft16<true>(v);
#pragma unroll
            for (int k1 = 1; k1 < 16; ++k1) v[k1] = cmul(v[k1], tw[k1]);
#pragma unroll
            for (int k1 = 0; k1 < 16; ++k1) xch[k1 * IXCH_STRIDE + j] = v[k1];
            __syncwarp(hmask);
#pragma unroll
            for (int n2 = 0; n2 < 16; ++n2) v[n2] = xch[j * IXCH_STRIDE + n2];
            __syncwarp(hmask);
            dft16<true>(v);                                            // z[j + 16*m] = x[2(j+16m)] + i x[2(j+16m)+1]
#pragma unroll
            for (int m = 0; m < 16; ++m) X[j + 16 * m] = make_float2(v[m].x * win[m].x, v[m].y * win[m].y);
        }
        __syncthreads();

        // ---- 3. gather overlap-add: sample n = 128*h + r, padded position p = n + 256 = 128*(h+2) + r,
        //         contributions from frames t = h-1 .. h+2 at offsets 384+r, 256+r, 128+r, r
        const int hops = min(IS_HOPS, (n_frames - 1) - h0);
        float* __restrict__ dst = audio + clip * (long long)out_len + (long long)h0 * ADN_HOP;
        for (int i = tid * 4; i < hops * ADN_HOP; i += IS_THREADS * 4) {
            const int hl = i >> 7, r = i & 127;                        // local hop, offset within hop (multiple of 4)
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f), wss = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int q = 0; q < 4; ++q) {                              // frame slot hl + q  <->  t = h0 + hl - 1 + q
                const int t = tA + hl + q;
                if (t >= 0 && t < n_frames) {
                    const int off = 384 - 128 * q + r;
                    const float2* ys = &sm.slot[hl + q][off >> 1];     // slots are only 8-byte aligned (257 float2)
                    const float2 y0 = ys[0], y1 = ys[1];
                    const float4 w = *reinterpret_cast<const float4*>(&sm.w2[off]);
                    acc.x += y0.x; acc.y += y0.y; acc.z += y1.x; acc.w += y1.y;
                    wss.x += w.x; wss.y += w.y; wss.z += w.z; wss.w += w.w;
                }
            }
            const float tiny = 1.17549435e-38f;
            float4 o;
            o.x = wss.x > tiny ? acc.x / wss.x : acc.x;
            o.y = wss.y > tiny ? acc.y / wss.y : acc.y;
            o.z = wss.z > tiny ? acc.z / wss.z : acc.z;
            o.w = wss.w > tiny ? acc.w / wss.w : acc.w;
            __stcs(reinterpret_cast<float4*>(dst + i), o);
        }
        __syncthreads();
    }
}

static int launch_istft(const float* mag, const float* phasor, int spec_is_complex, uint64_t seed, int64_t n_clips,
                        int64_t n_frames, float* audio, cudaStream_t stream) {
    if (n_clips < 0 || n_frames < 1 || n_frames > ((int64_t)1 << 23)) return ADN_ERR_ARG;
    if (n_clips == 0 || n_frames == 1) return ADN_OK;          // hop*(T-1) = 0 samples
    if (!audio) return ADN_ERR_ARG;
    if (spec_is_complex ? !phasor : !mag) return ADN_ERR_ARG;
    if ((reinterpret_cast<uintptr_t>(audio) & 15) != 0) return ADN_ERR_ARG;
    int st = check_device();
    if (st != ADN_OK) return st;
    const int hops_total = (int)n_frames - 1;
    const int tiles_per_clip = (hops_total + IS_HOPS - 1) / IS_HOPS;
    const long long total = (long long)n_clips * tiles_per_clip;
    const size_t smem = sizeof(IstftSmem);
    const long long max_grid = (long long)num_sms() * 2 * 8;
    const int grid = (int)(total < max_grid ? total : max_grid);
    const float2* ph = reinterpret_cast<const float2*>(phasor);
#define ADN_ISTFT_LAUNCH(MODE)                                                                                     \
    do {                                                                                                           \
        static unsigned char smem_set[64] = {0};                                                                   \
        ADN_CUDA_TRY(ensure_dyn_smem(istft_kernel<MODE>, (int)smem, smem_set));                                    \
        istft_kernel<MODE><<<grid, IS_THREADS, smem, stream>>>(mag, ph, (unsigned long long)seed, n_clips, (int)n_frames, \
                                                               tiles_per_clip, audio);                             \
    } while (0)
    if (spec_is_complex) ADN_ISTFT_LAUNCH(1);
    else if (phasor) ADN_ISTFT_LAUNCH(0);
    else ADN_ISTFT_LAUNCH(2);
#undef ADN_ISTFT_LAUNCH
    ADN_LAUNCH_CHECK();
    return ADN_OK;
}

}  // namespace adn

extern "C" int adn_istft_ola_f32(const float* mag, const float* phasor_c64, int spec_is_complex, uint64_t seed,
                                 int64_t n_clips, int64_t n_frames, float* audio, void* stream) {
    return adn::launch_istft(mag, phasor_c64, spec_is_complex, seed, n_clips, n_frames, audio, (cudaStream_t)stream);
}
