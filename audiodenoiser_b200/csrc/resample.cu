// resample.cu -- mono mix-down + polyphase FIR sample-rate conversion in front of the STFT: the numeric part of
// `librosa.load(path, sr=8000)` after decoding (code/create_train_dataset.py:204,225, code/create_test_dataset.py:139,
// code/test.py:80: every script resamples to SAMPLE_RATE = 8000 on load), so that native-rate clips (44.1 kHz IRMAS / UrbanSound)
// can enter the device path directly.
//
// librosa 0.10's default resampler is soxr_hq -- a third-party library that is neither in the reference tree nor in this image, and
// whose filter the reference does not specify.  This kernel implements the published polyphase algorithm of
// scipy.signal.resample_poly (zero-stuff by `up`, Kaiser-windowed sinc low-pass, keep every `down`-th sample, centre-aligned), which
// the oracle (oracle/resample_oracle.py) calls in float64: parity is exact against that algorithm and only approximate against soxr.
//
//   y[m] = sum_{t < T} xm[j(m) - t] * P[r(m)][t]      c = (m + pre_remove) * down,   j = c / up,   r = c % up
//   P[r][t] = g[r + t * up]  (g = the zero-padded, up-scaled prototype filter; built on the host, audiodenoiser_b200/resample.py)
//   xm = mean over the input channels (librosa.to_mono before the resample)
//
// One CTA per 256 consecutive output samples of a clip: the phase table (up x T floats, 35 KB for 441 -> 80) is loaded once per
// persistent CTA, the input window of a tile (256 * down / up + T samples, mono-mixed while loading) once per tile; the T-tap dot
// products then run out of shared memory.  HBM traffic = 4 B per input sample and channel + 4 B per output sample.
#include "adn_common.cuh"

namespace adn {

constexpr int RS_TILE = 256;

__global__ void __launch_bounds__(RS_TILE)
resample_poly_kernel(const float* __restrict__ x, int channels, long long len_in, int up, int down, const float* __restrict__ table,
                     int taps, int taps_pitch, long long pre_remove, long long len_out, long long tiles_per_clip, long long n_tiles,
                     int window, float* __restrict__ y) {
    extern __shared__ float s_dyn[];
    float* s_tab = s_dyn;                                  // [up][taps_pitch]
    float* s_x = s_dyn + (size_t)up * taps_pitch;           // [window]
    for (int i = threadIdx.x; i < up * taps_pitch; i += RS_TILE) s_tab[i] = table[i];
    const float inv_c = 1.f / (float)channels;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long clip = tile / tiles_per_clip;
        const long long m0 = (tile - clip * tiles_per_clip) * RS_TILE;
        // input window of this tile: j(m0) - (taps_pitch - 1) .. j(m0 + RS_TILE - 1); the table rows are zero-padded to taps_pitch
        // (a multiple of 4), so the dot product runs over whole float4 groups of taps
        const long long j_first = ((m0 + pre_remove) * down) / up - (taps_pitch - 1);
        const float* xc = x + clip * channels * len_in;
        __syncthreads();                                   // previous tile's window (and, first time, the table) is done with
        for (int i = threadIdx.x; i < window; i += RS_TILE) {
            const long long j = j_first + i;
            float v = 0.f;
            if (j >= 0 && j < len_in) {
                for (int c = 0; c < channels; ++c) v += xc[(long long)c * len_in + j];
                v *= inv_c;
            }
            s_x[i] = v;
        }
        __syncthreads();
        const long long m = m0 + threadIdx.x;
        if (m < len_out) {
            const long long c = (m + pre_remove) * down;
            const long long j = c / up;
            const int r = (int)(c - j * up);
            const float* p = s_tab + (size_t)r * taps_pitch;
            const float* xs = s_x + (j - j_first);           // xs[-t] = xm[j - t]
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            for (int t = 0; t < taps_pitch; t += 4) {          // one LDS.128 of taps per four scalar window reads
                const float4 p4 = *reinterpret_cast<const float4*>(p + t);
                a0 = fmaf(xs[-t], p4.x, a0);
                a1 = fmaf(xs[-t - 1], p4.y, a1);
                a2 = fmaf(xs[-t - 2], p4.z, a2);
                a3 = fmaf(xs[-t - 3], p4.w, a3);
            }
            y[clip * len_out + m] = (a0 + a1) + (a2 + a3);
        }
    }
}

}  // namespace adn

using namespace adn;

// x: (n_clips, channels, len_in) float32 planar; table: (up, taps_pitch) float32 polyphase filter (device, 16-byte aligned, rows
// zero-padded to taps_pitch = a multiple of 4); y: (n_clips, len_out).
// len_out = ceil(len_in * up / down); pre_remove = the number of leading full-rate output samples scipy discards for alignment.
extern "C" int adn_resample_poly_f32(const float* x, int64_t n_clips, int channels, int64_t len_in, int up, int down,
                                     const float* table, int taps, int taps_pitch, int64_t pre_remove, int64_t len_out, float* y,
                                     void* stream) {
    if (n_clips < 0 || channels <= 0 || len_in <= 0 || up <= 0 || down <= 0 || taps <= 0 || taps_pitch < taps || (taps_pitch & 3) ||
        pre_remove < 0 || len_out <= 0)
        return ADN_ERR_ARG;
    if (n_clips == 0) return ADN_OK;
    if (!x || !table || !y || (reinterpret_cast<uintptr_t>(table) & 15)) return ADN_ERR_ARG;
    int st = check_device(); if (st != ADN_OK) return st;
    const long long window = ((long long)(RS_TILE - 1) * down) / up + taps_pitch + 2;
    const long long smem = ((long long)up * taps_pitch + window) * 4;
    if (smem > 200 * 1024) return ADN_ERR_ARG;              // ratios far outside audio practice (the table must fit in shared memory)
    const long long tiles_per_clip = (len_out + RS_TILE - 1) / RS_TILE;
    const long long n_tiles = tiles_per_clip * n_clips;
    static unsigned char smem_set[64] = {0};
    ADN_CUDA_TRY(ensure_dyn_smem(resample_poly_kernel, 200 * 1024, smem_set));
    const long long cap = (long long)num_sms() * 4;
    const unsigned grid = (unsigned)(n_tiles < cap ? n_tiles : cap);
    resample_poly_kernel<<<grid, RS_TILE, (size_t)smem, (cudaStream_t)stream>>>(x, channels, len_in, up, down, table, taps, taps_pitch,
                                                                           pre_remove, len_out, tiles_per_clip, n_tiles, (int)window, y);
    ADN_LAUNCH_CHECK();
    return ADN_OK;
}
