// noise.cu -- the arithmetic of the reference's add_noise (code/create_train_dataset.py:105-159, duplicated at
// code/create_test_dataset.py:43-133) on the device, batched over clips, so that the dataset-creation path
// (clean chunk -> noisy chunk -> |STFT|) never leaves the GPU.
//
//   "white" / "urban" (:138-157)   noise scaled to the requested SNR against the clean chunk:
//                                  clean_rms = sqrt(mean(clean^2) + 1e-12), noise_rms likewise, scale = clean_rms / 10^(snr/20) / noise_rms
//                                  (noise dropped when noise_rms <= 1e-9), out = clip(clean + scale * noise, -1, 1)
//   "noise_cancellation" (:123-135) per 2 s block, with probability 0.8 the first half is attenuated: out = clip(clean - 0.8 clean, -1, 1)
// The random draws (np.random.randn, the snippet start, random.random() per block) stay on the host, in the reference's order, and
// are passed in; the reverb branch (pedalboard / JUCE) is out of scope.  One CTA per clip: two fp64 sums, then the mix -- HBM-bound on
// 12 bytes per sample (the second read of a 64 KB chunk hits L2).
#include "adn_common.cuh"

namespace adn {

__global__ void __launch_bounds__(256)
mix_noise_snr_kernel(const float* __restrict__ clean, const float* __restrict__ noise, long long length, float inv_snr_linear,
                     float* __restrict__ out) {
    __shared__ double red[2][8];
    __shared__ float s_scale;
    const float* c = clean + blockIdx.x * length;
    const float* n = noise + blockIdx.x * length;
    double sc = 0.0, sn = 0.0;
    for (long long i = threadIdx.x; i < length; i += blockDim.x) {
        const float a = c[i], b = n[i];
        sc += (double)a * a; sn += (double)b * b;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sc += __shfl_xor_sync(0xffffffffu, sc, o); sn += __shfl_xor_sync(0xffffffffu, sn, o); }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = sc; red[1][threadIdx.x >> 5] = sn; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int k = 0; k < 8; ++k) { a += red[0][k]; b += red[1][k]; }
        const double clean_rms = sqrt(a / (double)length + 1e-12), noise_rms = sqrt(b / (double)length + 1e-12);
        s_scale = noise_rms > 1e-9 ? (float)(clean_rms * (double)inv_snr_linear / noise_rms) : 0.f;
    }
    __syncthreads();
    const float scale = s_scale;
    float* o = out + blockIdx.x * length;
    for (long long i = threadIdx.x; i < length; i += blockDim.x)
        o[i] = fminf(fmaxf(fmaf(n[i], scale, c[i]), -1.f), 1.f);
}

__global__ void __launch_bounds__(256)
mix_noise_cancel_kernel(const float* __restrict__ clean, const unsigned char* __restrict__ flags, long long length, int block, int half,
                        int blocks_per_clip, float factor, float* __restrict__ out) {
    const float* c = clean + blockIdx.x * length;
    const unsigned char* f = flags + (long long)blockIdx.x * blocks_per_clip;
    float* o = out + blockIdx.x * length;
    for (long long i = threadIdx.x; i < length; i += blockDim.x) {
        const long long b = i / block;
        const bool hit = f[b] && (i - b * block) < half;
        const float v = c[i];
        o[i] = fminf(fmaxf(hit ? fmaf(factor, v, v) : v, -1.f), 1.f);
    }
}

}  // namespace adn

using namespace adn;

extern "C" int adn_mix_noise_snr_f32(const float* clean, const float* noise, int64_t n_clips, int64_t length, float snr_db, float* out,
                                     void* stream) {
    if (n_clips < 0 || length <= 0) return ADN_ERR_ARG;
    if (n_clips == 0) return ADN_OK;
    if (!clean || !noise || !out || n_clips > 0x7fffffffLL) return ADN_ERR_ARG;
    int st = check_device(); if (st != ADN_OK) return st;
    const float inv = (float)(1.0 / pow(10.0, (double)snr_db / 20.0));
    mix_noise_snr_kernel<<<(unsigned)n_clips, 256, 0, (cudaStream_t)stream>>>(clean, noise, length, inv, out);
    ADN_LAUNCH_CHECK();
    return ADN_OK;
}

extern "C" int adn_mix_noise_cancel_f32(const float* clean, const unsigned char* block_flags, int64_t n_clips, int64_t length, int block,
                                        int half, float factor, float* out, void* stream) {
    if (n_clips < 0 || length <= 0 || block <= 0 || half < 0) return ADN_ERR_ARG;
    if (n_clips == 0) return ADN_OK;
    if (!clean || !block_flags || !out || n_clips > 0x7fffffffLL) return ADN_ERR_ARG;
    int st = check_device(); if (st != ADN_OK) return st;
    const int bpc = (int)((length + block - 1) / block);
    mix_noise_cancel_kernel<<<(unsigned)n_clips, 256, 0, (cudaStream_t)stream>>>(clean, block_flags, length, block, half, bpc, factor, out);
    ADN_LAUNCH_CHECK();
    return ADN_OK;
}
