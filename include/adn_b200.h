/*
 * adn_b200.h -- C ABI of libadn_b200.so: the B200 (sm_100a) kernels behind the AudioDenoiser hot path.
 *
 * The reference (jimonld2000/AudioDenoiser) has no FFI of its own: its hot path is Python calling
 * librosa / torch.  Each entry point below therefore names the reference *call site* it replaces
 * (paths relative to the reference root).  The Python host layer (audiodenoiser_b200/) binds these
 * with ctypes and re-exposes the reference's function names; INTEGRATION.md shows the stub a
 * reference maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name ends in _host;
 *   - every function returns an adn_status (0 = ok) and never throws; adn_error_string() explains it,
 *     adn_last_cuda_error() returns the cudaError_t behind ADN_ERR_CUDA for the calling thread;
 *   - launches are asynchronous on `stream` (a cudaStream_t passed as void*); _host variants copy in,
 *     run, copy out and synchronise the stream before returning;
 *   - re-entrant: no mutable global state (twiddle/window tables are compile-time constants);
 *   - there is NO CPU fallback: on a device that is not sm_100 the calls return ADN_ERR_DEVICE.
 *
 * Layouts
 *   wave      (n_clips, length) float32, row stride `clip_stride` elements
 *   mag       (n_clips, 257, T) float32, T contiguous            -- the reference .npy layout
 *   phasor    (n_clips, 257, T) complex64 (re,im interleaved)    -- test.py:36 `angles`
 *   audio     (n_clips, 128*(T-1)) float32
 *   act       NHWC bfloat16 (n, h, w, c), c contiguous           -- internal UNet activations
 */
#ifndef ADN_B200_H
#define ADN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADN_N_FFT 512
#define ADN_HOP 128
#define ADN_N_BINS 257

typedef enum {
    ADN_OK = 0,
    ADN_ERR_ARG = 1,      /* bad size / null pointer / unsupported shape */
    ADN_ERR_CUDA = 2,     /* a CUDA runtime call failed: see adn_last_cuda_error() */
    ADN_ERR_DEVICE = 3,   /* current device is not compute capability 10.x */
    ADN_ERR_DRIVER = 4,   /* cuTensorMapEncodeTiled could not be resolved / failed */
    ADN_ERR_SHORT = 5     /* input shorter than n_fft with center=0 (librosa ParameterError) */
} adn_status;

int adn_version(void);
const char* adn_error_string(int status);
int adn_last_cuda_error(void);
/* ADN_OK iff the current CUDA device can run this library (sm_100). */
int adn_device_check(void);

/* ------------------------------------------------------------------ spectral front / back end */

/* Frame count of librosa.stft(n_fft=512, hop_length=128, center=...): 1 + (L [+512] - 512)/128, or -1
 * when the input is too short.  create_train_dataset.py:167-172 (center=0), create_test_dataset.py:39. */
int64_t adn_stft_num_frames(int64_t length, int center);

/* |STFT|: framing + periodic Hann + 512-point real FFT + magnitude, one fused kernel.
 * Replaces librosa.stft + librosa.magphase in audio_to_magnitude_spectrogram
 * (create_train_dataset.py:162-174, center=0) and audio_to_spectrogram (create_test_dataset.py:35-41, center=1,
 * zero padding).  mag must hold n_clips*257*T floats. */
int adn_stft_mag_f32(const float* wave, int64_t n_clips, int64_t length, int64_t clip_stride, int center,
                     float* mag, void* stream);

/* Complex STFT (same framing), out (n_clips,257,T) complex64.  Replaces librosa.stft at test.py:41. */
int adn_stft_complex_f32(const float* wave, int64_t n_clips, int64_t length, int64_t clip_stride, int center,
                         float* spec_c64, void* stream);

/* Inverse STFT with overlap-add: (mag * phasor) -> irFFT -> * Hann -> gather-form overlap-add -> / sum(w^2) -> trim.
 * Replaces `magnitude_spectrogram * angles` + librosa.istft(hop_length=128) in griffin_lim_reconstruction
 * (test.py:36-37,40,48).  Atomic-free and deterministic: every output sample adds its <=4 frame contributions in
 * ascending frame order.  phasor may be NULL: a unit phasor with uniform random phase is then generated on the
 * device from `seed` (the reference draws it from the unseeded numpy RNG, test.py:36).
 * mag may be NULL when `spec_is_complex` != 0, in which case phasor holds the complex spectrogram itself. */
int adn_istft_ola_f32(const float* mag, const float* phasor_c64, int spec_is_complex, uint64_t seed,
                      int64_t n_clips, int64_t n_frames, float* audio, void* stream);

/* The unit phasor adn_istft_ola_f32 generates for `seed` when phasor == NULL, written out as (n_clips,257,T) complex64:
 * the device-side stand-in for `angles = np.exp(2j * np.pi * np.random.rand(*mag.shape))` (test.py:36).  Lets a caller
 * (and the parity tests) reproduce a seeded reconstruction with an explicit phasor, bit for bit. */
int adn_random_phasor_c64(uint64_t seed, int64_t n_clips, int64_t n_frames, float* phasor_c64, void* stream);

/* Host-buffer variants (pageable or pinned host memory; allocate device scratch internally, synchronise). */
int adn_stft_mag_host_f32(const float* wave_host, int64_t n_clips, int64_t length, int center, float* mag_host);
int adn_istft_ola_host_f32(const float* mag_host, const float* phasor_c64_host, uint64_t seed,
                           int64_t n_clips, int64_t n_frames, float* audio_host);

/* ------------------------------------------------------------------ UNet (code/model.py) */

/* Weight packing, done once per checkpoint load (model.py:53-68 state_dict layout, fp32 on device):
 *   conv3x3 (Co,Ci,3,3) f32  ->  bf16 [Co][tap=ky*3+kx][Ci]      (K-major GEMM B operand, K = 9*Ci)
 *   convT   (Ci,Co,2,2) f32  ->  bf16 [q=dy*2+dx][Co][Ci]         (K-major, N = 4*Co, K = Ci)
 *   BN fold: scale = gamma / sqrt(var + eps); shift = (conv_bias - mean) * scale + beta   (fp32) */
int adn_pack_conv3x3_weight_bf16(const float* w, int c_out, int c_in, void* packed_bf16, void* stream);
int adn_pack_convt2x2_weight_bf16(const float* w, int c_in, int c_out, void* packed_bf16, void* stream);
int adn_fold_bn_f32(const float* conv_bias, const float* gamma, const float* beta, const float* mean,
                    const float* var, float eps, int channels, float* scale, float* shift, void* stream);

/* First layer, downconv1.conv.double_conv.0-2 (model.py:11-13 with Ci=1): direct 3x3 conv from the fp32
 * magnitude (n,1,h,w) + folded BN + ReLU -> NHWC bf16 (n,h,w,64).  w: (64,1,3,3) fp32. */
int adn_conv3x3_c1_bn_relu_bf16(const float* x, int n, int h, int w, const float* weight, const float* scale,
                                const float* shift, void* out_bf16, void* stream);

/* Conv3x3(pad 1) + folded BN + ReLU as a tcgen05/TMEM implicit GEMM (model.py:11-16).  The input is the channel
 * concatenation [src0, src1] (model.py:49 torch.cat([x2, x1])) without materialising it; src1 may be NULL (c1 = 0).
 * src1 has spatial size (h1,w1) <= (h,w): the missing bottom row / right column read as zero, which is the
 * F.pad of model.py:44-47 (diff//2 = 0 before, diff after).  c0, c1 multiples of 64; c_out multiple of 64.
 * If pool_out != NULL the 2x2/2 max-pool (model.py:31, floor) of the result is also written, (n,h/2,w/2,c_out). */
int adn_conv3x3_bn_relu_bf16(const void* src0, int c0, const void* src1, int c1, int h1, int w1,
                             int n, int h, int w, const void* w_packed, int c_out,
                             const float* scale, const float* shift, void* out_bf16, void* pool_out, void* stream);

/* Same conv with the 1x1 head fused into the epilogue (model.py:68,93: out = Conv2d(64,1,1), no activation):
 * requires c_out == 64; writes (n,1,h,w) fp32 and does not store the 64-channel activation. */
int adn_conv3x3_bn_relu_head_f32(const void* src0, int c0, const void* src1, int c1, int h1, int w1,
                                 int n, int h, int w, const void* w_packed, int c_out,
                                 const float* scale, const float* shift,
                                 const float* head_w, const float* head_b, float* out_f32, void* stream);

/* ConvTranspose2d(k=2,s=2) + bias as a tcgen05 GEMM with a pixel-shuffle store (model.py:38,43):
 * src (n,h,w,c_in) -> out (n,2h,2w,c_out).  bias: (c_out) fp32. */
int adn_convt2x2_bf16(const void* src, int c_in, int n, int h, int w, const void* w_packed, int c_out,
                      const float* bias, void* out_bf16, void* stream);

/* MaxPool2d(2) on NHWC bf16 (model.py:26,31), floor semantics: (n,h,w,c) -> (n,h/2,w/2,c). */
int adn_maxpool2x2_bf16(const void* src, int n, int h, int w, int c, void* out, void* stream);

/* Layout helpers for tests / host glue: NHWC bf16 <-> NCHW fp32. */
int adn_nhwc_bf16_to_nchw_f32(const void* src, int n, int h, int w, int c, float* dst, void* stream);
int adn_nchw_f32_to_nhwc_bf16(const float* src, int n, int c, int h, int w, void* dst, void* stream);

/* ------------------------------------------------------------------ loader transform / loss statistics */

/* SpectrogramDataset.__getitem__ (data_loader.py:37-72) on the device: float32 -> float16 round trip ->
 * zero-pad / crop (f_in,t_in) -> (f_out,t_out) -> float32.  src (n,f_in,t_in), dst (n,f_out,t_out). */
int adn_spec_f16_crop_f32(const float* src, int64_t n, int f_in, int t_in, int f_out, int t_out, float* dst, void* stream);

/* Partial sums for the statistics the clip-sharded path all-reduces:
 * sums[0] += sum|pred-target|, sums[1] += sum target^2, sums[2] += sum (target-pred)^2, over `count` elements.
 * sums is a device array of 3 doubles that the caller zeroes. */
int adn_spec_error_sums_f64(const float* pred, const float* target, int64_t count, double* sums, void* stream);

/* CombinedPerceptualLoss.forward (loss.py:83-95) = 0.4 * MultiScaleSTFTLoss (loss.py:12-35) + 0.4 * MelSpectrogramLoss
 * (loss.py:44-69) + 0.2 * L1Loss (loss.py:76,86) on (batch,1,freq,frames) float32 magnitude tensors (test.py:118-122,
 * train.py:68,85).  out4 (device, 4 floats) = {total, stft, mel, l1}.  mel_fb_32x64: the torchaudio HTK filterbank
 * melscale_fbanks(32, 0, 4000, 64, 8000, norm=None) as (32, 64) float32 on the device.  workspace: adn_loss_workspace_bytes()
 * bytes of device scratch, 256-byte aligned.  Requires 31 < frames <= 8192 (reflect padding of the 63-point mel STFT).
 * Deterministic: ordered two-stage reductions, no atomics. */
int64_t adn_loss_workspace_bytes(int64_t batch, int freq, int frames);
int adn_combined_loss_f32(const float* pred, const float* target, int64_t batch, int freq, int frames,
                          const float* mel_fb_32x64, void* workspace, float* out4, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ADN_B200_H */
