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
/* First 16 hex digits of the sha256 over the sources (csrc/ and include/) this binary was built from; the Python binding refuses a
 * library whose hash differs from the sources beside it (audiodenoiser_b200/build.py:source_hash). */
const char* adn_source_hash(void);
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

/* |STFT| with the SpectrogramDataset transform (data_loader.py:41-72: float32 -> float16 -> float32, zero-pad / crop to
 * (f_out, t_out)) as a second output of the same pass: crop (n_clips, f_out, t_out) float32.  mag may be NULL,
 * in which case only the frames that survive the crop are transformed (the training-tensor path: waveform -> (256,64)). */
int adn_stft_mag_crop_f16_f32(const float* wave, int64_t n_clips, int64_t length, int64_t clip_stride, int center, float* mag,
                              float* crop, int f_out, int t_out, void* stream);

/* Complex STFT (same framing), out (n_clips,257,T) complex64.  Replaces librosa.stft at test.py:41. */
int adn_stft_complex_f32(const float* wave, int64_t n_clips, int64_t length, int64_t clip_stride, int center,
                         float* spec_c64, void* stream);

/* Inverse STFT with overlap-add: (mag * phasor) -> irFFT -> * Hann -> gather-form overlap-add -> / sum(w^2) -> trim.
 * Replaces `magnitude_spectrogram * angles` + librosa.istft(hop_length=128) in griffin_lim_reconstruction
 * (test.py:36-37,40,48).  Atomic-free and deterministic: every output sample adds its <=4 frame contributions in
 * ascending frame order.  phasor may be NULL: a unit phasor with uniform random phase is then generated on the
 * device from `seed` (the reference draws it from the unseeded numpy RNG, test.py:36).
 * spec_is_complex == 1: mag may be NULL and phasor holds the complex spectrogram itself (librosa.istft at test.py:40).
 * spec_is_complex == 2 (opt-in, not what the reference computes): phasor holds an arbitrary complex
 * spectrogram C and the kernel inverts mag * C / |C| -- e.g. the denoised magnitude with the NOISY input's phase. */
int adn_istft_ola_f32(const float* mag, const float* phasor_c64, int spec_is_complex, uint64_t seed,
                      int64_t n_clips, int64_t n_frames, float* audio, void* stream);

/* The random-phase form of adn_istft_ola_f32 with a DEVICE-side call counter: the phase is drawn from seed + *seed_counter_dev.
 * test.py:36 draws a fresh phase per call; a CUDA-graph replay bakes kernel arguments in, so the per-call part of the seed lives in
 * device memory and is advanced inside the same graph by adn_u64_add (one thread).  adn_random_phasor_c64(seed + counter) exports
 * the phasor of any call. */
int adn_istft_ola_counter_f32(const float* mag, uint64_t seed, const uint64_t* seed_counter_dev, int64_t n_clips, int64_t n_frames,
                              float* audio, void* stream);
int adn_u64_add(uint64_t* counter_dev, uint64_t inc, void* stream);

/* The unit phasor adn_istft_ola_f32 generates for `seed` when phasor == NULL, written out as (n_clips,257,T) complex64:
 * the device-side stand-in for `angles = np.exp(2j * np.pi * np.random.rand(*mag.shape))` (test.py:36).  Lets a caller
 * (and the parity tests) reproduce a seeded reconstruction with an explicit phasor, bit for bit. */
int adn_random_phasor_c64(uint64_t seed, int64_t n_clips, int64_t n_frames, float* phasor_c64, void* stream);

/* Host-buffer variants (pageable or pinned host memory; allocate device scratch internally, synchronise). */
int adn_stft_mag_host_f32(const float* wave_host, int64_t n_clips, int64_t length, int center, float* mag_host);
int adn_istft_ola_host_f32(const float* mag_host, const float* phasor_c64_host, uint64_t seed,
                           int64_t n_clips, int64_t n_frames, float* audio_host);

/* ------------------------------------------------------------------ noise synthesis
 * add_noise of create_train_dataset.py:105-159 / create_test_dataset.py:43-133, batched: clean, noise, out are (n_clips, length)
 * float32, contiguous.  The random draws stay with the host (the reference's RNG calls, in its order) and are passed in.
 *   snr:    out = clip(clean + noise * clean_rms / 10^(snr_db/20) / noise_rms, -1, 1), rms = sqrt(mean(x^2) + 1e-12); noise is
 *           dropped when noise_rms <= 1e-9 (:148-157).  "white": noise = N(0,1) samples; "urban": the tiled / cropped snippet.
 *   cancel: block_flags (n_clips, ceil(length/block)) uint8; in every flagged block the first `half` samples become
 *           clean + factor * clean (factor = -0.8, block = 16000, half = 8000 in the reference, :123-135), then clip. */
int adn_mix_noise_snr_f32(const float* clean, const float* noise, int64_t n_clips, int64_t length, float snr_db, float* out,
                          void* stream);
int adn_mix_noise_cancel_f32(const float* clean, const unsigned char* block_flags, int64_t n_clips, int64_t length, int block,
                             int half, float factor, float* out, void* stream);

/* ------------------------------------------------------------------ resample front-end
 * The numeric tail of librosa.load(path, sr=8000) (create_train_dataset.py:204,225; create_test_dataset.py:139; test.py:80):
 * librosa.to_mono + librosa.resample.  x: (n_clips, channels, len_in) float32 planar, mixed down to mono inside the kernel;
 * y: (n_clips, len_out), len_out = ceil(len_in * up / down).  Polyphase FIR of scipy.signal.resample_poly:
 *   y[m] = sum_{t < taps} xm[j - t] * table[r * taps_pitch + t],  c = (m + pre_remove) * down, j = c / up, r = c % up.
 * table: (up, taps_pitch) float32 on the device, 16-byte aligned, rows zero-padded to taps_pitch = a multiple of 4
 * (audiodenoiser_b200/resample.py builds it: Kaiser beta 5 windowed sinc, half
 * length 10 * max(up, down), unit DC gain, scaled by up).  The reference's own resampler (soxr_hq inside librosa) is an absent
 * third-party library with an unspecified filter: parity is against the published scipy algorithm, approximate against soxr. */
int adn_resample_poly_f32(const float* x, int64_t n_clips, int channels, int64_t len_in, int up, int down, const float* table,
                          int taps, int taps_pitch, int64_t pre_remove, int64_t len_out, float* y, void* stream);

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

/* UpSampleLayer's ConvTranspose2d + F.pad + cat + first Conv3x3 + BN + ReLU (model.py:41-49, :11-13) as ONE implicit GEMM with the
 * ConvTranspose merged into the conv weights (csrc/conv_upm.cu): skip (n,h,w,c0) and low (n,hl,wl,cl) NHWC bf16 -> out (n,h,w,c_out),
 * h - 2*hl and w - 2*wl in {0,1} (F.pad puts the missing row / column at the end), c0, cl multiples of 64, c_out a multiple of 128.
 * w_merged / shift_m / wb come from adn_pack_upmerged_weight_bf16.  Eval mode only (the training step keeps the two layers apart). */
int adn_conv3x3_upmerged_bn_relu_bf16(const void* skip, int c0, const void* low, int cl, int hl, int wl, int n, int h, int w,
                                      const void* w_merged, int c_out, const float* scale, const float* shift_m, const float* wb,
                                      void* out_bf16, void* stream);

/* Checkpoint-load-time merge for the entry above.  w3: the level's first conv weight (c_out, c0 + cup, 3, 3) f32 (input channels
 * [skip, up], model.py:49); wt / bt: ConvTranspose2d weight (cl, cup, 2, 2) and bias (cup) f32; scale / shift: the folded BatchNorm of
 * the conv (adn_fold_bn_f32).  Writes w_merged bf16 [c_out][9*c0 + 16*cl] (skip taps, then per parity class and 2x2 tap the product
 * of the conv and ConvTranspose weights, summed in fp32), shift_m (c_out) = shift + the interior ConvTranspose-bias term, and
 * wb (9, c_out) = the per-tap bias terms the kernel takes back out on border pixels. */
int adn_pack_upmerged_weight_bf16(const float* w3, const float* wt, const float* bt, const float* scale, const float* shift,
                                  int c_out, int c0, int cup, int cl, void* w_merged, float* shift_m, float* wb, void* stream);

/* The same merged layer for c_out == 128 (decoder level of 128 channels): both column-parity classes of a tile share one
 * 256-column accumulator, so the skip planes are loaded once for two classes (csrc/conv_upm.cu, conv3x3_upm2_kernel).
 * bsh / b1: adn_pack_upmerged_pair_weight_bf16 of the w_merged tensor above; adn_upmerged_pair_weight_elems(.., which) gives
 * their element counts (which = 0: bsh, 1: b1).  shift_m / wb as above.
 * cl == 0 (low = wb = NULL, w_merged = the [c_out][9][c0] pack of adn_pack_conv3x3_weight_bf16): a plain Conv3x3 + BN + ReLU over
 * `skip` in the same parity-class formulation, used for the 128-output-channel layers without a fused pool. */
int64_t adn_upmerged_pair_weight_elems(int c_out, int c0, int cl, int which);
int adn_pack_upmerged_pair_weight_bf16(const void* w_merged, int c_out, int c0, int cl, void* bsh, void* b1, void* stream);
int adn_conv3x3_upmerged_pair_bn_relu_bf16(const void* skip, int c0, const void* low, int cl, int hl, int wl, int n, int h, int w,
                                           const void* bsh, const void* b1, int c_out, const float* scale, const float* shift_m,
                                           const float* wb, void* out_bf16, void* stream);

/* Conv3x3 + BN + ReLU + fused MaxPool2d(2) (DownSampleLayer, model.py:29-32) for c_out == 128 in the same parity-class formulation:
 * the 2x2 pool window of a half-resolution pixel is exactly its four parity classes, so the pooled value is the running maximum over
 * the two column classes of a tile and the two row-parity tiles of a region (one CTA pair runs both back to back).  src (n,h,w,c_in),
 * out (n,h,w,128), pool_out (n,h/2,w/2,128); bsh / b1 from adn_pack_upmerged_pair_weight_bf16(w_packed [128][9][c_in], cl = 0). */
int adn_conv3x3_pair_bn_relu_pool_bf16(const void* src, int c_in, int n, int h, int w, const void* bsh, const void* b1, int c_out,
                                       const float* scale, const float* shift, void* out_bf16, void* pool_out, void* stream);

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
/* The bookkeeping slots of the same 8-double statistics vector: sums8[3] = numel, and when loss4 (the device 4-vector of
 * adn_combined_loss_f32: total, stft, mel, l1 -- test.py:118-122) is given, sums8[4] = n_clips, sums8[5..7] = n_clips x the three terms,
 * so that the all-reduced vector yields exact full-batch means for unequal shards. */
int adn_stats_pack_f64(double* sums8, int64_t numel, int64_t n_clips, const float* loss4, void* stream);
/* Zero `bytes` bytes of device memory on `stream` (cudaMemsetAsync): how the caller zeroes the statistics vector above. */
int adn_zero_bytes(void* dev, int64_t bytes, void* stream);
/* Device-to-device cudaMemcpyAsync and v[0..n) += inc: the bookkeeping of a training step (loss vector, head-bias gradient, the 18
 * num_batches_tracked counters of nn.BatchNorm2d) without framework kernels. */
int adn_copy_bytes(void* dst_dev, const void* src_dev, int64_t bytes, void* stream);
int adn_i64_add_n(int64_t* v_dev, int n, int64_t inc, void* stream);

/* CombinedPerceptualLoss.forward (loss.py:83-95) = 0.4 * MultiScaleSTFTLoss (loss.py:12-35) + 0.4 * MelSpectrogramLoss
 * (loss.py:44-69) + 0.2 * L1Loss (loss.py:76,86) on (batch,1,freq,frames) float32 magnitude tensors (test.py:118-122,
 * train.py:68,85).  out4 (device, 4 floats) = {total, stft, mel, l1}.  mel_fb_32x64: the torchaudio HTK filterbank
 * melscale_fbanks(32, 0, 4000, 64, 8000, norm=None) as (32, 64) float32 on the device.  workspace: adn_loss_workspace_bytes()
 * bytes of device scratch, 256-byte aligned.  Requires 31 < frames <= 8192 (reflect padding of the 63-point mel STFT).
 * Deterministic: ordered two-stage reductions, no atomics. */
int64_t adn_loss_workspace_bytes(int64_t batch, int freq, int frames);
int adn_combined_loss_f32(const float* pred, const float* target, int64_t batch, int freq, int frames,
                          const float* mel_fb_32x64, void* workspace, float* out4, void* stream);


/* ------------------------------------------------------------------ training step (code/train.py:65-72)
 *
 * train.py:65-72 is zero_grad -> model(noisy) in train() mode -> criterion -> loss.backward() -> clip_grad_norm_(1.0) ->
 * AdamW.step().  The entry points below are the device side of that body; the host side (audiodenoiser_b200/training.py)
 * strings them together layer by layer.  Activation gradients are NHWC bf16, parameter gradients fp32 in the reference
 * state_dict layout.  `workspace` is adn_train_workspace_bytes() bytes of device scratch (256-byte aligned). */
int64_t adn_train_workspace_bytes(void);

/* Conv3x3 without the activation: out = conv * scale + shift [-> ReLU if relu != 0].  Train-mode pre-BatchNorm output
 * (scale = 1, shift = conv bias; model.py:11,14 before :12,15) and the data gradient of the backward pass (weights packed by
 * adn_pack_conv3x3_dgrad_weight_bf16, scale = 1, shift = 0).  Same tcgen05 kernel and argument meaning as
 * adn_conv3x3_bn_relu_bf16; first-layer (Ci = 1) variant below. */
int adn_conv3x3_affine_bf16(const void* src0, int c0, const void* src1, int c1, int h1, int w1, int n, int h, int w,
                            const void* w_packed, int c_out, const float* scale, const float* shift, int relu, void* out_bf16,
                            void* stream);
int adn_conv3x3_c1_affine_bf16(const float* x, int n, int h, int w, const float* weight, const float* scale, const float* shift,
                               int relu, void* out_bf16, void* stream);

/* Data-gradient weight packing: conv (Co,Ci,3,3) f32 -> bf16 [Ci][8 - tap][Co]; convT (Ci,Co,2,2) f32 -> bf16 [Ci][q*Co + co]. */
int adn_pack_conv3x3_dgrad_weight_bf16(const float* w, int c_out, int c_in, void* packed_bf16, void* stream);
int adn_pack_convt2x2_dgrad_weight_bf16(const float* w, int c_in, int c_out, void* packed_bf16, void* stream);

/* All weight packs of a training step in ONE launch.  table_dev: device array of n_entries records
 *   struct { const float* w; void* fwd_bf16; void* dgrad_bf16; int32_t c_out, c_in, kind, pad; }   (32 bytes)
 * kind 0 = Conv2d 3x3 (fwd [Co][tap][Ci], dgrad [Ci][8-tap][Co]); kind 1 = ConvTranspose2d 2x2 (fwd [q][Co][Ci], dgrad [Ci][q*Co+co]). */
int adn_pack_weights_table_bf16(const void* table_dev, int n_entries, void* stream);
/* The same with a selection: which = 1 forward packs only, 2 data-gradient packs only, 3 both (the training step packs the forward
 * operands after AdamW and the data-gradient operands on a side stream under the next forward). */
int adn_pack_weights_table_sel_bf16(const void* table_dev, int n_entries, int which, void* stream);
/* The same over a ONE-dimensional grid: record field `pad` = the number of blocks entry i gets (kind 0: one block per 32 x 32 x 9 tile,
 * (c_out/32)*(c_in/32); kind 1: any count >= 1), total_blocks = their sum.  The 2-D forms above launch 1 024 blocks per entry. */
int adn_pack_weights_table_flat_bf16(const void* table_dev, int n_entries, int total_blocks, int which, void* stream);

/* nn.BatchNorm2d in train() mode (model.py:12,15; eps 1e-5, momentum 0.1): batch statistics of z (pixels, c) NHWC bf16 ->
 * scale = gamma * invstd, shift = beta - mean * scale, mean, invstd; running_mean / running_var are updated in place
 * (unbiased variance) unless NULL.  Then y = max(z * scale + shift, 0) (BatchNorm + ReLU, model.py:12-13). */
int adn_bn_train_stats_f32(const void* z_bf16, int64_t pixels, int c, const float* gamma, const float* beta, float eps,
                           float momentum, float* running_mean, float* running_var, float* scale, float* shift, float* mean,
                           float* invstd, void* workspace, void* stream);
int adn_bn_relu_apply_bf16(const void* z_bf16, const float* scale, const float* shift, int64_t pixels, int c, void* y_bf16,
                           void* stream);

/* Backward of ReLU o BatchNorm2d(train): dy (pixel stride dy_ld >= c channels: may be a channel slice) and the saved z ->
 * dz (dense), d_gamma, d_beta (fp32, overwritten). */
int adn_bn_relu_backward_bf16(const void* dy_bf16, int dy_ld, const void* z_bf16, int64_t pixels, int c, const float* scale,
                              const float* shift, const float* mean, const float* invstd, float* d_gamma, float* d_beta,
                              void* dz_bf16, void* workspace, void* stream);

/* out[ch] = sum over pixels of x[p][ch] (ConvTranspose2d bias gradient; x may be a channel slice with pixel stride x_ld). */
int adn_channel_sum_f32(const void* x_bf16, int x_ld, int64_t pixels, int c, float* out, void* workspace, void* stream);

/* MaxPool2d(2) backward (gradient to the first maximum of each window, like torch) + the skip-connection gradient
 * (model.py:31,49): out = d_dec + route(d_pool).  y: the pooled layer's input (n,h,w,c); d_dec may be NULL or a channel slice. */
int adn_maxpool2x2_backward_add_bf16(const void* y_bf16, const void* d_pool_bf16, const void* d_dec_bf16, int dec_ld, int n, int h,
                                     int w, int c, void* out_bf16, void* stream);

/* 1x1 head Conv2d(64,1,1) (model.py:68,93): forward (pixels,64) bf16 -> (pixels) fp32; backward -> dy bf16, d_w (64), d_b (1). */
int adn_head1x1_forward_f32(const void* y_bf16, const float* w, const float* b, int64_t pixels, float* out, void* stream);
int adn_head1x1_backward(const void* y_bf16, const float* d_out, const float* w, int64_t pixels, void* dy_bf16, float* d_w,
                         float* d_b, void* workspace, void* stream);

/* Weight gradients, WRITTEN to fp32 buffers in the reference layout (every element of the addressed slice is overwritten, which
 * is what zero_grad() + backward() of train.py:66,69 leaves behind).
 *   conv3x3:  d_weight (Co, ci_total, 3, 3), columns [ci_off, ci_off + c_in) from input x (n,h1,w1,c_in) [(h1,w1) <= (h,w): the
 *             zero-padded up-sampled half of a concatenated input, model.py:44-49]; tcgen05 GEMM contracting over pixels,
 *             split-K partials folded in a fixed order (deterministic); wgrad_workspace: adn_wgrad_workspace_bytes() bytes.
 *   conv3x3_c1: first layer, d_weight (64,1,3,3), overwritten (deterministic reduction).
 *   convt2x2: d_weight (Ci, Co, 2, 2); d_up = channels [up_off, up_off + c_out) of an (n,2h,2w,up_ld) gradient tensor. */
int64_t adn_wgrad_workspace_bytes(void);   /* split-K partial tiles of the two tensor-core weight-gradient entry points */
int adn_conv3x3_wgrad_f32(const void* dz_bf16, int c_out, const void* x_bf16, int c_in, int h1, int w1, int n, int h, int w,
                          float* d_weight, int ci_off, int ci_total, void* wgrad_workspace, void* stream);
int adn_conv3x3_c1_wgrad_f32(const void* dz_bf16, const float* x, int n, int h, int w, float* d_weight, void* workspace, void* stream);
int adn_convt2x2_wgrad_f32(const void* x_bf16, int c_in, const void* d_up_bf16, int up_ld, int up_off, int c_out, int n, int h, int w,
                           float* d_weight, void* wgrad_workspace, void* stream);

/* Data gradient of ConvTranspose2d(k=2,s=2): d_in (n,h,w,c_in) from channels [out_off, out_off + c_out) of d_out (n,2h,2w,out_ld). */
int adn_convt2x2_dgrad_bf16(const void* d_out_bf16, int out_ld, int out_off, int c_out, int n, int h, int w, const void* w_packed,
                            int c_in, void* d_in_bf16, void* stream);

/* Gradient of c_stft * stft + c_mel * mel + c_l1 * l1 (loss.py:83-95) with respect to pred; train.py:69 is (0.4, 0.4, 0.2). */
int64_t adn_loss_backward_workspace_bytes(int64_t batch, int freq, int frames);
int adn_combined_loss_backward_f32(const float* pred, const float* target, int64_t batch, int freq, int frames,
                                   const float* mel_fb_32x64, float c_stft, float c_mel, float c_l1, void* workspace,
                                   float* d_pred, void* stream);

/* torch.nn.utils.clip_grad_norm_(params, max_norm) (train.py:70) over one flat fp32 gradient buffer:
 * norm_and_coef[0] = total L2 norm, [1] = min(1, max_norm / (norm + 1e-6)).  torch.optim.AdamW.step (train.py:71,124) on flat
 * buffers, gradients scaled by norm_and_coef[1] (pass NULL for no clipping); `step` is the 1-based step count. */
int adn_grad_norm_f32(const float* grads, int64_t count, float max_norm, float* norm_and_coef, void* workspace, void* stream);
int adn_adamw_step_f32(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t count,
                       const float* norm_and_coef, float lr, float beta1, float beta2, float eps, float weight_decay, int64_t step,
                       void* stream);
/* Same update with the step count kept on the device (step_counter[0], a float the call advances by one before use), so that a
 * CUDA-graph capture of the whole training step can be replayed. */
int adn_adamw_step_dev_f32(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t count,
                           const float* norm_and_coef, float* step_counter, float lr, float beta1, float beta2, float eps,
                           float weight_decay, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ADN_B200_H */
