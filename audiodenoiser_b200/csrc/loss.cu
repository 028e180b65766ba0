// loss.cu -- CombinedPerceptualLoss forward of the reference (code/loss.py:6-95) as three small sm_100a kernels.
//
//   loss.py:14-20 / 46-52   mean over the frequency axis -> per-sample time envelope (B, T)         \  loss_envelope_kernel
//   loss.py:76,86           nn.L1Loss(pred, target): mean |pred - target| over every element        /  (one pass over the inputs)
//   loss.py:22-35           three rectangular-window STFTs of the envelope (n_fft 63/32/16, hop 16/8/4, center=True with
//                           zero padding, onesided), |.|, L1, averaged over the scales               \  loss_spectral_kernel
//   loss.py:40-42,60-69     torchaudio MelSpectrogram(sr 8000, n_fft 63, hop 16, n_mels 64): periodic Hann(63), reflect
//                           padding, power spectrum, HTK filterbank (32 x 64, no norm), L1           /  (one CTA per sample)
//   loss.py:88-95           total = 0.4 stft + 0.4 mel + 0.2 l1                                        loss_finalize_kernel
//
// The envelopes are tiny (B x T), so the transforms are direct DFTs from shared memory with an exact-angle twiddle table
// instead of an FFT: a few hundred thousand MACs per sample.  All reductions are two-stage and ordered (no atomics), so the
// four scalars are bit-reproducible run to run.  HBM-bound on the single read of pred and target (8 bytes per element).
#include "adn_common.cuh"

namespace adn {

constexpr int LOSS_MAX_T = 8192;
constexpr int ENV_TX = 32, ENV_FY = 8;
constexpr int MEL_NFFT = 63, MEL_BINS = 32, MEL_N = 64, MEL_HOP = 16;

// ------------------------------------------------------------------------------------------------ envelope + L1
// grid (ceil(T/32), B); block (32, 8): lane x owns time column t, the 8 y-slices stride the frequency axis.
__global__ void __launch_bounds__(ENV_TX * ENV_FY)
loss_envelope_kernel(const float* __restrict__ pred, const float* __restrict__ target, int F, int T,
                     float* __restrict__ env, double* __restrict__ l1_partial) {
    __shared__ float sp[ENV_FY][ENV_TX], stg[ENV_FY][ENV_TX];
    __shared__ double sl[ENV_FY];
    const int b = blockIdx.y;
    const int t = blockIdx.x * ENV_TX + threadIdx.x;
    const size_t base = (size_t)b * F * T;
    float ap = 0.f, at = 0.f, al = 0.f;
    if (t < T) {
        for (int f = threadIdx.y; f < F; f += ENV_FY) {
            const float p = __ldcs(pred + base + (size_t)f * T + t), g = __ldcs(target + base + (size_t)f * T + t);
            ap += p; at += g; al += fabsf(p - g);
        }
    }
    sp[threadIdx.y][threadIdx.x] = ap;
    stg[threadIdx.y][threadIdx.x] = at;
    double l = (double)al;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
    if (threadIdx.x == 0) sl[threadIdx.y] = l;
    __syncthreads();
    if (threadIdx.y == 0) {
        if (t < T) {
            float a = 0.f, c = 0.f;
#pragma unroll
            for (int y = 0; y < ENV_FY; ++y) { a += sp[y][threadIdx.x]; c += stg[y][threadIdx.x]; }
            env[((size_t)b * 2 + 0) * T + t] = a / (float)F;
            env[((size_t)b * 2 + 1) * T + t] = c / (float)F;
        }
        if (threadIdx.x == 0) {
            double s = 0.0;
#pragma unroll
            for (int y = 0; y < ENV_FY; ++y) s += sl[y];
            l1_partial[(size_t)b * gridDim.x + blockIdx.x] = s;
        }
    }
}

// ------------------------------------------------------------------------------------------------ spectral terms
__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    return s;       // valid in thread 0
}

// one CTA per sample; sums[b] = { sum|dmag| scale 63, scale 32, scale 16, sum|dmel| }
__global__ void __launch_bounds__(256)
loss_spectral_kernel(const float* __restrict__ env, int T, const float* __restrict__ mel_fb, double* __restrict__ sums) {
    extern __shared__ float smem[];
    float* xp = smem;                       // [T]
    float* xt = xp + T;                     // [T]
    float* tw_c = xt + T;                   // [64] cos(2 pi m / n)
    float* tw_s = tw_c + 64;                // [64] sin(2 pi m / n)
    float* win = tw_s + 64;                 // [64] periodic Hann(63)
    float* fb = win + 64;                   // [32][64]
    float* pw = fb + MEL_BINS * MEL_N;      // [2][frames_chunk = 8][32] power spectra
    __shared__ double red[8];
    const int b = blockIdx.x, tid = threadIdx.x;
    for (int i = tid; i < T; i += blockDim.x) { xp[i] = env[((size_t)b * 2) * T + i]; xt[i] = env[((size_t)b * 2 + 1) * T + i]; }
    for (int i = tid; i < MEL_BINS * MEL_N; i += blockDim.x) fb[i] = mel_fb[i];
    if (tid < 64) win[tid] = tid < MEL_NFFT ? 0.5f - 0.5f * cospif(2.0f * (float)tid / (float)MEL_NFFT) : 0.f;

    // ---- loss.py:22-35: rectangular-window magnitudes at three scales
    const int nffts[3] = {63, 32, 16}, hops[3] = {16, 8, 4};
    // blockIdx.y selects ONE of the four transforms (three rectangular scales, mel): 4 x batch CTAs instead of batch (64 CTAs on 148
    // SMs each running the four transforms back to back took 110 us per step at the benchmark shape)
    for (int s = 0; s < 3; ++s) {
        if ((int)blockIdx.y != s) continue;
        const int n = nffts[s], hop = hops[s], pad = n / 2, bins = n / 2 + 1;
        const int frames = 1 + (T + 2 * pad - n) / hop;
        __syncthreads();
        if (tid < n) sincospif(2.0f * (float)tid / (float)n, &tw_s[tid], &tw_c[tid]);
        __syncthreads();
        double acc = 0.0;
        for (int item = tid; item < frames * bins; item += blockDim.x) {
            const int fr = item / bins, k = item - fr * bins;
            const int start = fr * hop - pad;
            float pr = 0.f, pi = 0.f, tr = 0.f, ti = 0.f;
            int m = 0;                                             // (k * i) mod n, advanced incrementally
            for (int i = 0; i < n; ++i) {
                const int idx = start + i;
                if (idx >= 0 && idx < T) {
                    const float c = tw_c[m], sn = tw_s[m];
                    const float a = xp[idx], g = xt[idx];
                    pr = fmaf(a, c, pr); pi = fmaf(a, sn, pi);
                    tr = fmaf(g, c, tr); ti = fmaf(g, sn, ti);
                }
                m += k; if (m >= n) m -= n;
            }
            acc += (double)fabsf(sqrtf(pr * pr + pi * pi) - sqrtf(tr * tr + ti * ti));
        }
        const double tot = block_sum(acc, red);
        if (tid == 0) sums[(size_t)b * 4 + s] = tot;
    }

    // ---- loss.py:40-69: mel power spectrogram (Hann 63, reflect padding), 8 frames per pass
    if (blockIdx.y == 3) {
        const int pad = MEL_NFFT / 2;
        const int frames = 1 + (T + 2 * pad - MEL_NFFT) / MEL_HOP;
        __syncthreads();
        if (tid < MEL_NFFT) sincospif(2.0f * (float)tid / (float)MEL_NFFT, &tw_s[tid], &tw_c[tid]);
        __syncthreads();
        double acc = 0.0;
        for (int f0 = 0; f0 < frames; f0 += 8) {
            const int nf = min(8, frames - f0);
            // power spectra: item = (signal, frame, bin)
            for (int item = tid; item < 2 * nf * MEL_BINS; item += blockDim.x) {
                const int sig = item / (nf * MEL_BINS);
                const int rem = item - sig * nf * MEL_BINS;
                const int fr = rem / MEL_BINS, k = rem - fr * MEL_BINS;
                const float* x = sig ? xt : xp;
                const int start = (f0 + fr) * MEL_HOP - pad;
                float re = 0.f, im = 0.f;
                int m = 0;
                for (int i = 0; i < MEL_NFFT; ++i) {
                    int idx = start + i;
                    if (idx < 0) idx = -idx;                       // reflect (no edge repeat), torch pad_mode='reflect'
                    if (idx >= T) idx = 2 * (T - 1) - idx;
                    const float v = x[idx] * win[i];
                    re = fmaf(v, tw_c[m], re); im = fmaf(v, tw_s[m], im);
                    m += k; if (m >= MEL_NFFT) m -= MEL_NFFT;
                }
                pw[(sig * 8 + fr) * MEL_BINS + k] = re * re + im * im;
            }
            __syncthreads();
            for (int item = tid; item < nf * MEL_N; item += blockDim.x) {
                const int fr = item / MEL_N, mel = item - fr * MEL_N;
                float mp = 0.f, mt = 0.f;
#pragma unroll 8
                for (int k = 0; k < MEL_BINS; ++k) {
                    const float w = fb[k * MEL_N + mel];
                    mp = fmaf(pw[fr * MEL_BINS + k], w, mp);
                    mt = fmaf(pw[(8 + fr) * MEL_BINS + k], w, mt);
                }
                acc += (double)fabsf(mp - mt);
            }
            __syncthreads();
        }
        const double tot = block_sum(acc, red);
        if (tid == 0) sums[(size_t)b * 4 + 3] = tot;
    }
}

// ------------------------------------------------------------------------------------------------ finalize
__global__ void __launch_bounds__(256)
loss_finalize_kernel(const double* __restrict__ sums, const double* __restrict__ l1_partial, int n_l1, int B, int F,
                     int T, float* __restrict__ out4) {
    // 256 threads, strided partial sums folded in a fixed order (deterministic); one thread walking the B x 4 + n_l1 partials took 90 us
    __shared__ double red[8];
    __shared__ double tot[5];
    for (int j = 0; j < 5; ++j) {
        double a = 0.0;
        if (j < 4) { for (int b = threadIdx.x; b < B; b += blockDim.x) a += sums[(size_t)b * 4 + j]; }
        else { for (int i = threadIdx.x; i < n_l1; i += blockDim.x) a += l1_partial[i]; }
        const double t = block_sum(a, red);
        if (threadIdx.x == 0) tot[j] = t;
        __syncthreads();
    }
    if (threadIdx.x != 0) return;
    const double s[4] = {tot[0], tot[1], tot[2], tot[3]}, l1 = tot[4];
    const int nffts[3] = {63, 32, 16}, hops[3] = {16, 8, 4};
    double stft = 0.0;
    for (int j = 0; j < 3; ++j) {
        const int n = nffts[j], pad = n / 2, bins = n / 2 + 1;
        const int frames = 1 + (T + 2 * pad - n) / hops[j];
        stft += s[j] / ((double)B * bins * frames);
    }
    stft /= 3.0;
    const int mframes = 1 + (T + 2 * (MEL_NFFT / 2) - MEL_NFFT) / MEL_HOP;
    const double mel = s[3] / ((double)B * MEL_N * mframes);
    const double l1m = l1 / ((double)B * F * T);
    out4[0] = (float)(0.4 * stft + 0.4 * mel + 0.2 * l1m);
    out4[1] = (float)stft;
    out4[2] = (float)mel;
    out4[3] = (float)l1m;
}

// ------------------------------------------------------------------------------------------------ backward
// d(c_stft * stft + c_mel * mel + c_l1 * l1) / d pred.  Both spectral terms depend on pred only through the envelope
// e[b][t] = mean_f pred[b][f][t], so the kernel below produces dE (B, T) and the elementwise pass adds dE / F to the L1 term.
//   phase 1 (per transform): the forward DFT again, storing for every (frame, bin) the cotangent of (re, im):
//        rectangular scales: sign(|P| - |T|) * (re, im) / |P| * c_stft / (3 B bins frames)       (d|z| = 0 at z = 0, like torch)
//        mel:                2 (re, im) * sum_m sign(mel_p - mel_t)[m] fb[k][m] * c_mel / (B 64 frames)
//   phase 2: every envelope sample GATHERS its contributions (frames covering it x bins) in a fixed order: deterministic.
__global__ void __launch_bounds__(256)
loss_backward_env_kernel(const float* __restrict__ env, int T, int B, const float* __restrict__ mel_fb, float c_stft, float c_mel,
                         float* __restrict__ scratch, long long scratch_per_sample, float* __restrict__ d_env) {
    extern __shared__ float smem[];
    float* xp = smem;                       // [T]
    float* xt = xp + T;                     // [T]
    float* de = xt + T;                     // [T]
    float* tw_c = de + T;                   // [64]
    float* tw_s = tw_c + 64;                // [64]
    float* win = tw_s + 64;                 // [64]
    float* fb = win + 64;                   // [32][64]
    float* msgn = fb + MEL_BINS * MEL_N;    // [8 frames][64 mels] sign(mel_p - mel_t)
    float* pw = msgn + 8 * MEL_N;           // [2][8][32]
    const int b = blockIdx.x, tid = threadIdx.x;
    float* G = scratch + (long long)b * scratch_per_sample;
    for (int i = tid; i < T; i += blockDim.x) { xp[i] = env[((size_t)b * 2) * T + i]; xt[i] = env[((size_t)b * 2 + 1) * T + i]; de[i] = 0.f; }
    for (int i = tid; i < MEL_BINS * MEL_N; i += blockDim.x) fb[i] = mel_fb[i];
    if (tid < 64) win[tid] = tid < MEL_NFFT ? 0.5f - 0.5f * cospif(2.0f * (float)tid / (float)MEL_NFFT) : 0.f;

    const int nffts[3] = {63, 32, 16}, hops[3] = {16, 8, 4};
    for (int s = 0; s < 3; ++s) {
        const int n = nffts[s], hop = hops[s], pad = n / 2, bins = n / 2 + 1;
        const int frames = 1 + (T + 2 * pad - n) / hop;
        const float coef = c_stft / (3.0f * (float)B * (float)bins * (float)frames);
        __syncthreads();
        if (tid < n) sincospif(2.0f * (float)tid / (float)n, &tw_s[tid], &tw_c[tid]);
        __syncthreads();
        for (int item = tid; item < frames * bins; item += blockDim.x) {
            const int fr = item / bins, k = item - fr * bins;
            const int start = fr * hop - pad;
            float pr = 0.f, pi = 0.f, tr = 0.f, ti = 0.f;
            int m = 0;
            for (int i = 0; i < n; ++i) {
                const int idx = start + i;
                if (idx >= 0 && idx < T) {
                    const float c = tw_c[m], sn = tw_s[m];
                    const float a = xp[idx], g = xt[idx];
                    pr = fmaf(a, c, pr); pi = fmaf(a, sn, pi);
                    tr = fmaf(g, c, tr); ti = fmaf(g, sn, ti);
                }
                m += k; if (m >= n) m -= n;
            }
            const float mp = sqrtf(pr * pr + pi * pi), mt = sqrtf(tr * tr + ti * ti);
            const float sg = (mp > mt) ? 1.f : (mp < mt) ? -1.f : 0.f;
            const float w = mp > 0.f ? coef * sg / mp : 0.f;
            G[2 * item] = w * pr;
            G[2 * item + 1] = w * pi;
        }
        __syncthreads();
        for (int j = tid; j < T; j += blockDim.x) {
            // frames with start <= j < start + n, start = fr*hop - pad
            const int fr_lo = max(0, (j + pad - n + hop) / hop), fr_hi = min(frames - 1, (j + pad) / hop);
            float acc = 0.f;
            for (int fr = fr_lo; fr <= fr_hi; ++fr) {
                const int i = j - (fr * hop - pad);
                if (i < 0 || i >= n) continue;
                int m = 0;
                for (int k = 0; k < bins; ++k) {
                    acc = fmaf(G[2 * (fr * bins + k)], tw_c[m], acc);
                    acc = fmaf(G[2 * (fr * bins + k) + 1], tw_s[m], acc);
                    m += i; if (m >= n) m -= n;
                }
            }
            de[j] += acc;
        }
        __syncthreads();
    }
    {   // mel term
        const int pad = MEL_NFFT / 2;
        const int frames = 1 + (T + 2 * pad - MEL_NFFT) / MEL_HOP;
        const float coef = c_mel / ((float)B * (float)MEL_N * (float)frames);
        __syncthreads();
        if (tid < MEL_NFFT) sincospif(2.0f * (float)tid / (float)MEL_NFFT, &tw_s[tid], &tw_c[tid]);
        __syncthreads();
        for (int f0 = 0; f0 < frames; f0 += 8) {
            const int nf = min(8, frames - f0);
            for (int item = tid; item < 2 * nf * MEL_BINS; item += blockDim.x) {
                const int sig = item / (nf * MEL_BINS);
                const int rem = item - sig * nf * MEL_BINS;
                const int fr = rem / MEL_BINS, k = rem - fr * MEL_BINS;
                const float* x = sig ? xt : xp;
                const int start = (f0 + fr) * MEL_HOP - pad;
                float re = 0.f, im = 0.f;
                int m = 0;
                for (int i = 0; i < MEL_NFFT; ++i) {
                    int idx = start + i;
                    if (idx < 0) idx = -idx;
                    if (idx >= T) idx = 2 * (T - 1) - idx;
                    const float v = x[idx] * win[i];
                    re = fmaf(v, tw_c[m], re); im = fmaf(v, tw_s[m], im);
                    m += k; if (m >= MEL_NFFT) m -= MEL_NFFT;
                }
                pw[(sig * 8 + fr) * MEL_BINS + k] = re * re + im * im;
                if (sig == 0) { G[2 * ((f0 + fr) * MEL_BINS + k)] = re; G[2 * ((f0 + fr) * MEL_BINS + k) + 1] = im; }
            }
            __syncthreads();
            for (int item = tid; item < nf * MEL_N; item += blockDim.x) {
                const int fr = item / MEL_N, mel = item - fr * MEL_N;
                float mp = 0.f, mt = 0.f;
#pragma unroll 8
                for (int k = 0; k < MEL_BINS; ++k) {
                    const float w = fb[k * MEL_N + mel];
                    mp = fmaf(pw[fr * MEL_BINS + k], w, mp);
                    mt = fmaf(pw[(8 + fr) * MEL_BINS + k], w, mt);
                }
                msgn[fr * MEL_N + mel] = (mp > mt) ? 1.f : (mp < mt) ? -1.f : 0.f;
            }
            __syncthreads();
            for (int item = tid; item < nf * MEL_BINS; item += blockDim.x) {        // cotangent of (re, im) of the pred spectrum
                const int fr = item / MEL_BINS, k = item - fr * MEL_BINS;
                float g = 0.f;
#pragma unroll 8
                for (int mel = 0; mel < MEL_N; ++mel) g = fmaf(msgn[fr * MEL_N + mel], fb[k * MEL_N + mel], g);
                g *= 2.f * coef;
                G[2 * ((f0 + fr) * MEL_BINS + k)] *= g;
                G[2 * ((f0 + fr) * MEL_BINS + k) + 1] *= g;
            }
            __syncthreads();
        }
        for (int j = tid; j < T; j += blockDim.x) {
            // padded positions that reflect onto j: q = j, q = -j (j >= 1), q = 2(T-1) - j (j <= T-2)
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                int q;
                if (c == 0) q = j;
                else if (c == 1) { if (j < 1) continue; q = -j; }
                else { if (j > T - 2) continue; q = 2 * (T - 1) - j; }
                if (q < -pad || q > T - 1 + pad) continue;
                const int fr_lo = max(0, (q + pad - MEL_NFFT + MEL_HOP) / MEL_HOP), fr_hi = min(frames - 1, (q + pad) / MEL_HOP);
                for (int fr = fr_lo; fr <= fr_hi; ++fr) {
                    const int i = q - (fr * MEL_HOP - pad);
                    if (i < 0 || i >= MEL_NFFT) continue;
                    float a2 = 0.f;
                    int m = 0;
                    for (int k = 0; k < MEL_BINS; ++k) {
                        a2 = fmaf(G[2 * (fr * MEL_BINS + k)], tw_c[m], a2);
                        a2 = fmaf(G[2 * (fr * MEL_BINS + k) + 1], tw_s[m], a2);
                        m += i; if (m >= MEL_NFFT) m -= MEL_NFFT;
                    }
                    acc = fmaf(a2, win[i], acc);
                }
            }
            de[j] += acc;
        }
        __syncthreads();
    }
    for (int i = tid; i < T; i += blockDim.x) d_env[(size_t)b * T + i] = de[i];
}

// d_pred[b][f][t] = c_l1 * sign(pred - target) / (B F T) + d_env[b][t] / F
__global__ void __launch_bounds__(256)
loss_backward_apply_kernel(const float* __restrict__ pred, const float* __restrict__ target, const float* __restrict__ d_env, int F, int T,
                           long long total, float c_l1_over_n, float inv_f, float* __restrict__ d_pred) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(i % T);
        const long long b = i / ((long long)F * T);
        const float d = __ldcs(pred + i) - __ldcs(target + i);
        const float sg = (d > 0.f) ? 1.f : (d < 0.f) ? -1.f : 0.f;
        d_pred[i] = fmaf(sg, c_l1_over_n, d_env[b * T + t] * inv_f);
    }
}

static inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace adn

using namespace adn;

extern "C" int64_t adn_loss_workspace_bytes(int64_t batch, int freq, int frames) {
    if (batch < 0 || freq <= 0 || frames <= 0) return -1;
    const size_t env = align256((size_t)batch * 2 * frames * sizeof(float));
    const size_t l1 = align256((size_t)batch * ((frames + ENV_TX - 1) / ENV_TX) * sizeof(double));
    const size_t sums = align256((size_t)batch * 4 * sizeof(double));
    return (int64_t)(env + l1 + sums);
}

extern "C" int adn_combined_loss_f32(const float* pred, const float* target, int64_t batch, int freq, int frames,
                                     const float* mel_fb_32x64, void* workspace, float* out4, void* stream) {
    if (batch <= 0 || freq <= 0 || frames <= 0) return ADN_ERR_ARG;
    if (!pred || !target || !mel_fb_32x64 || !workspace || !out4) return ADN_ERR_ARG;
    if (frames > LOSS_MAX_T || frames <= MEL_NFFT / 2 || batch > 65535) return ADN_ERR_ARG;   // reflect padding needs T > 31
    int st = check_device();
    if (st != ADN_OK) return st;
    cudaStream_t s = (cudaStream_t)stream;
    char* ws = static_cast<char*>(workspace);
    float* env = reinterpret_cast<float*>(ws);
    const int tx = (frames + ENV_TX - 1) / ENV_TX;
    double* l1p = reinterpret_cast<double*>(ws + align256((size_t)batch * 2 * frames * sizeof(float)));
    double* sums = reinterpret_cast<double*>(reinterpret_cast<char*>(l1p) + align256((size_t)batch * tx * sizeof(double)));
    loss_envelope_kernel<<<dim3(tx, (unsigned)batch), dim3(ENV_TX, ENV_FY), 0, s>>>(pred, target, freq, frames, env, l1p);
    ADN_LAUNCH_CHECK();
    const size_t smem = (size_t)(2 * frames + 3 * 64 + MEL_BINS * MEL_N + 2 * 8 * MEL_BINS) * sizeof(float);
    static unsigned char smem_set[64] = {0};
    ADN_CUDA_TRY(ensure_dyn_smem(loss_spectral_kernel, (int)((2 * LOSS_MAX_T + 3 * 64 + MEL_BINS * MEL_N + 2 * 8 * MEL_BINS) * sizeof(float)), smem_set));
    loss_spectral_kernel<<<dim3((unsigned)batch, 4), 256, smem, s>>>(env, frames, mel_fb_32x64, sums);
    ADN_LAUNCH_CHECK();
    loss_finalize_kernel<<<1, 256, 0, s>>>(sums, l1p, (int)(batch * tx), (int)batch, freq, frames, out4);
    ADN_LAUNCH_CHECK();
    return ADN_OK;
}

static inline long long loss_bwd_scratch_floats(int T) {
    long long f63 = 1 + (T - 1) / 16, f32 = 1 + T / 8, f16 = 1 + T / 4;
    long long m = f63 * 32;
    if (f32 * 17 > m) m = f32 * 17;
    if (f16 * 9 > m) m = f16 * 9;
    return 2 * m + 16;
}

extern "C" int64_t adn_loss_backward_workspace_bytes(int64_t batch, int freq, int frames) {
    if (batch < 0 || freq <= 0 || frames <= 0) return -1;
    return (int64_t)(align256((size_t)batch * 2 * frames * sizeof(float)) + align256((size_t)batch * frames * sizeof(float)) +
                     align256((size_t)batch * loss_bwd_scratch_floats(frames) * sizeof(float)) +
                     align256((size_t)batch * ((frames + ENV_TX - 1) / ENV_TX) * sizeof(double)));
}

// Gradient of c_stft * stft + c_mel * mel + c_l1 * l1 (the three terms of CombinedPerceptualLoss, loss.py:83-95) with respect
// to pred; loss.backward() at train.py:69 is (c_stft, c_mel, c_l1) = (0.4, 0.4, 0.2).  d_pred: (batch,1,freq,frames) float32.
extern "C" int adn_combined_loss_backward_f32(const float* pred, const float* target, int64_t batch, int freq, int frames,
                                              const float* mel_fb_32x64, float c_stft, float c_mel, float c_l1, void* workspace,
                                              float* d_pred, void* stream) {
    if (batch <= 0 || freq <= 0 || frames <= 0) return ADN_ERR_ARG;
    if (!pred || !target || !mel_fb_32x64 || !workspace || !d_pred) return ADN_ERR_ARG;
    if (frames > LOSS_MAX_T / 2 || frames <= MEL_NFFT / 2 || batch > 65535) return ADN_ERR_ARG;
    int st = check_device();
    if (st != ADN_OK) return st;
    cudaStream_t s = (cudaStream_t)stream;
    char* ws = static_cast<char*>(workspace);
    float* env = reinterpret_cast<float*>(ws); ws += align256((size_t)batch * 2 * frames * sizeof(float));
    float* d_env = reinterpret_cast<float*>(ws); ws += align256((size_t)batch * frames * sizeof(float));
    float* scratch = reinterpret_cast<float*>(ws); ws += align256((size_t)batch * loss_bwd_scratch_floats(frames) * sizeof(float));
    double* l1p = reinterpret_cast<double*>(ws);
    const int tx = (frames + ENV_TX - 1) / ENV_TX;
    loss_envelope_kernel<<<dim3(tx, (unsigned)batch), dim3(ENV_TX, ENV_FY), 0, s>>>(pred, target, freq, frames, env, l1p);
    ADN_LAUNCH_CHECK();
    const size_t smem = (size_t)(3 * frames + 3 * 64 + MEL_BINS * MEL_N + 8 * MEL_N + 2 * 8 * MEL_BINS) * sizeof(float);
    static unsigned char smem_set[64] = {0};
    ADN_CUDA_TRY(ensure_dyn_smem(loss_backward_env_kernel, (int)((3 * (LOSS_MAX_T / 2) + 3 * 64 + MEL_BINS * MEL_N + 8 * MEL_N + 2 * 8 * MEL_BINS) * sizeof(float)), smem_set));
    loss_backward_env_kernel<<<(unsigned)batch, 256, smem, s>>>(env, frames, (int)batch, mel_fb_32x64, c_stft, c_mel, scratch,
                                                               loss_bwd_scratch_floats(frames), d_env);
    ADN_LAUNCH_CHECK();
    const long long total = (long long)batch * freq * frames;
    long long g = (total + 255) / 256; const long long cap = (long long)num_sms() * 16; if (g > cap) g = cap;
    loss_backward_apply_kernel<<<(int)g, 256, 0, s>>>(pred, target, d_env, freq, frames, total, c_l1 / (float)((double)batch * freq * frames),
                                                     1.0f / (float)freq, d_pred);
    ADN_LAUNCH_CHECK();
    return ADN_OK;
}
