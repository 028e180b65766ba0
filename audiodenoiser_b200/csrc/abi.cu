// abi.cu -- status / error plumbing of the C ABI and the host-buffer convenience entry points.
#include <cstdlib>
#include <mutex>
#include "adn_common.cuh"

namespace adn {

static thread_local int tl_last_cuda_error = 0;

void set_last_cuda_error(cudaError_t e) { tl_last_cuda_error = (int)e; }

int check_device() {
    // cached per device ordinal; the answer never changes for a device, so a benign race only repeats the query
    static int cache[64] = {0};   // 0 = unknown, 1 = ok, 2 = unsupported
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { set_last_cuda_error(e); return ADN_ERR_CUDA; }
    if (dev >= 0 && dev < 64 && cache[dev]) return cache[dev] == 1 ? ADN_OK : ADN_ERR_DEVICE;
    int major = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) { set_last_cuda_error(e); return ADN_ERR_CUDA; }
    const int ok = (major == 10);
    if (dev >= 0 && dev < 64) cache[dev] = ok ? 1 : 2;
    return ok ? ADN_OK : ADN_ERR_DEVICE;
}

}  // namespace adn

extern "C" int adn_version(void) { return 200; }

#ifndef ADN_SOURCE_HASH
#define ADN_SOURCE_HASH "unstamped"
#endif
extern "C" const char* adn_source_hash(void) { return ADN_SOURCE_HASH; }

extern "C" const char* adn_error_string(int status) {
    switch (status) {
        case ADN_OK: return "ok";
        case ADN_ERR_ARG: return "invalid argument (size, alignment, null pointer or unsupported shape)";
        case ADN_ERR_CUDA: return "CUDA runtime error (see adn_last_cuda_error)";
        case ADN_ERR_DEVICE: return "current CUDA device is not sm_100 (B200); there is no fallback path";
        case ADN_ERR_DRIVER: return "cuTensorMapEncodeTiled unavailable or failed";
        case ADN_ERR_SHORT: return "input is too short for n_fft=512 with center=0";
        default: return "unknown status";
    }
}

extern "C" int adn_last_cuda_error(void) { return adn::tl_last_cuda_error; }

extern "C" int adn_device_check(void) { return adn::check_device(); }

// ------------------------------------------------------------------------------------------------ host variants
namespace {
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
};
}  // namespace

extern "C" int adn_stft_mag_host_f32(const float* wave_host, int64_t n_clips, int64_t length, int center, float* mag_host) {
    const int64_t T = adn_stft_num_frames(length, center);
    if (n_clips < 0) return ADN_ERR_ARG;
    if (T < 0) return ADN_ERR_SHORT;
    if (n_clips == 0) return ADN_OK;
    if (!wave_host || !mag_host) return ADN_ERR_ARG;
    int st = adn::check_device();
    if (st != ADN_OK) return st;
    // pad the row stride to a multiple of 4 floats so every clip takes the 128-bit load path
    const int64_t stride = (length + 3) & ~(int64_t)3;
    DevBuf w, m;
    ADN_CUDA_TRY(w.alloc((size_t)n_clips * (stride ? stride : 4) * sizeof(float)));
    ADN_CUDA_TRY(m.alloc((size_t)n_clips * ADN_N_BINS * T * sizeof(float)));
    cudaStream_t s = nullptr;   // legacy default stream: synchronous with respect to the caller
    if (length > 0)
        ADN_CUDA_TRY(cudaMemcpy2DAsync(w.p, stride * sizeof(float), wave_host, length * sizeof(float), length * sizeof(float),
                                       (size_t)n_clips, cudaMemcpyHostToDevice, s));
    st = adn_stft_mag_f32((const float*)w.p, n_clips, length, stride ? stride : 4, center, (float*)m.p, s);
    if (st != ADN_OK) return st;
    ADN_CUDA_TRY(cudaMemcpyAsync(mag_host, m.p, (size_t)n_clips * ADN_N_BINS * T * sizeof(float), cudaMemcpyDeviceToHost, s));
    ADN_CUDA_TRY(cudaStreamSynchronize(s));
    return ADN_OK;
}

extern "C" int adn_istft_ola_host_f32(const float* mag_host, const float* phasor_c64_host, uint64_t seed, int64_t n_clips,
                                      int64_t n_frames, float* audio_host) {
    if (n_clips < 0 || n_frames < 1) return ADN_ERR_ARG;
    if (n_clips == 0 || n_frames == 1) return ADN_OK;
    if (!mag_host || !audio_host) return ADN_ERR_ARG;
    int st = adn::check_device();
    if (st != ADN_OK) return st;
    const size_t n_spec = (size_t)n_clips * ADN_N_BINS * n_frames;
    const size_t n_out = (size_t)n_clips * ADN_HOP * (n_frames - 1);
    DevBuf m, p, a;
    ADN_CUDA_TRY(m.alloc(n_spec * sizeof(float)));
    ADN_CUDA_TRY(a.alloc(n_out * sizeof(float)));
    cudaStream_t s = nullptr;
    ADN_CUDA_TRY(cudaMemcpyAsync(m.p, mag_host, n_spec * sizeof(float), cudaMemcpyHostToDevice, s));
    if (phasor_c64_host) {
        ADN_CUDA_TRY(p.alloc(n_spec * 2 * sizeof(float)));
        ADN_CUDA_TRY(cudaMemcpyAsync(p.p, phasor_c64_host, n_spec * 2 * sizeof(float), cudaMemcpyHostToDevice, s));
    }
    st = adn_istft_ola_f32((const float*)m.p, phasor_c64_host ? (const float*)p.p : nullptr, 0, seed, n_clips, n_frames,
                           (float*)a.p, s);
    if (st != ADN_OK) return st;
    ADN_CUDA_TRY(cudaMemcpyAsync(audio_host, a.p, n_out * sizeof(float), cudaMemcpyDeviceToHost, s));
    ADN_CUDA_TRY(cudaStreamSynchronize(s));
    return ADN_OK;
}

// cudaMemsetAsync behind the C ABI: the statistics vector of a step is zeroed by a memset node, not by a framework fill kernel
extern "C" int adn_zero_bytes(void* dev, int64_t bytes, void* stream) {
    if (!dev || bytes < 0) return ADN_ERR_ARG;
    if (bytes == 0) return ADN_OK;
    ADN_CUDA_TRY(cudaMemsetAsync(dev, 0, (size_t)bytes, (cudaStream_t)stream));
    return ADN_OK;
}

// cudaMemcpyAsync (device to device) behind the C ABI: small bookkeeping copies of the training step as memcpy nodes
extern "C" int adn_copy_bytes(void* dst_dev, const void* src_dev, int64_t bytes, void* stream) {
    if (!dst_dev || !src_dev || bytes < 0) return ADN_ERR_ARG;
    if (bytes == 0) return ADN_OK;
    ADN_CUDA_TRY(cudaMemcpyAsync(dst_dev, src_dev, (size_t)bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return ADN_OK;
}

namespace adn {
__global__ void i64_add_n_kernel(long long* v, int n, long long inc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] += inc;
}
}  // namespace adn

// v[0..n) += inc on the device (the num_batches_tracked counters of the 18 BatchNorm2d layers, one launch per training step)
extern "C" int adn_i64_add_n(int64_t* v_dev, int n, int64_t inc, void* stream) {
    if (!v_dev || n <= 0) return ADN_ERR_ARG;
    int st = adn::check_device();
    if (st != ADN_OK) return st;
    adn::i64_add_n_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(reinterpret_cast<long long*>(v_dev), n, (long long)inc);
    ADN_LAUNCH_CHECK();
    return ADN_OK;
}
