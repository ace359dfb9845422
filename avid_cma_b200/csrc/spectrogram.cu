// Log-spectrogram of the audio clips on the GPU (include/avid_b200.h, avid_log_spectrogram): the input-side step the reference
// runs in its CPU loader workers with librosa (datasets/preprocessing.py:158-186, LogSpectrogram.__call__):
//     S = |stft(sig, n_fft = 2 * bins_arg, hop)|^2                       (hann window, centred frames, reflect padding)
//     S = [S[0], mean of the bin pairs (S[1], S[2]), (S[3], S[4]), ...]  -> n_fft / 4 + 1 bins
//     S = S[:, :num_frames];  dB = 10 log10(max(S, 1e-10));  dB = max(dB, max(dB) - top_db)       (power_to_db, ref = 1)
//     out = (dB - mean[bin]) / (std[bin] + 1e-5), transposed to (1, frames, bins)
// One CTA per (frame, clip): the windowed frame goes through a radix-2 FFT in shared memory (n_fft <= 2048); a second pass applies
// the per-clip top_db floor and the per-bin normalisation.  fp32 throughout (librosa computes in complex64 as well).
#include <math.h>
#include "common.cuh"

namespace avid {

constexpr int kSpecMaxFft = 2048;

__device__ __forceinline__ int reflect_index(int i, int n) {      // numpy pad mode 'reflect' (no edge repeat), |i| may exceed n once
    if (n == 1) return 0;
    const int period = 2 * (n - 1);
    i %= period;
    if (i < 0) i += period;
    return i < n ? i : period - i;
}

// monotone float <-> int map so that atomicMax on ints orders floats
__device__ __forceinline__ int float_order(float f) {
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float order_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void __launch_bounds__(256) logspec_power_kernel(const float* __restrict__ wave, int num_samples, int n_fft, int log2n, int hop,
                                                            int num_frames, float* __restrict__ out, int* __restrict__ clip_max) {
    __shared__ float2 x[kSpecMaxFft];
    __shared__ float2 tw[kSpecMaxFft / 2];
    __shared__ float s_max[8];
    const int f = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const float* sig = wave + (size_t)b * num_samples;
    const int start = f * hop - n_fft / 2;                          // librosa centre=True: frame f is centred on sample f * hop
    for (int i = tid; i < n_fft; i += 256) {
        const float w = 0.5f - 0.5f * cospif(2.0f * (float)i / (float)n_fft);      // periodic hann (scipy get_window('hann', fftbins=True))
        const float v = sig[reflect_index(start + i, num_samples)] * w;
        x[__brev((unsigned)i) >> (32 - log2n)] = make_float2(v, 0.f);              // bit-reversed order for the in-place DIT FFT
    }
    for (int i = tid; i < n_fft / 2; i += 256) {
        float s, c;
        sincospif(-2.0f * (float)i / (float)n_fft, &s, &c);
        tw[i] = make_float2(c, s);
    }
    __syncthreads();
    for (int len = 2, stride = n_fft / 2; len <= n_fft; len <<= 1, stride >>= 1) {
        const int half = len >> 1;
        for (int i = tid; i < n_fft / 2; i += 256) {
            const int grp = i / half, pos = i - grp * half;
            const int a = grp * len + pos, c2 = a + half;
            const float2 w = tw[pos * stride];
            const float2 u = x[a], v = x[c2];
            const float2 t = make_float2(v.x * w.x - v.y * w.y, v.x * w.y + v.y * w.x);
            x[a] = make_float2(u.x + t.x, u.y + t.y);
            x[c2] = make_float2(u.x - t.x, u.y - t.y);
        }
        __syncthreads();
    }
    // bins: [P0, mean(P1, P2), mean(P3, P4), ...]  (preprocessing.py:170-171)
    const int bins = n_fft / 4 + 1;
    float mx = -INFINITY;
    float* dst = out + ((size_t)b * num_frames + f) * bins;
    for (int j = tid; j < bins; j += 256) {
        float p;
        if (j == 0) {
            p = x[0].x * x[0].x + x[0].y * x[0].y;
        } else {
            const float2 a = x[2 * j - 1], c2 = x[2 * j];
            p = 0.5f * ((a.x * a.x + a.y * a.y) + (c2.x * c2.x + c2.y * c2.y));
        }
        const float db = 10.0f * log10f(fmaxf(p, 1e-10f));
        dst[j] = db;
        mx = fmaxf(mx, db);
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    if ((tid & 31) == 0) s_max[tid >> 5] = mx;
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < 8; ++w) mx = fmaxf(mx, s_max[w]);
        atomicMax(clip_max + b, float_order(mx));
    }
}

__global__ void logspec_init_kernel(int* clip_max, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) clip_max[i] = float_order(-INFINITY);
}

__global__ void __launch_bounds__(256) logspec_finish_kernel(float* __restrict__ out, const int* __restrict__ clip_max, int per_clip, int bins,
                                                             float top_db, const float* __restrict__ mean, const float* __restrict__ stdv, int64_t total) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int b = (int)(i / per_clip), bin = (int)(i % bins);
        float v = out[i];
        if (top_db >= 0.f) v = fmaxf(v, order_float(clip_max[b]) - top_db);
        if (mean) v = (v - mean[bin]) / (stdv[bin] + 1e-5f);
        out[i] = v;
    }
}

}  // namespace avid

using namespace avid;

extern "C" {

size_t avid_log_spectrogram_workspace_bytes(int32_t batch) { return batch > 0 ? sizeof(int) * (size_t)batch : 0; }

int avid_log_spectrogram(const float* wave, int32_t batch, int32_t num_samples, int32_t n_fft, int32_t hop, int32_t num_frames, float top_db,
                         const float* mean, const float* stdv, float* out, void* workspace, size_t workspace_bytes, void* stream) {
    AVID_REQUIRE(wave && out && batch > 0 && num_samples > 1 && hop > 0 && num_frames > 0, "log_spectrogram: bad arguments");
    AVID_REQUIRE(n_fft >= 8 && n_fft <= kSpecMaxFft && (n_fft & (n_fft - 1)) == 0, "log_spectrogram: n_fft=%d must be a power of two in [8, %d]", n_fft,
                 kSpecMaxFft);
    AVID_REQUIRE((mean == nullptr) == (stdv == nullptr), "log_spectrogram: give both mean and std or neither");
    AVID_REQUIRE(num_frames <= num_samples / hop + 1, "log_spectrogram: %d frames requested, the clip has %d", num_frames, num_samples / hop + 1);
    AVID_REQUIRE(n_fft / 2 < num_samples, "log_spectrogram: reflect padding needs more than n_fft / 2 samples");
    AVID_REQUIRE(workspace && workspace_bytes >= avid_log_spectrogram_workspace_bytes(batch), "log_spectrogram: workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int* clip_max = static_cast<int*>(workspace);
    int log2n = 0;
    while ((1 << log2n) < n_fft) ++log2n;
    logspec_init_kernel<<<(batch + 255) / 256, 256, 0, st>>>(clip_max, batch);
    int rc;
    if ((rc = check_launch("logspec_init_kernel"))) return rc;
    logspec_power_kernel<<<dim3(num_frames, batch), 256, 0, st>>>(wave, num_samples, n_fft, log2n, hop, num_frames, out, clip_max);
    if ((rc = check_launch("logspec_power_kernel"))) return rc;
    const int bins = n_fft / 4 + 1;
    const int64_t total = (int64_t)batch * num_frames * bins;
    int64_t blocks = (total + 255) / 256;
    if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
    logspec_finish_kernel<<<(unsigned)blocks, 256, 0, st>>>(out, clip_max, num_frames * bins, bins, top_db, mean, stdv, total);
    return check_launch("logspec_finish_kernel");
}

}  // extern "C"
