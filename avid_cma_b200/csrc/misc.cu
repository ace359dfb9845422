// Layout conversions, the fused Adam step and small elementwise helpers (include/avid_b200.h).
#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>
#include "common.cuh"

namespace avid {

// [n, c, thw] -> [n, thw, cpad] (channels-last, zero-padded to cpad channels), c <= 4
__global__ void __launch_bounds__(256) nchw_to_nhwc_small_kernel(const float* __restrict__ in, float* __restrict__ out, int n, int c,
                                                                 int64_t thw, int cpad) {
    const int64_t total = (int64_t)n * thw;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t img = i / thw, s = i - img * thw;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        for (int ch = 0; ch < c; ++ch) v[ch] = __ldg(in + (img * c + ch) * thw + s);
        for (int q = 0; q < cpad; q += 4)
            *reinterpret_cast<float4*>(out + i * cpad + q) = q == 0 ? make_float4(v[0], v[1], v[2], v[3]) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// generic batched 2-D transpose through shared memory: in [b, rows, cols] -> out [b, cols, rows]
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t rows, int64_t cols) {
    __shared__ float tile[32][33];
    const int64_t b = blockIdx.z;
    const float* src = in + (size_t)b * rows * cols;
    float* dst = out + (size_t)b * rows * cols;
    const int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (int j = ty; j < 32; j += 8)
        if (r0 + j < rows && c0 + tx < cols) tile[j][tx] = src[(r0 + j) * cols + c0 + tx];
    __syncthreads();
    for (int j = ty; j < 32; j += 8)
        if (c0 + j < cols && r0 + tx < rows) dst[(c0 + j) * rows + r0 + tx] = tile[tx][j];
}

// PyTorch filter [co, ci, taps] -> tap-major [taps, ci_pad, co] and its transpose [taps, co, ci_pad]
__global__ void filter_to_tap_kernel(const float* __restrict__ w, float* __restrict__ w_tap, float* __restrict__ w_tap_t, int co, int ci,
                                     int taps, int ci_pad) {
    const int64_t total = (int64_t)taps * ci_pad * co;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int o = (int)(i % co);
        int64_t r = i / co;
        const int c = (int)(r % ci_pad), t = (int)(r / ci_pad);
        const float v = c < ci ? __ldg(w + ((size_t)o * ci + c) * taps + t) : 0.f;
        w_tap[i] = v;
        if (w_tap_t) w_tap_t[((size_t)t * co + o) * ci_pad + c] = v;
    }
}

// tap-major [taps, ci_pad, co] -> PyTorch [co, ci, taps]
__global__ void filter_from_tap_kernel(const float* __restrict__ w_tap, float* __restrict__ w, int co, int ci, int taps, int ci_pad) {
    const int64_t total = (int64_t)co * ci * taps;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int t = (int)(i % taps);
        int64_t r = i / taps;
        const int c = (int)(r % ci), o = (int)(r / ci);
        w[i] = __ldg(w_tap + ((size_t)t * ci_pad + c) * co + o);
    }
}

__global__ void __launch_bounds__(256) add_inplace_kernel(float* __restrict__ a, const float* __restrict__ b, int64_t n4) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 x = reinterpret_cast<float4*>(a)[i];
        const float4 y = __ldg(reinterpret_cast<const float4*>(b) + i);
        x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;
        reinterpret_cast<float4*>(a)[i] = x;
    }
}

// torch.optim.Adam (L2 weight decay folded into the gradient, no amsgrad) on flat fp32 buffers
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, int64_t n, float lr, float beta1, float beta2, float eps,
                                                   float weight_decay, float bc1, float bc2_sqrt, float grad_scale) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float pi = p[i];
        const float gi = fmaf(weight_decay, pi, g[i] * grad_scale);
        const float mi = beta1 * m[i] + (1.f - beta1) * gi;
        const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = pi - (lr / bc1) * (mi / denom);
    }
}

// PyTorch filter [co, ci, taps] -> bf16 (hi, lo) planes of both tensor-core operand layouts in one pass:
//   fwd planes [taps][co][ci] (K-major in ci), dgrad planes [taps][ci][co] (K-major in co)
__global__ void __launch_bounds__(256) filter_to_planes_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ f_hi,
                                                               __nv_bfloat16* __restrict__ f_lo, __nv_bfloat16* __restrict__ d_hi,
                                                               __nv_bfloat16* __restrict__ d_lo, int co, int ci, int taps) {
    // two sweeps, each in the index order of the plane it WRITES: 2-byte stores with a stride of co (or ci) elements fill one
    // 32-byte sector per element, strided 4-byte reads of the (L2-resident, <= 9.4 MB) filter are much cheaper
    const int64_t total = (int64_t)taps * ci * co;
    const int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = i0; i < total; i += stride) {           // forward planes: ((t * co + o) * ci + c)
        const int c = (int)(i % ci);
        int64_t r = i / ci;
        const int o = (int)(r % co), t = (int)(r / co);
        const float v = __ldg(w + ((size_t)o * ci + c) * taps + t);
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        f_hi[i] = h;
        if (f_lo) f_lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
    for (int64_t j = i0; j < total; j += stride) {           // dgrad planes: ((t * ci + c) * co + o)
        const int o = (int)(j % co);
        int64_t r = j / co;
        const int c = (int)(r % ci), t = (int)(r / ci);
        const float v = __ldg(w + ((size_t)o * ci + c) * taps + t);
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        d_hi[j] = h;
        if (d_lo) d_lo[j] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
}

// The same two conversions for MANY filters in one launch (a tower converts ~35 filters per step; as separate launches they cost
// more in launch latency than in work).  Block b works on a 2048-element chunk of the filter whose block range contains b.
constexpr int kFilterMax = 48, kFilterChunk = 2048;
struct FilterBatch {
    const float* src[kFilterMax];     // to_planes: PyTorch filter [co][ci][taps];  from_tap: tap-major gradient [taps][ci_pad][co]
    void* dst[4][kFilterMax];         // to_planes: f_hi, f_lo, d_hi, d_lo planes;  from_tap: dst[0] = PyTorch-layout gradient
    int co[kFilterMax], ci[kFilterMax], taps[kFilterMax], ci_pad[kFilterMax];
    int block_begin[kFilterMax + 1];
    int count;
};

__global__ void __launch_bounds__(256) filter_to_planes_multi_kernel(const FilterBatch b) {
    int t = 0;
    while (t + 1 < b.count && (int)blockIdx.x >= b.block_begin[t + 1]) ++t;
    const int co = b.co[t], ci = b.ci[t], taps = b.taps[t];
    const float* __restrict__ w = b.src[t];
    __nv_bfloat16* f_hi = static_cast<__nv_bfloat16*>(b.dst[0][t]);
    __nv_bfloat16* f_lo = static_cast<__nv_bfloat16*>(b.dst[1][t]);
    __nv_bfloat16* d_hi = static_cast<__nv_bfloat16*>(b.dst[2][t]);
    __nv_bfloat16* d_lo = static_cast<__nv_bfloat16*>(b.dst[3][t]);
    const int total = taps * ci * co;
    const int base = (blockIdx.x - b.block_begin[t]) * kFilterChunk;
    const int end = min(total, base + kFilterChunk);
    for (int i = base + threadIdx.x; i < end; i += 256) {         // forward planes ((t * co + o) * ci + c)
        const int c = i % ci, r = i / ci;
        const int o = r % co, tp = r / co;
        const float v = __ldg(w + ((size_t)o * ci + c) * taps + tp);
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        f_hi[i] = h;
        if (f_lo) f_lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
    for (int j = base + threadIdx.x; j < end; j += 256) {         // dgrad planes ((t * ci + c) * co + o)
        const int o = j % co, r = j / co;
        const int c = r % ci, tp = r / ci;
        const float v = __ldg(w + ((size_t)o * ci + c) * taps + tp);
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        d_hi[j] = h;
        if (d_lo) d_lo[j] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
}

__global__ void __launch_bounds__(256) filter_from_tap_multi_kernel(const FilterBatch b) {
    int t = 0;
    while (t + 1 < b.count && (int)blockIdx.x >= b.block_begin[t + 1]) ++t;
    const int co = b.co[t], ci = b.ci[t], taps = b.taps[t], ci_pad = b.ci_pad[t];
    const float* __restrict__ w_tap = b.src[t];
    float* __restrict__ w = static_cast<float*>(b.dst[0][t]);
    const int total = co * ci * taps;
    const int base = (blockIdx.x - b.block_begin[t]) * kFilterChunk;
    const int end = min(total, base + kFilterChunk);
    for (int i = base + threadIdx.x; i < end; i += 256) {
        const int tp = i % taps, r = i / taps;
        const int c = r % ci, o = r / ci;
        w[i] = __ldg(w_tap + ((size_t)tp * ci_pad + c) * co + o);
    }
}

// torch.optim.Adam on up to kAdamMaxTensors tensors per launch: block b works on a 4096-element chunk of the tensor whose
// block range contains b
constexpr int kAdamMaxTensors = 32, kAdamChunk = 4096;
struct AdamBatch {
    float* p[kAdamMaxTensors];
    const float* g[kAdamMaxTensors];
    float* m[kAdamMaxTensors];
    float* v[kAdamMaxTensors];
    int64_t n[kAdamMaxTensors];
    int block_begin[kAdamMaxTensors + 1];
    int count;
};
__global__ void __launch_bounds__(256) adam_multi_kernel(const AdamBatch b, float lr, float beta1, float beta2, float eps, float weight_decay,
                                                         float bc1, float bc2_sqrt, float grad_scale) {
    int t = 0;
    while (t + 1 < b.count && (int)blockIdx.x >= b.block_begin[t + 1]) ++t;
    const int64_t base = (int64_t)(blockIdx.x - b.block_begin[t]) * kAdamChunk;
    const int64_t end = min(b.n[t], base + kAdamChunk);
    float* __restrict__ p = b.p[t];
    const float* __restrict__ g = b.g[t];
    float* __restrict__ m = b.m[t];
    float* __restrict__ v = b.v[t];
    for (int64_t i = base + threadIdx.x; i < end; i += 256) {
        const float pi = p[i];
        const float gi = fmaf(weight_decay, pi, g[i] * grad_scale);
        const float mi = beta1 * m[i] + (1.f - beta1) * gi;
        const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = pi - (lr / bc1) * (mi / denom);
    }
}

static unsigned grid_for(int64_t n) {
    int64_t b = (n + 255) / 256;
    const int64_t cap = 16 * kNumSMs;
    return (unsigned)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace avid

using namespace avid;

extern "C" {

int avid_nchw_to_nhwc(const float* in, float* out, int32_t n, int32_t c, int64_t thw, int32_t c_pad, void* stream) {
    AVID_REQUIRE(in && out && n > 0 && c > 0 && thw > 0 && c_pad >= c, "nchw_to_nhwc: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (c <= 4) {
        AVID_REQUIRE(c_pad % 4 == 0, "nchw_to_nhwc: c_pad=%d must be a multiple of 4", c_pad);
        nchw_to_nhwc_small_kernel<<<grid_for((int64_t)n * thw), 256, 0, st>>>(in, out, n, c, thw, c_pad);
        return check_launch("nchw_to_nhwc_small_kernel");
    }
    AVID_REQUIRE(c_pad == c, "nchw_to_nhwc: channel padding is only supported for c <= 4");
    AVID_REQUIRE((thw + 31) / 32 < 65536 && n < 65536, "nchw_to_nhwc: tensor too large");
    transpose_kernel<<<dim3((unsigned)((thw + 31) / 32), (c + 31) / 32, n), 256, 0, st>>>(in, out, c, thw);
    return check_launch("transpose_kernel");
}

int avid_nhwc_to_nchw(const float* in, float* out, int32_t n, int32_t c, int64_t thw, void* stream) {
    AVID_REQUIRE(in && out && n > 0 && c > 0 && thw > 0, "nhwc_to_nchw: bad arguments");
    AVID_REQUIRE((thw + 31) / 32 < 65536 && n < 65536, "nhwc_to_nchw: tensor too large");
    transpose_kernel<<<dim3((c + 31) / 32, (unsigned)((thw + 31) / 32), n), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, out, thw, c);
    return check_launch("transpose_kernel");
}

int avid_filter_to_tapmajor(const float* w_oihw, float* w_tap, float* w_tap_t, int32_t co, int32_t ci, int32_t taps, int32_t ci_pad, void* stream) {
    AVID_REQUIRE(w_oihw && w_tap && co > 0 && ci > 0 && taps > 0 && ci_pad >= ci, "filter_to_tapmajor: bad arguments");
    filter_to_tap_kernel<<<grid_for((int64_t)taps * ci_pad * co), 256, 0, static_cast<cudaStream_t>(stream)>>>(w_oihw, w_tap, w_tap_t, co, ci, taps, ci_pad);
    return check_launch("filter_to_tap_kernel");
}

int avid_filter_from_tapmajor(const float* w_tap, float* w_oihw, int32_t co, int32_t ci, int32_t taps, int32_t ci_pad, void* stream) {
    AVID_REQUIRE(w_oihw && w_tap && co > 0 && ci > 0 && taps > 0 && ci_pad >= ci, "filter_from_tapmajor: bad arguments");
    filter_from_tap_kernel<<<grid_for((int64_t)taps * ci * co), 256, 0, static_cast<cudaStream_t>(stream)>>>(w_tap, w_oihw, co, ci, taps, ci_pad);
    return check_launch("filter_from_tap_kernel");
}

int avid_filter_to_planes(const float* w_oihw, void* fwd_hi, void* fwd_lo, void* dgrad_hi, void* dgrad_lo, int32_t co, int32_t ci, int32_t taps,
                          void* stream) {
    AVID_REQUIRE(w_oihw && fwd_hi && dgrad_hi && co > 0 && ci > 0 && taps > 0, "filter_to_planes: bad arguments");
    AVID_REQUIRE((fwd_lo == nullptr) == (dgrad_lo == nullptr), "filter_to_planes: give both lo planes or neither");
    filter_to_planes_kernel<<<grid_for((int64_t)taps * ci * co), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        w_oihw, static_cast<__nv_bfloat16*>(fwd_hi), static_cast<__nv_bfloat16*>(fwd_lo), static_cast<__nv_bfloat16*>(dgrad_hi),
        static_cast<__nv_bfloat16*>(dgrad_lo), co, ci, taps);
    return check_launch("filter_to_planes_kernel");
}

int avid_adam_step_multi(float* const* params, const float* const* grads, float* const* exp_avgs, float* const* exp_avg_sqs,
                         const int64_t* sizes, int32_t count, int64_t step, float lr, float beta1, float beta2, float eps,
                         float weight_decay, float grad_scale, void* stream) {
    AVID_REQUIRE(params && grads && exp_avgs && exp_avg_sqs && sizes && count > 0 && step > 0, "adam_step_multi: bad arguments");
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    for (int first = 0; first < count; first += kAdamMaxTensors) {
        AdamBatch b;
        b.count = count - first < kAdamMaxTensors ? count - first : kAdamMaxTensors;
        int blocks = 0;
        for (int i = 0; i < b.count; ++i) {
            AVID_REQUIRE(params[first + i] && grads[first + i] && exp_avgs[first + i] && exp_avg_sqs[first + i] && sizes[first + i] > 0,
                         "adam_step_multi: tensor %d has a NULL pointer or no elements", first + i);
            b.p[i] = params[first + i];  b.g[i] = grads[first + i];  b.m[i] = exp_avgs[first + i];  b.v[i] = exp_avg_sqs[first + i];
            b.n[i] = sizes[first + i];
            b.block_begin[i] = blocks;
            blocks += (int)((sizes[first + i] + kAdamChunk - 1) / kAdamChunk);
        }
        b.block_begin[b.count] = blocks;
        adam_multi_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(b, lr, beta1, beta2, eps, weight_decay, (float)bc1,
                                                                                (float)sqrt(bc2), grad_scale);
        int rc = check_launch("adam_multi_kernel");
        if (rc) return rc;
    }
    return AVID_OK;
}

static int fill_filter_batch(FilterBatch* b, int first, int count, const float* const* src, const int32_t* co, const int32_t* ci, const int32_t* taps,
                             const int32_t* ci_pad, const char* what) {
    b->count = count - first < kFilterMax ? count - first : kFilterMax;
    int blocks = 0;
    for (int i = 0; i < b->count; ++i) {
        const int k = first + i;
        AVID_REQUIRE(src[k] && co[k] > 0 && ci[k] > 0 && taps[k] > 0, "%s: filter %d has a NULL pointer or an empty shape", what, k);
        const int64_t n = (int64_t)co[k] * ci[k] * taps[k];
        AVID_REQUIRE(n < ((int64_t)1 << 31), "%s: filter %d too large", what, k);
        b->src[i] = src[k];  b->co[i] = co[k];  b->ci[i] = ci[k];  b->taps[i] = taps[k];  b->ci_pad[i] = ci_pad ? ci_pad[k] : ci[k];
        b->block_begin[i] = blocks;
        blocks += (int)((n + kFilterChunk - 1) / kFilterChunk);
    }
    b->block_begin[b->count] = blocks;
    return AVID_OK;
}

int avid_filter_to_planes_multi(const float* const* w_oihw, void* const* fwd_hi, void* const* fwd_lo, void* const* dgrad_hi, void* const* dgrad_lo,
                                const int32_t* co, const int32_t* ci, const int32_t* taps, int32_t count, void* stream) {
    AVID_REQUIRE(w_oihw && fwd_hi && dgrad_hi && co && ci && taps && count > 0, "filter_to_planes_multi: bad arguments");
    AVID_REQUIRE((fwd_lo == nullptr) == (dgrad_lo == nullptr), "filter_to_planes_multi: give both lo plane lists or neither");
    for (int first = 0; first < count; first += kFilterMax) {
        FilterBatch b;
        int rc = fill_filter_batch(&b, first, count, w_oihw, co, ci, taps, nullptr, "filter_to_planes_multi");
        if (rc) return rc;
        for (int i = 0; i < b.count; ++i) {
            AVID_REQUIRE(fwd_hi[first + i] && dgrad_hi[first + i], "filter_to_planes_multi: filter %d has a NULL plane", first + i);
            b.dst[0][i] = fwd_hi[first + i];  b.dst[1][i] = fwd_lo ? fwd_lo[first + i] : nullptr;
            b.dst[2][i] = dgrad_hi[first + i];  b.dst[3][i] = dgrad_lo ? dgrad_lo[first + i] : nullptr;
        }
        filter_to_planes_multi_kernel<<<b.block_begin[b.count], 256, 0, static_cast<cudaStream_t>(stream)>>>(b);
        if ((rc = check_launch("filter_to_planes_multi_kernel"))) return rc;
    }
    return AVID_OK;
}

int avid_filter_from_tapmajor_multi(const float* const* w_tap, float* const* w_oihw, const int32_t* co, const int32_t* ci, const int32_t* taps,
                                    const int32_t* ci_pad, int32_t count, void* stream) {
    AVID_REQUIRE(w_tap && w_oihw && co && ci && taps && ci_pad && count > 0, "filter_from_tapmajor_multi: bad arguments");
    for (int first = 0; first < count; first += kFilterMax) {
        FilterBatch b;
        int rc = fill_filter_batch(&b, first, count, w_tap, co, ci, taps, ci_pad, "filter_from_tapmajor_multi");
        if (rc) return rc;
        for (int i = 0; i < b.count; ++i) {
            AVID_REQUIRE(w_oihw[first + i] && ci_pad[first + i] >= ci[first + i], "filter_from_tapmajor_multi: filter %d: bad destination / ci_pad", first + i);
            b.dst[0][i] = w_oihw[first + i];
        }
        filter_from_tap_multi_kernel<<<b.block_begin[b.count], 256, 0, static_cast<cudaStream_t>(stream)>>>(b);
        if ((rc = check_launch("filter_from_tap_multi_kernel"))) return rc;
    }
    return AVID_OK;
}

int avid_zero_bytes(void* p, size_t bytes, void* stream) {
    AVID_REQUIRE(p && bytes > 0, "zero_bytes: bad arguments");
    cudaError_t e = cudaMemsetAsync(p, 0, bytes, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) { set_error("zero_bytes: %s", cudaGetErrorString(e)); return AVID_ECUDA; }
    return AVID_OK;
}

int avid_add_inplace(float* a, const float* b, int64_t n, void* stream) {
    AVID_REQUIRE(a && b && n > 0 && n % 4 == 0, "add_inplace: n=%lld must be a positive multiple of 4", (long long)n);
    add_inplace_kernel<<<grid_for(n / 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(a, b, n / 4);
    return check_launch("add_inplace_kernel");
}

int avid_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, int64_t step,
                   float lr, float beta1, float beta2, float eps, float weight_decay, float grad_scale, void* stream) {
    AVID_REQUIRE(param && grad && exp_avg && exp_avg_sq && n > 0 && step > 0, "adam_step: bad arguments");
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    adam_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                                            weight_decay, (float)bc1, (float)sqrt(bc2), grad_scale);
    return check_launch("adam_kernel");
}

}  // extern "C"
