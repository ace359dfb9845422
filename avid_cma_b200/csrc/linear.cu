// Projection heads (models/av_wrapper.py:17-33): y = x W^T + b (+ReLU) and its backward, W in the
// PyTorch [out, in] layout.  The heads are 0.01 % of the step's FLOPs; these are
// plain fp32 kernels sized for rows <= a few hundred, in/out <= 1024.
#include "common.cuh"

namespace avid {

// warp per output feature o: W[o, :] stays in registers, rows stream through
template <int IN4>   // in_f / 128 float4 per lane
__global__ void __launch_bounds__(256) linear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                                         float* __restrict__ y, int rows, int in_f, int out_f, int relu) {
    const int o = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (o >= out_f) return;
    float4 wr[IN4];
#pragma unroll
    for (int j = 0; j < IN4; ++j) wr[j] = __ldg(reinterpret_cast<const float4*>(w + (size_t)o * in_f) + lane + 32 * j);
    const float bias = b ? b[o] : 0.f;
    for (int r = 0; r < rows; ++r) {
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < IN4; ++j) {
            const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (size_t)r * in_f) + lane + 32 * j);
            acc = fmaf(xv.x, wr[j].x, fmaf(xv.y, wr[j].y, fmaf(xv.z, wr[j].z, fmaf(xv.w, wr[j].w, acc))));
        }
        acc = warp_sum(acc) + bias;
        if (relu) acc = fmaxf(acc, 0.f);
        if (lane == 0) y[(size_t)r * out_f + o] = acc;
    }
}

// dy <- dy * (y > 0)
__global__ void relu_mask_kernel(float* dy, const float* __restrict__ y, int64_t n) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n && !(y[i] > 0.f)) dy[i] = 0.f;
}

// dx[r, i] = sum_o dy[r, o] W[o, i]: thread per (r, i), coalesced along i
__global__ void __launch_bounds__(256) linear_dx_kernel(const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx,
                                                        int rows, int in_f, int out_f) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * in_f) return;
    const int r = idx / in_f, i = idx - r * in_f;
    // four independent accumulation chains, 8 loads in flight: the problem is a single partial wave, i.e. latency bound
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    int o = 0;
#pragma unroll 2
    for (; o + 4 <= out_f; o += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] = fmaf(__ldg(dy + (size_t)r * out_f + o + u), __ldg(w + (size_t)(o + u) * in_f + i), acc[u]);
    }
    for (; o < out_f; ++o) acc[0] = fmaf(__ldg(dy + (size_t)r * out_f + o), __ldg(w + (size_t)o * in_f + i), acc[0]);
    dx[idx] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
}

// dW[o, i] = sum_r dy[r, o] x[r, i]; db[o] = sum_r dy[r, o]
__global__ void __launch_bounds__(256) linear_dw_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dw,
                                                        float* __restrict__ db, int rows, int in_f, int out_f) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= out_f * in_f) return;
    const int o = idx / in_f, i = idx - o * in_f;
    float acc = 0.f, accb = 0.f;
    for (int r = 0; r < rows; ++r) {
        const float d = __ldg(dy + (size_t)r * out_f + o);
        acc = fmaf(d, __ldg(x + (size_t)r * in_f + i), acc);
        accb += d;
    }
    dw[idx] = acc;
    if (i == 0 && db) db[o] = accb;
}

}  // namespace avid

using namespace avid;

extern "C" {

int avid_linear_forward(const float* x, const float* w, const float* b, float* y,
                        int32_t rows, int32_t in_f, int32_t out_f, int32_t relu, void* stream) {
    AVID_REQUIRE(x && w && y && rows > 0 && out_f > 0, "linear_forward: bad arguments");
    AVID_REQUIRE(in_f > 0 && in_f % 128 == 0 && in_f <= 1024, "linear_forward: in_features=%d must be a multiple of 128, <= 1024", in_f);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const unsigned grid = (out_f * 32 + 255) / 256;
    switch (in_f / 128) {
        case 1: linear_fwd_kernel<1><<<grid, 256, 0, st>>>(x, w, b, y, rows, in_f, out_f, relu); break;
        case 2: linear_fwd_kernel<2><<<grid, 256, 0, st>>>(x, w, b, y, rows, in_f, out_f, relu); break;
        case 4: linear_fwd_kernel<4><<<grid, 256, 0, st>>>(x, w, b, y, rows, in_f, out_f, relu); break;
        case 8: linear_fwd_kernel<8><<<grid, 256, 0, st>>>(x, w, b, y, rows, in_f, out_f, relu); break;
        default: set_error("linear_forward: in_features=%d unsupported (128, 256, 512 or 1024)", in_f); return AVID_EUNSUPPORTED;
    }
    return check_launch("linear_fwd_kernel");
}

int avid_linear_backward(const float* x, const float* w, const float* y, float* dy, float* dx, float* dw, float* db,
                         int32_t rows, int32_t in_f, int32_t out_f, int32_t relu, void* stream) {
    AVID_REQUIRE(x && w && dy && dw && rows > 0 && in_f > 0 && out_f > 0, "linear_backward: bad arguments");
    AVID_REQUIRE(!relu || y, "linear_backward: relu needs the forward output");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc;
    if (relu) {
        const int64_t n = (int64_t)rows * out_f;
        relu_mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dy, y, n);
        if ((rc = check_launch("relu_mask_kernel"))) return rc;
    }
    if (dx) {
        linear_dx_kernel<<<(rows * in_f + 255) / 256, 256, 0, st>>>(dy, w, dx, rows, in_f, out_f);
        if ((rc = check_launch("linear_dx_kernel"))) return rc;
    }
    linear_dw_kernel<<<(out_f * in_f + 255) / 256, 256, 0, st>>>(dy, x, dw, db, rows, in_f, out_f);
    return check_launch("linear_dw_kernel");
}

}  // extern "C"
