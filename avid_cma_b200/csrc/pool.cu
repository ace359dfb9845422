// Pooling layers of the towers on channels-last activations (include/avid_b200.h):
// nn.MaxPool3d((1,3,3),(1,2,2),(0,1,1)) of the video stem (models/video.py:23) and the global
// nn.AdaptiveMaxPool{2d,3d}(1) (models/video.py:41, models/audio.py:31).  Bandwidth-bound: one
// float4 of channels per thread, every global access a coalesced run of channels.
#include <math.h>
#include "common.cuh"

namespace avid {

__device__ __forceinline__ float4 max4(const float4& a, const float4& b) {
    return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
}

// forward also records WHICH window position (dh * 3 + dw, first maximum in scan order like ATen) won, one byte per output
__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, uint8_t* __restrict__ argmax,
                                                          int nt, int h, int w, int c4, int ho, int wo) {
    const int64_t total = (int64_t)nt * ho * wo * c4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t r0 = (uint32_t)i / (uint32_t)c4;          // 32-bit index arithmetic (size checked by the host)
        const int cc = (int)((uint32_t)i - r0 * (uint32_t)c4);
        const uint32_t r1 = r0 / (uint32_t)wo;
        const int ow = (int)(r0 - r1 * (uint32_t)wo);
        const uint32_t img = r1 / (uint32_t)ho;
        const int oh = (int)(r1 - img * (uint32_t)ho);
        float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        uint8_t am[4] = {0, 0, 0, 0};
#pragma unroll
        for (int dh = 0; dh < 3; ++dh) {
            const int ih = oh * 2 - 1 + dh;
            if (ih < 0 || ih >= h) continue;
#pragma unroll
            for (int dw = 0; dw < 3; ++dw) {
                const int iw = ow * 2 - 1 + dw;
                if (iw < 0 || iw >= w) continue;
                const float4 v4 = __ldg(reinterpret_cast<const float4*>(x) + ((img * h + ih) * w + iw) * c4 + cc);
                const float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (v[j] > m[j] || (v[j] != v[j] && m[j] == m[j])) {      // strictly greater keeps the first maximum; NaN propagates
                        m[j] = v[j];
                        am[j] = (uint8_t)(dh * 3 + dw);
                    }
            }
        }
        reinterpret_cast<float4*>(y)[i] = make_float4(m[0], m[1], m[2], m[3]);
        if (argmax) reinterpret_cast<uchar4*>(argmax)[i] = make_uchar4(am[0], am[1], am[2], am[3]);
    }
}

// Gather formulation of the backward: one thread per float4 of INPUT channels looks at the (at most 2 x 2) windows that
// contain its pixel and takes their dy where the recorded argmax is this pixel -- every dx element is written exactly once
// (no zero fill, no atomics, x is not read).
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const uint8_t* __restrict__ argmax, const float* __restrict__ dy,
                                                          float* __restrict__ dx, int nt, int h, int w, int c4, int ho, int wo) {
    const int64_t total = (int64_t)nt * h * w * c4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t r0 = (uint32_t)i / (uint32_t)c4;
        const int cc = (int)((uint32_t)i - r0 * (uint32_t)c4);
        const uint32_t r1 = r0 / (uint32_t)w;
        const int iw = (int)(r0 - r1 * (uint32_t)w);
        const uint32_t img = r1 / (uint32_t)h;
        const int ih = (int)(r1 - img * (uint32_t)h);
        float o[4] = {0.f, 0.f, 0.f, 0.f};
        for (int oh = ih >> 1; oh <= ((ih + 1) >> 1) && oh < ho; ++oh)
            for (int ow = iw >> 1; ow <= ((iw + 1) >> 1) && ow < wo; ++ow) {
                const int64_t oi = ((img * ho + oh) * wo + ow) * c4 + cc;
                const uint8_t pos = (uint8_t)((ih - (oh * 2 - 1)) * 3 + (iw - (ow * 2 - 1)));
                const uchar4 am = __ldg(reinterpret_cast<const uchar4*>(argmax) + oi);
                if (am.x != pos && am.y != pos && am.z != pos && am.w != pos) continue;
                const float4 g4 = __ldg(reinterpret_cast<const float4*>(dy) + oi);
                if (am.x == pos) o[0] += g4.x;
                if (am.y == pos) o[1] += g4.y;
                if (am.z == pos) o[2] += g4.z;
                if (am.w == pos) o[3] += g4.w;
            }
        reinterpret_cast<float4*>(dx)[i] = make_float4(o[0], o[1], o[2], o[3]);
    }
}

__global__ void __launch_bounds__(128) global_maxpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int* __restrict__ argmax,
                                                                 int n, int64_t thw, int c) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * c) return;
    const int img = i / c, ch = i - img * c;
    const float* p = x + (size_t)img * thw * c + ch;
    float m = -INFINITY;
    int am = 0;
    for (int64_t s = 0; s < thw; ++s) {
        const float v = __ldg(p + s * c);
        if (v > m) { m = v; am = (int)s; }
    }
    y[i] = m;
    if (argmax) argmax[i] = am;
}

__global__ void __launch_bounds__(128) global_maxpool_bwd_kernel(const float* __restrict__ dy, const int* __restrict__ argmax,
                                                                 float* __restrict__ dx, int n, int64_t thw, int c) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * c) return;
    const int img = i / c, ch = i - img * c;
    dx[((size_t)img * thw + argmax[i]) * c + ch] = dy[i];
}

static unsigned grid_for(int64_t n) {
    int64_t b = (n + 255) / 256;
    const int64_t cap = 16 * kNumSMs;
    return (unsigned)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace avid

using namespace avid;

extern "C" {

int avid_maxpool_1x3x3_forward(const float* x, float* y, uint8_t* argmax, int32_t nt, int32_t h, int32_t w, int32_t c, int32_t ho, int32_t wo,
                               void* stream) {
    AVID_REQUIRE(x && y && nt > 0 && h > 0 && w > 0 && c > 0 && c % 4 == 0, "maxpool_forward: bad arguments");
    AVID_REQUIRE((int64_t)nt * h * w * (c / 4) < ((int64_t)1 << 31), "maxpool_forward: tensor too large for 32-bit indexing");
    AVID_REQUIRE(ho == (h + 2 - 3) / 2 + 1 && wo == (w + 2 - 3) / 2 + 1, "maxpool_forward: output extent mismatch");
    maxpool_fwd_kernel<<<grid_for((int64_t)nt * ho * wo * (c / 4)), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, argmax, nt, h, w, c / 4, ho, wo);
    return check_launch("maxpool_fwd_kernel");
}

int avid_maxpool_1x3x3_backward(const uint8_t* argmax, const float* dy, float* dx,
                                int32_t nt, int32_t h, int32_t w, int32_t c, int32_t ho, int32_t wo, void* stream) {
    AVID_REQUIRE(argmax && dy && dx && nt > 0 && h > 0 && w > 0 && c > 0 && c % 4 == 0, "maxpool_backward: bad arguments");
    AVID_REQUIRE((int64_t)nt * h * w * (c / 4) < ((int64_t)1 << 31), "maxpool_backward: tensor too large for 32-bit indexing");
    AVID_REQUIRE(ho == (h + 2 - 3) / 2 + 1 && wo == (w + 2 - 3) / 2 + 1, "maxpool_backward: output extent mismatch");
    maxpool_bwd_kernel<<<grid_for((int64_t)nt * h * w * (c / 4)), 256, 0, static_cast<cudaStream_t>(stream)>>>(argmax, dy, dx, nt, h, w, c / 4, ho, wo);
    return check_launch("maxpool_bwd_kernel");
}

int avid_global_maxpool_forward(const float* x, float* y, int32_t* argmax, int32_t n, int64_t thw, int32_t c, void* stream) {
    AVID_REQUIRE(x && y && n > 0 && thw > 0 && c > 0, "global_maxpool_forward: bad arguments");
    global_maxpool_fwd_kernel<<<(n * c + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(x, y, argmax, n, thw, c);
    return check_launch("global_maxpool_fwd_kernel");
}

int avid_global_maxpool_backward(const float* dy, const int32_t* argmax, float* dx, int32_t n, int64_t thw, int32_t c, void* stream) {
    AVID_REQUIRE(dy && argmax && dx && n > 0 && thw > 0 && c > 0, "global_maxpool_backward: bad arguments");
    global_maxpool_bwd_kernel<<<(n * c + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(dy, argmax, dx, n, thw, c);
    return check_launch("global_maxpool_bwd_kernel");
}

}  // extern "C"
