// Train-mode BatchNorm{2d,3d} (+ReLU) on channels-last activations x[rows, c] (include/avid_b200.h).
// PyTorch semantics of the reference's nn.BatchNorm layers (models/network_blocks.py:19-21,36-45,
// models/video.py:21, models/audio.py:23): batch statistics with the biased variance, eps inside the
// square root, running statistics updated with the unbiased variance.  Sums are accumulated in
// fp64 so the statistics do not depend on the (atomic) summation order to fp32 precision.
#include <cuda_bf16.h>
#include <math.h>
#include "common.cuh"

namespace avid {

constexpr int kEwUnroll = 4;      // independent 16-byte loads per thread and stream in the elementwise BatchNorm passes

// The video stem is conv -> BN -> ReLU -> MaxPool3d((1,3,3),(1,2,2),(0,1,1)) (models/video.py:20-23).  In the fused path the
// ReLU output is never written: the forward pools relu(bn(z)) on the fly and records the winning window position; the
// backward kernels obtain "dy" (the gradient at the ReLU output) by gathering the pooled gradient through that argmax.
struct PoolGather {
    const uint8_t* argmax;      // [nt, ho, wo, c] winning position dh*3+dw, or nullptr: dy is a plain tensor
    const float* dyp;           // [nt, ho, wo, c] gradient at the pooled output
    int h, w, ho, wo;
    const float* z;             // [nt, h, w, c] conv output (only read by the pooled reduce when gamma == 0)
};

// gradient at the (unpooled) pixel `row` = (img * h + ih) * w + iw, channels 4*lane_c .. 4*lane_c+3.  32-bit index arithmetic
// (the host checks that the tensor has fewer than 2^31 float4): 64-bit divisions by run-time extents cost more than the loads.
__device__ __forceinline__ float4 pool_gather(const PoolGather& g, uint32_t row, int lane_c, int c4) {
    const uint32_t r = row / (uint32_t)g.w;
    const int iw = (int)(row - r * (uint32_t)g.w);
    const uint32_t img = r / (uint32_t)g.h;
    const int ih = (int)(r - img * (uint32_t)g.h);
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    const int oh0 = ih >> 1, oh1 = min((ih + 1) >> 1, g.ho - 1), ow0 = iw >> 1, ow1 = min((iw + 1) >> 1, g.wo - 1);
    for (int oh = oh0; oh <= oh1; ++oh)
        for (int ow = ow0; ow <= ow1; ++ow) {
            const uint32_t oi = ((img * (uint32_t)g.ho + oh) * (uint32_t)g.wo + ow) * (uint32_t)c4 + lane_c;
            const uint32_t pos = (uint32_t)((ih - (oh * 2 - 1)) * 3 + (iw - (ow * 2 - 1)));
            const uchar4 am = __ldg(reinterpret_cast<const uchar4*>(g.argmax) + oi);
            if (am.x != pos && am.y != pos && am.z != pos && am.w != pos) continue;
            const float4 d = __ldg(reinterpret_cast<const float4*>(g.dyp) + oi);
            if (am.x == pos) o.x += d.x;
            if (am.y == pos) o.y += d.y;
            if (am.z == pos) o.z += d.z;
            if (am.w == pos) o.w += d.w;
        }
    return o;
}

// one thread = one float4 of channels, striding over rows; blockDim.x = 256
// MODE 0: forward statistics of x.  MODE 1: backward sums (sum g, sum g*xhat), g = dy * relu'(bn(x)).  MODE 2: the same sums for a
// pooled layer, computed from the POOLED tensors only (x = pooled output p, dy = its gradient): g is non-zero only at window
// maxima, where y = p, so relu'(y) = [p > 0] and xhat = (p - beta) / gamma -- a pass over 1/4 of the elements, no gather.
template <int MODE>
__global__ void __launch_bounds__(256) bn_reduce_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                        const float* __restrict__ mean, const float* __restrict__ invstd,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        int64_t rows, int c, int64_t rows_per_block, double* __restrict__ out,
                                                        const PoolGather pool) {
    extern __shared__ double s_red[];   // [2][rpi][c]  (rpi = row groups per iteration)
    const int c4 = c >> 2;
    const int rpi = 256 / c4;                    // c4 <= 256 and divides 256 (c in 64..1024, power of two)
    const int lane_c = threadIdx.x % c4, rg = threadIdx.x / c4;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r1 = min(rows, r0 + rows_per_block);
    double a0[4] = {0, 0, 0, 0}, a1[4] = {0, 0, 0, 0};
    float mu[4], is[4], ga[4], be[4];
    if (MODE != 0) {
        *reinterpret_cast<float4*>(mu) = reinterpret_cast<const float4*>(mean)[lane_c];
        *reinterpret_cast<float4*>(is) = reinterpret_cast<const float4*>(invstd)[lane_c];
        *reinterpret_cast<float4*>(ga) = reinterpret_cast<const float4*>(gamma)[lane_c];
        *reinterpret_cast<float4*>(be) = reinterpret_cast<const float4*>(beta)[lane_c];
    }
    for (int64_t r = r0 + rg; r < r1; r += rpi) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + (size_t)r * c) + lane_c);
        const float xv[4] = {v.x, v.y, v.z, v.w};
        if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                a0[j] += (double)xv[j];
                a1[j] += (double)xv[j] * (double)xv[j];
            }
        } else if (MODE == 1) {
            const float4 d = __ldg(reinterpret_cast<const float4*>(dy + (size_t)r * c) + lane_c);
            const float dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float xh = (xv[j] - mu[j]) * is[j];
                const float g = (xh * ga[j] + be[j]) > 0.f ? dv[j] : 0.f;   // relu'(bn(x))
                a0[j] += (double)g;
                a1[j] += (double)g * (double)xh;
            }
        } else {
            const float4 d = __ldg(reinterpret_cast<const float4*>(dy + (size_t)r * c) + lane_c);
            const float dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (!(xv[j] > 0.f)) continue;                               // relu'(y) at the window maximum
                float xh;
                if (ga[j] != 0.f) {
                    xh = (xv[j] - be[j]) / ga[j];
                } else {        // degenerate affine: recover xhat from the conv output at the recorded argmax
                    const uint32_t rr = (uint32_t)r, r2 = rr / (uint32_t)pool.wo, img = r2 / (uint32_t)pool.ho;
                    const int ow = (int)(rr - r2 * (uint32_t)pool.wo), oh = (int)(r2 - img * (uint32_t)pool.ho);
                    const int am = pool.argmax[(size_t)r * c + lane_c * 4 + j];
                    const int ih = oh * 2 - 1 + am / 3, iw = ow * 2 - 1 + am % 3;
                    xh = (pool.z[(((size_t)img * pool.h + ih) * pool.w + iw) * c + lane_c * 4 + j] - mu[j]) * is[j];
                }
                a0[j] += (double)dv[j];
                a1[j] += (double)dv[j] * (double)xh;
            }
        }
    }
    // block reduction over the row groups, then one fp64 atomic per channel per block
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        s_red[(size_t)rg * c + lane_c * 4 + j] = a0[j];
        s_red[(size_t)(rpi + rg) * c + lane_c * 4 + j] = a1[j];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * c; i += 256) {
        const int which = i / c, ch = i - which * c;
        double t = 0.0;
        for (int k = 0; k < rpi; ++k) t += s_red[(size_t)(which * rpi + k) * c + ch];
        atomicAdd(out + i, t);
    }
}

__global__ void bn_finalize_kernel(const double* __restrict__ stats, int64_t rows, int c, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float momentum, float* running_mean,
                                   float* running_var, float* mean, float* invstd, float* scale, float* shift) {
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= c) return;
    const double n = (double)rows;
    const double m = stats[ch] / n;
    double var = stats[c + ch] / n - m * m;
    if (var < 0.0) var = 0.0;
    const float is = (float)(1.0 / sqrt(var + (double)eps));
    mean[ch] = (float)m;
    invstd[ch] = is;
    const float sc = gamma[ch] * is;
    scale[ch] = sc;
    shift[ch] = beta[ch] - (float)m * sc;
    if (running_mean) {
        const double unbiased = rows > 1 ? var * n / (n - 1.0) : var;
        running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * (float)m;
        running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * (float)unbiased;
    }
}

// hi = bf16(v), lo = bf16(v - hi): the operand planes of the tcgen05 convolutions
__device__ __forceinline__ void store_planes(__nv_bfloat16* hi, __nv_bfloat16* lo, int64_t i, const float (&f)[4]) {
    __nv_bfloat16 h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        h[j] = __float2bfloat16_rn(f[j]);
        l[j] = __float2bfloat16_rn(f[j] - __bfloat162float(h[j]));
    }
    reinterpret_cast<uint2*>(hi)[i] = *reinterpret_cast<uint2*>(h);
    if (lo) reinterpret_cast<uint2*>(lo)[i] = *reinterpret_cast<uint2*>(l);
}

// Elementwise kernels: one float4 of channels per thread per iteration.  c4 (a power of two <= 256) divides the grid stride
// (gridDim.x * 256), so a thread sees the SAME four channels in every iteration: their constants live in registers.
__global__ void __launch_bounds__(256) bn_relu_forward_kernel(const float* __restrict__ x, const float* __restrict__ scale,
                                                              const float* __restrict__ shift, float* __restrict__ y,
                                                              __nv_bfloat16* __restrict__ y_hi, __nv_bfloat16* __restrict__ y_lo,
                                                              int64_t n4, int c4) {
    const int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    const int cc = (int)(i0 % c4);
    const float4 sc = __ldg(reinterpret_cast<const float4*>(scale) + cc);
    const float4 sh = __ldg(reinterpret_cast<const float4*>(shift) + cc);
    // kEwUnroll independent 16-byte loads per thread and input stream before the first use: ~100 KB in flight per SM (the
    // bandwidth-delay product of HBM3e is ~44 KB per SM; with 2 loads these passes ran at 4.6-5.1 TB/s of the 6.5 TB/s copy peak)
    for (int64_t i = i0; i < n4; i += kEwUnroll * stride) {
        float4 v[kEwUnroll];
#pragma unroll
        for (int u = 0; u < kEwUnroll; ++u) {
            const int64_t k = i + u * stride;
            v[u] = k < n4 ? __ldg(reinterpret_cast<const float4*>(x) + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < kEwUnroll; ++u) {
            const int64_t k = i + u * stride;
            if (k >= n4) break;
            float4 o;
            o.x = fmaxf(fmaf(v[u].x, sc.x, sh.x), 0.f);
            o.y = fmaxf(fmaf(v[u].y, sc.y, sh.y), 0.f);
            o.z = fmaxf(fmaf(v[u].z, sc.z, sh.z), 0.f);
            o.w = fmaxf(fmaf(v[u].w, sc.w, sh.w), 0.f);
            if (y) reinterpret_cast<float4*>(y)[k] = o;
            if (y_hi) {
                const float f[4] = {o.x, o.y, o.z, o.w};
                store_planes(y_hi, y_lo, k, f);
            }
        }
    }
}

// pooled[nt, ho, wo, c] = max over the 3x3 / stride 2 / pad 1 window of relu(z * scale + shift); one thread per float4 of
// pooled channels (its four channels are the same in every grid-stride iteration)
__global__ void __launch_bounds__(256) bn_relu_maxpool_forward_kernel(const float* __restrict__ z, const float* __restrict__ scale,
                                                                      const float* __restrict__ shift, float* __restrict__ p,
                                                                      __nv_bfloat16* __restrict__ p_hi, __nv_bfloat16* __restrict__ p_lo,
                                                                      uint8_t* __restrict__ argmax, int nt, int h, int w, int c4, int ho, int wo) {
    const int64_t total = (int64_t)nt * ho * wo * c4;
    const int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int cc = (int)(i0 % c4);
    const float4 sc4 = __ldg(reinterpret_cast<const float4*>(scale) + cc);
    const float4 sh4 = __ldg(reinterpret_cast<const float4*>(shift) + cc);
    const float sc[4] = {sc4.x, sc4.y, sc4.z, sc4.w}, sh[4] = {sh4.x, sh4.y, sh4.z, sh4.w};
    for (int64_t i = i0; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t r = (uint32_t)i / (uint32_t)c4;                 // 32-bit index arithmetic: the host checks the tensor size
        const uint32_t r1 = r / (uint32_t)wo;
        const int ow = (int)(r - r1 * (uint32_t)wo);
        const uint32_t img = r1 / (uint32_t)ho;
        const int oh = (int)(r1 - img * (uint32_t)ho);
        // all nine window loads are issued before the first use (clamped addresses + validity flags instead of skipped iterations: with
        // `continue` in the tap loops the loads were predicated one by one behind the running maximum -- ncu: 16 long-scoreboard stalls
        // per issued instruction, 48 % of the DRAM rate)
        float4 v9[9];
        bool ok9[9];
#pragma unroll
        for (int dh = 0; dh < 3; ++dh) {
            const int ih = oh * 2 - 1 + dh;
            const int ihc = min(max(ih, 0), h - 1);
#pragma unroll
            for (int dw = 0; dw < 3; ++dw) {
                const int iw = ow * 2 - 1 + dw;
                const int iwc = min(max(iw, 0), w - 1);
                ok9[dh * 3 + dw] = ih >= 0 && ih < h && iw >= 0 && iw < w;
                v9[dh * 3 + dw] = __ldg(reinterpret_cast<const float4*>(z) + ((img * (uint32_t)h + ihc) * (uint32_t)w + iwc) * (uint32_t)c4 + cc);
            }
        }
        float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        uint8_t am[4] = {0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            if (!ok9[k]) continue;
            const float v[4] = {v9[k].x, v9[k].y, v9[k].z, v9[k].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float y = fmaxf(fmaf(v[j], sc[j], sh[j]), 0.f);
                if (y > m[j]) { m[j] = y; am[j] = (uint8_t)k; }      // strictly greater keeps the first maximum
            }
        }
        if (p) reinterpret_cast<float4*>(p)[i] = make_float4(m[0], m[1], m[2], m[3]);
        if (p_hi) store_planes(p_hi, p_lo, i, m);
        if (argmax) reinterpret_cast<uchar4*>(argmax)[i] = make_uchar4(am[0], am[1], am[2], am[3]);
    }
}

__global__ void __launch_bounds__(256, 3) bn_relu_backward_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                     const float* __restrict__ mean, const float* __restrict__ invstd,
                                                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                     const double* __restrict__ sums, int64_t rows, int c,
                                                                     float* __restrict__ dx, __nv_bfloat16* __restrict__ dx_hi,
                                                                     __nv_bfloat16* __restrict__ dx_lo, float* dgamma, float* dbeta,
                                                                     const PoolGather pool) {
    const int c4 = c >> 2;
    const int64_t n4 = rows * c4;
    const float inv_n = 1.0f / (float)rows;
    if (blockIdx.x == 0)
        for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
            if (dbeta) dbeta[ch] = (float)sums[ch];
            if (dgamma) dgamma[ch] = (float)sums[c + ch];
        }
    const int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    const int cc = (int)(i0 % c4);
    // dx = k * (g - sg - xh * sgx),  xh = (x - mu) * is,  g = dy * [xh * ga + be > 0]
    float mu[4], is[4], ga[4], be[4], kk[4], sg[4], sgx[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int ch = cc * 4 + j;
        mu[j] = mean[ch];  is[j] = invstd[ch];  ga[j] = gamma[ch];  be[j] = beta[ch];
        kk[j] = ga[j] * is[j];
        sg[j] = (float)sums[ch] * inv_n;
        sgx[j] = (float)sums[c + ch] * inv_n;
    }
    for (int64_t i = i0; i < n4; i += kEwUnroll * stride) {
        float4 xv4[kEwUnroll], dv4[kEwUnroll];
#pragma unroll
        for (int u = 0; u < kEwUnroll; ++u) {
            const int64_t k = i + u * stride;
            const bool in = k < n4;
            xv4[u] = in ? __ldg(reinterpret_cast<const float4*>(x) + k) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (pool.argmax) dv4[u] = in ? pool_gather(pool, (uint32_t)k / (uint32_t)c4, cc, c4) : make_float4(0.f, 0.f, 0.f, 0.f);
            else dv4[u] = in ? __ldg(reinterpret_cast<const float4*>(dy) + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < kEwUnroll; ++u) {
            const int64_t k = i + u * stride;
            if (k >= n4) break;
            const float xv[4] = {xv4[u].x, xv4[u].y, xv4[u].z, xv4[u].w}, dv[4] = {dv4[u].x, dv4[u].y, dv4[u].z, dv4[u].w};
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float xh = (xv[j] - mu[j]) * is[j];
                const float g = (xh * ga[j] + be[j]) > 0.f ? dv[j] : 0.f;
                o[j] = kk[j] * (g - sg[j] - xh * sgx[j]);
            }
            if (dx) reinterpret_cast<float4*>(dx)[k] = make_float4(o[0], o[1], o[2], o[3]);
            if (dx_hi) store_planes(dx_hi, dx_lo, k, o);
        }
    }
}

// Backward apply of the pooled stem layer on 2 x 2 blocks of unpooled pixels: the block (rows 2a, 2a+1; columns 2b, 2b+1) is
// covered by the four pooling windows (a..a+1, b..b+1) only, so one thread loads 4 argmax entries + 4 pooled gradients (all
// independent, no data-dependent branches) for 4 output pixels instead of walking up to 4 windows per pixel.  Window (oh, ow)
// holds pixel (ih, iw) at position (ih - 2 oh + 1) * 3 + (iw - 2 ow + 1).
__global__ void __launch_bounds__(256, 3) bn_relu_pool_backward_apply_kernel(const float* __restrict__ z, const uint8_t* __restrict__ argmax,
                                                                          const float* __restrict__ dyp, const float* __restrict__ mean,
                                                                          const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                                          const float* __restrict__ beta, const double* __restrict__ sums,
                                                                          int nt, int h, int w, int c, int ho, int wo, float* __restrict__ dx,
                                                                          __nv_bfloat16* __restrict__ dx_hi, __nv_bfloat16* __restrict__ dx_lo,
                                                                          float* dgamma, float* dbeta) {
    const int c4 = c >> 2;
    const uint32_t hb = (uint32_t)(h + 1) >> 1, wb = (uint32_t)(w + 1) >> 1;
    const uint32_t total = (uint32_t)nt * hb * wb * (uint32_t)c4;
    const float inv_n = 1.0f / ((float)nt * (float)h * (float)w);
    if (blockIdx.x == 0)
        for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
            if (dbeta) dbeta[ch] = (float)sums[ch];
            if (dgamma) dgamma[ch] = (float)sums[c + ch];
        }
    const uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;      // stride is a multiple of c4
    const int cc = (int)(i0 % (uint32_t)c4);
    float mu[4], is[4], ga[4], be[4], kk[4], sg[4], sgx[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int ch = cc * 4 + j;
        mu[j] = mean[ch];  is[j] = invstd[ch];  ga[j] = gamma[ch];  be[j] = beta[ch];
        kk[j] = ga[j] * is[j];
        sg[j] = (float)sums[ch] * inv_n;
        sgx[j] = (float)sums[c + ch] * inv_n;
    }
    for (uint32_t i = i0; i < total; i += stride) {
        uint32_t r = i / (uint32_t)c4;
        const uint32_t b = r % wb;  r /= wb;
        const uint32_t a = r % hb;
        const uint32_t img = r / hb;
        // the four windows: [dh][dw] = (a + dh, b + dw)
        uchar4 am[2][2];
        float4 d[2][2];
#pragma unroll
        for (int dh = 0; dh < 2; ++dh)
#pragma unroll
            for (int dw = 0; dw < 2; ++dw) {
                const uint32_t oh = a + dh, ow = b + dw;
                am[dh][dw] = make_uchar4(255, 255, 255, 255);
                d[dh][dw] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (oh < (uint32_t)ho && ow < (uint32_t)wo) {
                    const uint32_t oi = ((img * (uint32_t)ho + oh) * (uint32_t)wo + ow) * (uint32_t)c4 + cc;
                    am[dh][dw] = __ldg(reinterpret_cast<const uchar4*>(argmax) + oi);
                    d[dh][dw] = __ldg(reinterpret_cast<const float4*>(dyp) + oi);
                }
            }
        float4 zv[2][2];
        uint32_t pi[2][2];
        bool ok[2][2];
#pragma unroll
        for (int py = 0; py < 2; ++py)
#pragma unroll
            for (int px = 0; px < 2; ++px) {
                const uint32_t ih = 2 * a + py, iw = 2 * b + px;
                ok[py][px] = ih < (uint32_t)h && iw < (uint32_t)w;
                pi[py][px] = ((img * (uint32_t)h + ih) * (uint32_t)w + iw) * (uint32_t)c4 + cc;
                zv[py][px] = ok[py][px] ? __ldg(reinterpret_cast<const float4*>(z) + pi[py][px]) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
        for (int py = 0; py < 2; ++py)
#pragma unroll
            for (int px = 0; px < 2; ++px) {
                if (!ok[py][px]) continue;
                // gradient at the ReLU output of this pixel: pooled gradients of the windows whose argmax is this pixel
                float g[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int dh = 0; dh <= py; ++dh)
#pragma unroll
                    for (int dw = 0; dw <= px; ++dw) {
                        const unsigned pos = (unsigned)((py - 2 * dh + 1) * 3 + (px - 2 * dw + 1));
                        const uchar4 m = am[dh][dw];
                        const float4 dd = d[dh][dw];
                        if (m.x == pos) g[0] += dd.x;
                        if (m.y == pos) g[1] += dd.y;
                        if (m.z == pos) g[2] += dd.z;
                        if (m.w == pos) g[3] += dd.w;
                    }
                const float xv[4] = {zv[py][px].x, zv[py][px].y, zv[py][px].z, zv[py][px].w};
                float o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float xh = (xv[j] - mu[j]) * is[j];
                    const float gg = (xh * ga[j] + be[j]) > 0.f ? g[j] : 0.f;
                    o[j] = kk[j] * (gg - sg[j] - xh * sgx[j]);
                }
                if (dx) reinterpret_cast<float4*>(dx)[pi[py][px]] = make_float4(o[0], o[1], o[2], o[3]);
                if (dx_hi) store_planes(dx_hi, dx_lo, (int64_t)pi[py][px], o);
            }
    }
}

static int check_bn_shape(int64_t rows, int c) {
    AVID_REQUIRE(rows > 0, "bn: rows must be positive");
    AVID_REQUIRE(c >= 4 && c <= 1024 && (c & (c - 1)) == 0, "bn: c=%d must be a power of two in [4,1024]", c);
    return AVID_OK;
}

static unsigned ew_grid(int64_t n4) {
    int64_t b = (n4 + 255) / 256;
    const int64_t cap = 16 * kNumSMs;
    return (unsigned)(b < cap ? (b < 1 ? 1 : b) : cap);
}

template <int MODE>
static int launch_reduce(const float* x, const float* dy, const float* mean, const float* invstd, const float* gamma,
                         const float* beta, int64_t rows, int c, double* out, cudaStream_t st, const PoolGather pool = PoolGather{}) {
    const int c4 = c >> 2, rpi = 256 / c4;
    int64_t blocks = 8 * kNumSMs;
    int64_t rpb = (rows + blocks - 1) / blocks;
    const int64_t min_rpb = 4 * rpi;
    if (rpb < min_rpb) rpb = min_rpb;
    blocks = (rows + rpb - 1) / rpb;
    const size_t smem = sizeof(double) * 2 * rpi * c;   // = 2 * 256 * 4 * 8 = 16 KB
    bn_reduce_kernel<MODE><<<(unsigned)blocks, 256, smem, st>>>(x, dy, mean, invstd, gamma, beta, rows, c, rpb, out, pool);
    return check_launch(MODE == 0 ? "bn_reduce_kernel<fwd>" : (MODE == 1 ? "bn_reduce_kernel<bwd>" : "bn_reduce_kernel<pooled bwd>"));
}

}  // namespace avid

using namespace avid;

extern "C" {

int avid_bn_stats(const float* x, int64_t rows, int32_t c, double* stats, void* stream) {
    int rc = check_bn_shape(rows, c);
    if (rc) return rc;
    AVID_REQUIRE(x && stats, "bn_stats: NULL pointer");
    return launch_reduce<0>(x, nullptr, nullptr, nullptr, nullptr, nullptr, rows, c, stats, static_cast<cudaStream_t>(stream));
}

int avid_bn_finalize(const double* stats, int64_t rows, int32_t c, const float* gamma, const float* beta,
                     float eps, float momentum, float* running_mean, float* running_var,
                     float* mean, float* invstd, float* scale, float* shift, void* stream) {
    int rc = check_bn_shape(rows, c);
    if (rc) return rc;
    AVID_REQUIRE(stats && gamma && beta && mean && invstd && scale && shift, "bn_finalize: NULL pointer");
    AVID_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "bn_finalize: running_mean / running_var must both be given or both NULL");
    bn_finalize_kernel<<<(c + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(stats, rows, c, gamma, beta, eps, momentum,
                                                                                       running_mean, running_var, mean, invstd, scale, shift);
    return check_launch("bn_finalize_kernel");
}

int avid_bn_relu_forward_ex(const float* x, const float* scale, const float* shift, float* y, void* y_hi, void* y_lo,
                            int64_t rows, int32_t c, void* stream) {
    int rc = check_bn_shape(rows, c);
    if (rc) return rc;
    AVID_REQUIRE(x && scale && shift && (y || y_hi), "bn_relu_forward: NULL pointer");
    AVID_REQUIRE(y_hi || !y_lo, "bn_relu_forward: a lo plane needs the hi plane");
    const int64_t n4 = rows * (c >> 2);
    bn_relu_forward_kernel<<<ew_grid(n4), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, scale, shift, y, static_cast<__nv_bfloat16*>(y_hi),
                                                                                      static_cast<__nv_bfloat16*>(y_lo), n4, c >> 2);
    return check_launch("bn_relu_forward_kernel");
}

int avid_bn_relu_forward(const float* x, const float* scale, const float* shift, float* y, int64_t rows, int32_t c, void* stream) {
    AVID_REQUIRE(y, "bn_relu_forward: NULL pointer");
    return avid_bn_relu_forward_ex(x, scale, shift, y, nullptr, nullptr, rows, c, stream);
}

int avid_bn_relu_backward_reduce(const float* x, const float* dy, const float* mean, const float* invstd,
                                 const float* gamma, const float* beta, int64_t rows, int32_t c, double* sums, void* stream) {
    int rc = check_bn_shape(rows, c);
    if (rc) return rc;
    AVID_REQUIRE(x && dy && mean && invstd && gamma && beta && sums, "bn_relu_backward_reduce: NULL pointer");
    return launch_reduce<1>(x, dy, mean, invstd, gamma, beta, rows, c, sums, static_cast<cudaStream_t>(stream));
}

int avid_bn_relu_backward_apply_ex(const float* x, const float* dy, const float* mean, const float* invstd,
                                   const float* gamma, const float* beta, const double* sums,
                                   int64_t rows, int32_t c, float* dx, void* dx_hi, void* dx_lo, float* dgamma, float* dbeta, void* stream) {
    int rc = check_bn_shape(rows, c);
    if (rc) return rc;
    AVID_REQUIRE(x && dy && mean && invstd && gamma && beta && sums && (dx || dx_hi), "bn_relu_backward_apply: NULL pointer");
    AVID_REQUIRE(dx_hi || !dx_lo, "bn_relu_backward_apply: a lo plane needs the hi plane");
    bn_relu_backward_apply_kernel<<<ew_grid(rows * (c >> 2)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, dy, mean, invstd, gamma, beta, sums, rows, c, dx, static_cast<__nv_bfloat16*>(dx_hi), static_cast<__nv_bfloat16*>(dx_lo), dgamma, dbeta,
        PoolGather{});
    return check_launch("bn_relu_backward_apply_kernel");
}

static int check_pool_shape(int32_t nt, int32_t h, int32_t w, int32_t c, int32_t ho, int32_t wo) {
    AVID_REQUIRE(nt > 0 && h > 0 && w > 0, "bn_relu_maxpool: bad extents");
    AVID_REQUIRE(ho == (h + 2 - 3) / 2 + 1 && wo == (w + 2 - 3) / 2 + 1, "bn_relu_maxpool: pooled extent mismatch");
    AVID_REQUIRE((int64_t)nt * h * w * (c / 4) < ((int64_t)1 << 31), "bn_relu_maxpool: tensor too large for 32-bit indexing");
    return check_bn_shape((int64_t)nt * h * w, c);
}

int avid_bn_relu_maxpool_forward(const float* z, const float* scale, const float* shift, float* p, void* p_hi, void* p_lo, uint8_t* argmax,
                                 int32_t nt, int32_t h, int32_t w, int32_t c, int32_t ho, int32_t wo, void* stream) {
    int rc = check_pool_shape(nt, h, w, c, ho, wo);
    if (rc) return rc;
    AVID_REQUIRE(z && scale && shift && (p || p_hi), "bn_relu_maxpool_forward: NULL pointer");
    AVID_REQUIRE(p_hi || !p_lo, "bn_relu_maxpool_forward: a lo plane needs the hi plane");
    bn_relu_maxpool_forward_kernel<<<ew_grid((int64_t)nt * ho * wo * (c >> 2)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        z, scale, shift, p, static_cast<__nv_bfloat16*>(p_hi), static_cast<__nv_bfloat16*>(p_lo), argmax, nt, h, w, c >> 2, ho, wo);
    return check_launch("bn_relu_maxpool_forward_kernel");
}

int avid_bn_relu_maxpool_backward_reduce(const float* z, const float* pooled, const uint8_t* argmax, const float* dyp, const float* mean,
                                         const float* invstd, const float* gamma, const float* beta, int32_t nt, int32_t h, int32_t w, int32_t c,
                                         int32_t ho, int32_t wo, double* sums, void* stream) {
    int rc = check_pool_shape(nt, h, w, c, ho, wo);
    if (rc) return rc;
    AVID_REQUIRE(z && pooled && argmax && dyp && mean && invstd && gamma && beta && sums, "bn_relu_maxpool_backward_reduce: NULL pointer");
    return launch_reduce<2>(pooled, dyp, mean, invstd, gamma, beta, (int64_t)nt * ho * wo, c, sums, static_cast<cudaStream_t>(stream),
                            PoolGather{argmax, dyp, h, w, ho, wo, z});
}

int avid_bn_relu_maxpool_backward_apply(const float* z, const uint8_t* argmax, const float* dyp, const float* mean, const float* invstd,
                                        const float* gamma, const float* beta, const double* sums, int32_t nt, int32_t h, int32_t w, int32_t c,
                                        int32_t ho, int32_t wo, float* dz, void* dz_hi, void* dz_lo, float* dgamma, float* dbeta, void* stream) {
    int rc = check_pool_shape(nt, h, w, c, ho, wo);
    if (rc) return rc;
    AVID_REQUIRE(z && argmax && dyp && mean && invstd && gamma && beta && sums && (dz || dz_hi), "bn_relu_maxpool_backward_apply: NULL pointer");
    AVID_REQUIRE(dz_hi || !dz_lo, "bn_relu_maxpool_backward_apply: a lo plane needs the hi plane");
    const int64_t blocks4 = (int64_t)nt * ((h + 1) / 2) * ((w + 1) / 2) * (c >> 2);
    bn_relu_pool_backward_apply_kernel<<<ew_grid(blocks4), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        z, argmax, dyp, mean, invstd, gamma, beta, sums, nt, h, w, c, ho, wo, dz, static_cast<__nv_bfloat16*>(dz_hi),
        static_cast<__nv_bfloat16*>(dz_lo), dgamma, dbeta);
    return check_launch("bn_relu_pool_backward_apply_kernel");
}

int avid_bn_relu_backward_apply(const float* x, const float* dy, const float* mean, const float* invstd,
                                const float* gamma, const float* beta, const double* sums,
                                int64_t rows, int32_t c, float* dx, float* dgamma, float* dbeta, void* stream) {
    AVID_REQUIRE(dx, "bn_relu_backward_apply: NULL pointer");
    return avid_bn_relu_backward_apply_ex(x, dy, mean, invstd, gamma, beta, sums, rows, c, dx, nullptr, nullptr, dgamma, dbeta, stream);
}

}  // extern "C"
