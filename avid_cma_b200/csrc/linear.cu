// Projection heads (models/av_wrapper.py:17-33): y = x W^T + b (+ReLU) and its backward, W in the
// PyTorch [out, in] layout.  The heads are 0.01 % of the step's FLOPs; these are
// plain fp32 kernels sized for rows <= a few hundred, in/out <= 1024.
#include <stdint.h>
#include "common.cuh"

namespace avid {

// One shared-memory tiled fp32 kernel serves the three products of a Linear layer through strides:
//     C[m, n] = sum_k A[m * a_m + k * a_k] * B[n * b_n + k * b_k]   (+ bias[n], ReLU)
//   forward  y[r, o]  = sum_i x[r, i]  W[o, i]      A = x  (in_f, 1)    B = W (in_f, 1)
//   dx       dx[r, i] = sum_o dy[r, o] W[o, i]      A = dy (out_f, 1)   B = W (1, in_f)
//   dW       dW[o, i] = sum_r dy[r, o] x[r, i]      A = dy (1, out_f)   B = x (1, in_f)
// 32 x 32 output tile, k-blocks of 32, 256 threads with 2 x 2 outputs each.  A tile is read with the threads running along
// whichever index is contiguous in memory, so all three cases load coalesced.  (The earlier warp-per-output kernels spent
// 45-80 us per layer on shuffles and dependent loads: 0.75 ms per step for 0.003 % of its FLOPs.)
constexpr int kLT = 32;

// one float4 of a 32 x 32 tile per thread, along whichever index is contiguous in memory (s_k == 1: along k, else along r);
// tiles that cross the matrix edge fall back to guarded scalar loads
struct TileFrag {
    float v[4];
};
__device__ __forceinline__ TileFrag fetch_frag(const float* __restrict__ src, int rows_total, int k_total, int r0, int k0, long s_r, long s_k, int tid) {
    const int c = (tid & 7) * 4, o = tid >> 3;
    const int r = s_k == 1 ? o : c, k = s_k == 1 ? c : o;
    TileFrag f;
    const bool full = s_k == 1 ? (r0 + r < rows_total && k0 + k + 3 < k_total) : (r0 + r + 3 < rows_total && k0 + k < k_total);
    const float* ptr = src + (long)(r0 + r) * s_r + (long)(k0 + k) * s_k;
    if (full && (reinterpret_cast<uintptr_t>(ptr) & 15) == 0) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(ptr));
        f.v[0] = q.x; f.v[1] = q.y; f.v[2] = q.z; f.v[3] = q.w;
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int rr = s_k == 1 ? r : r + j, kk = s_k == 1 ? k + j : k;
            f.v[j] = (r0 + rr < rows_total && k0 + kk < k_total) ? __ldg(src + (long)(r0 + rr) * s_r + (long)(k0 + kk) * s_k) : 0.f;
        }
    }
    return f;
}
__device__ __forceinline__ void store_frag(float (*dst)[kLT + 1], const TileFrag& f, long s_k, int tid) {
    const int c = (tid & 7) * 4, o = tid >> 3;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (s_k == 1) dst[c + j][o] = f.v[j];      // dst[k][r]
        else dst[o][c + j] = f.v[j];
    }
}

__global__ void __launch_bounds__(256) linear_gemm_kernel(const float* __restrict__ A, const float* __restrict__ B, const float* __restrict__ bias,
                                                          float* __restrict__ Cmat, int M, int N, int K, long a_m, long a_k, long b_n, long b_k,
                                                          int relu, float* __restrict__ colsum_a) {
    __shared__ float sa[kLT][kLT + 1], sb[kLT][kLT + 1];       // [k][m], [k][n]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * kLT, n0 = blockIdx.x * kLT;
    float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    float asum[2] = {0.f, 0.f};
    // the next k-block is fetched into registers while the current one is multiplied: the layers are tiny (one partial wave), so
    // the kernel is a latency chain and every un-overlapped global load shows
    TileFrag fa = fetch_frag(A, M, K, m0, 0, a_m, a_k, tid), fb = fetch_frag(B, N, K, n0, 0, b_n, b_k, tid);
    for (int k0 = 0; k0 < K; k0 += kLT) {
        store_frag(sa, fa, a_k, tid);
        store_frag(sb, fb, b_k, tid);
        __syncthreads();
        if (k0 + kLT < K) {
            fa = fetch_frag(A, M, K, m0, k0 + kLT, a_m, a_k, tid);
            fb = fetch_frag(B, N, K, n0, k0 + kLT, b_n, b_k, tid);
        }
#pragma unroll
        for (int k = 0; k < kLT; ++k) {
            const float a0 = sa[k][ty], a1 = sa[k][ty + 16], b0 = sb[k][tx], b1 = sb[k][tx + 16];
            acc[0][0] = fmaf(a0, b0, acc[0][0]);  acc[0][1] = fmaf(a0, b1, acc[0][1]);
            acc[1][0] = fmaf(a1, b0, acc[1][0]);  acc[1][1] = fmaf(a1, b1, acc[1][1]);
            asum[0] += a0;  asum[1] += a1;
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int m = m0 + ty + 16 * i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int n = n0 + tx + 16 * j;
            if (n >= N) continue;
            float v = acc[i][j] + (bias ? bias[n] : 0.f);
            if (relu) v = fmaxf(v, 0.f);
            Cmat[(long)m * N + n] = v;
        }
        if (colsum_a && blockIdx.x == 0 && tx == 0) colsum_a[m] = asum[i];      // db[o] = sum_r dy[r, o] (the dW launch: A = dy^T)
    }
}

// dy <- dy * (y > 0)
__global__ void relu_mask_kernel(float* dy, const float* __restrict__ y, int64_t n) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n && !(y[i] > 0.f)) dy[i] = 0.f;
}

static int launch_gemm(const float* A, const float* B, const float* bias, float* C, int M, int N, int K, long a_m, long a_k, long b_n, long b_k,
                       int relu, float* colsum_a, cudaStream_t st, const char* what) {
    dim3 grid((N + kLT - 1) / kLT, (M + kLT - 1) / kLT);
    linear_gemm_kernel<<<grid, 256, 0, st>>>(A, B, bias, C, M, N, K, a_m, a_k, b_n, b_k, relu, colsum_a);
    return check_launch(what);
}

}  // namespace avid

using namespace avid;

extern "C" {

int avid_linear_forward(const float* x, const float* w, const float* b, float* y,
                        int32_t rows, int32_t in_f, int32_t out_f, int32_t relu, void* stream) {
    AVID_REQUIRE(x && w && y && rows > 0 && out_f > 0 && in_f > 0, "linear_forward: bad arguments");
    return launch_gemm(x, w, b, y, rows, out_f, in_f, in_f, 1, in_f, 1, relu, nullptr, static_cast<cudaStream_t>(stream), "linear_gemm_kernel(fwd)");
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
    if (dx && (rc = launch_gemm(dy, w, nullptr, dx, rows, in_f, out_f, out_f, 1, 1, in_f, 0, nullptr, st, "linear_gemm_kernel(dx)"))) return rc;
    return launch_gemm(dy, x, nullptr, dw, out_f, in_f, rows, 1, out_f, 1, in_f, 0, db, st, "linear_gemm_kernel(dw)");
}

}  // extern "C"
