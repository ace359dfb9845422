// Memory-bank maintenance kernels (include/avid_b200.h): momentum update of the rows touched by a
// step (criterions/avid.py:103-129) and row-wise L2 normalisation (avid.py:92,95).  One warp per
// 512-byte row, one float4 per lane: every access is a fully coalesced 512-byte request.
#include "common.cuh"

namespace avid {

__device__ __forceinline__ float4 scale4(const float4& v, float s) { return make_float4(v.x * s, v.y * s, v.z * s, v.w * s); }
__device__ __forceinline__ float sq4(const float4& v) { return v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w; }

// warp w handles (instance i = w / 2, bank = w % 2)
__global__ void __launch_bounds__(256) bank_update_kernel(float* bank_v, float* bank_a, int64_t row_begin, int64_t row_end,
                                                          const float* emb_v, const float* emb_a, const int64_t* y, int n,
                                                          float mom_v, float mom_a) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= 2 * n) return;
    const int i = w >> 1, which = w & 1;
    const int64_t row = y[i];
    if (row < row_begin || row >= row_end) return;
    float* bank = which ? bank_a : bank_v;
    const float* emb = which ? emb_a : emb_v;
    const float mom = which ? mom_a : mom_v;
    float4 e = reinterpret_cast<const float4*>(emb + (size_t)i * kD)[lane];
    e = scale4(e, 1.0f / fmaxf(sqrtf(warp_sum(sq4(e))), 1e-12f));           // F.normalize of the embedding (avid.py:52-53)
    float4* dst = reinterpret_cast<float4*>(bank + (size_t)(row - row_begin) * kD) + lane;
    float4 m = *dst;
    const float om = 1.0f - mom;
    m = make_float4(m.x * mom + e.x * om, m.y * mom + e.y * om, m.z * mom + e.z * om, m.w * mom + e.w * om);
    m = scale4(m, 1.0f / fmaxf(sqrtf(warp_sum(sq4(m))), 1e-12f));           // avid.py:122,128
    *dst = m;
}

__global__ void __launch_bounds__(256) rows_normalize_kernel(float* x, int64_t rows) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps) {
        float4* p = reinterpret_cast<float4*>(x + (size_t)r * kD) + lane;
        float4 v = *p;
        *p = scale4(v, 1.0f / fmaxf(sqrtf(warp_sum(sq4(v))), 1e-12f));
    }
}

}  // namespace avid

using namespace avid;

extern "C" {

int avid_bank_update(float* bank_video, float* bank_audio, int64_t row_begin, int64_t row_end,
                     const float* emb_video, const float* emb_audio, const int64_t* y, int32_t n,
                     float momentum_video, float momentum_audio, void* stream) {
    AVID_REQUIRE(bank_video && bank_audio && emb_video && emb_audio && y, "bank_update: NULL pointer");
    AVID_REQUIRE(n > 0 && row_begin >= 0 && row_begin < row_end, "bank_update: bad sizes");
    const int warps = 2 * n;
    bank_update_kernel<<<(warps * 32 + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        bank_video, bank_audio, row_begin, row_end, emb_video, emb_audio, y, n, momentum_video, momentum_audio);
    return check_launch("bank_update_kernel");
}

int avid_rows_l2_normalize(float* x, int64_t rows, void* stream) {
    AVID_REQUIRE(x && rows > 0, "rows_l2_normalize: bad arguments");
    int64_t blocks = (rows + 7) / 8;
    if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
    rows_normalize_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, rows);
    return check_launch("rows_normalize_kernel");
}

}  // extern "C"
