// Memory-bank maintenance kernels (include/avid_b200.h): momentum update of the rows touched by a
// step (criterions/avid.py:103-129) and row-wise L2 normalisation (avid.py:92,95).  One warp per
// 512-byte row, one float4 per lane: every access is a fully coalesced 512-byte request.
#include "common.cuh"

namespace avid {

__device__ __forceinline__ float4 scale4(const float4& v, float s) { return make_float4(v.x * s, v.y * s, v.z * s, v.w * s); }
__device__ __forceinline__ float sq4(const float4& v) { return v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w; }

// warp w handles (instance i = w / 2, bank = w % 2)
// instance i of the n gathered instances = row (i % gb) of rank-record (i / gb); records are `stride` bytes apart (gb = n, stride 0: dense)
__device__ __forceinline__ const char* rec_ptr(const void* base, int i, int gb, size_t stride, size_t elem_bytes) {
    const int g = i / gb;
    return static_cast<const char*>(base) + (size_t)g * stride + (size_t)(i - g * gb) * elem_bytes;
}

__global__ void __launch_bounds__(256) bank_update_kernel(float* bank_v, float* bank_a, int64_t row_begin, int64_t row_end,
                                                          const float* emb_v, const float* emb_a, const int64_t* y, int n, int gb, size_t stride,
                                                          float mom_v, float mom_a) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= 2 * n) return;
    const int i = w >> 1, which = w & 1;
    const int64_t row = *reinterpret_cast<const int64_t*>(rec_ptr(y, i, gb, stride, 8));
    if (row < row_begin || row >= row_end) return;
    // Duplicate instance ids in the gathered batch (DistributedSampler padding, clips_per_video > 1): the reference's
    // index_select + index_copy_ leaves ONE complete update per row (avid.py:119-129).  Here the LAST occurrence owns the row
    // (two warps doing the read-modify-write of the same 512 bytes would interleave their lanes and tear it).
    bool later = false;
    for (int j = i + 1 + lane; j < n; j += 32) later |= (*reinterpret_cast<const int64_t*>(rec_ptr(y, j, gb, stride, 8)) == row);
    if (__any_sync(0xffffffffu, later)) return;
    float* bank = which ? bank_a : bank_v;
    const float* emb = which ? emb_a : emb_v;
    const float mom = which ? mom_a : mom_v;
    float4 e = reinterpret_cast<const float4*>(rec_ptr(emb, i, gb, stride, kD * 4))[lane];
    e = scale4(e, 1.0f / fmaxf(sqrtf(warp_sum(sq4(e))), 1e-12f));           // F.normalize of the embedding (avid.py:52-53)
    float4* dst = reinterpret_cast<float4*>(bank + (size_t)(row - row_begin) * kD) + lane;
    float4 m = *dst;
    const float om = 1.0f - mom;
    m = make_float4(m.x * mom + e.x * om, m.y * mom + e.y * om, m.z * mom + e.z * om, m.w * mom + e.w * om);
    m = scale4(m, 1.0f / fmaxf(sqrtf(warp_sum(sq4(m))), 1e-12f));           // avid.py:122,128
    *dst = m;
}

// init_memory (avid.py:88-96): N(0,1) rows, L2-normalised.  Row r of bank `which` depends only on (seed, which, r): Philox
// counter (r * 32 + lane), Box-Muller on the four 32-bit words.  Every rank of a sharded run fills its own rows and a replicated
// run fills all of them -- the same values either way, so no (N,128) broadcast from rank 0 is needed.
__global__ void __launch_bounds__(256) bank_init_kernel(float* bank, int64_t row_begin, int64_t rows, uint64_t seed, int which) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps) {
        uint32_t w[4];
        Philox::generate(seed ^ (which ? 0x9E3779B97F4A7C15ull : 0ull), (uint64_t)(row_begin + r) * 32u + (uint64_t)lane, w);
        const float kInv = 2.3283064365386963e-10f;      // 2^-32
        const float u0 = ((float)w[0] + 0.5f) * kInv, u1 = (float)w[1] * kInv, u2 = ((float)w[2] + 0.5f) * kInv, u3 = (float)w[3] * kInv;
        const float r0 = sqrtf(-2.0f * logf(fminf(u0, 1.0f))), r1 = sqrtf(-2.0f * logf(fminf(u2, 1.0f)));
        float s0, c0, s1, c1;
        sincospif(2.0f * u1, &s0, &c0);
        sincospif(2.0f * u3, &s1, &c1);
        float4 v = make_float4(r0 * c0, r0 * s0, r1 * c1, r1 * s1);
        v = scale4(v, 1.0f / fmaxf(sqrtf(warp_sum(sq4(v))), 1e-12f));
        reinterpret_cast<float4*>(bank + (size_t)r * kD)[lane] = v;
    }
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
                     const float* emb_video, const float* emb_audio, const int64_t* y, int32_t n, int32_t group_batch, int64_t group_stride,
                     float momentum_video, float momentum_audio, void* stream) {
    AVID_REQUIRE(bank_video && bank_audio && emb_video && emb_audio && y, "bank_update: NULL pointer");
    AVID_REQUIRE(n > 0 && row_begin >= 0 && row_begin < row_end, "bank_update: bad sizes");
    AVID_REQUIRE(group_batch >= 0 && group_stride >= 0 && group_stride % 16 == 0 && (group_batch == 0 || n % group_batch == 0),
                 "bank_update: bad record layout (group_batch %d, stride %lld)", group_batch, (long long)group_stride);
    const int warps = 2 * n;
    bank_update_kernel<<<(warps * 32 + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        bank_video, bank_audio, row_begin, row_end, emb_video, emb_audio, y, n, group_batch > 0 ? group_batch : n,
        group_batch > 0 ? (size_t)group_stride : 0, momentum_video, momentum_audio);
    return check_launch("bank_update_kernel");
}

int avid_bank_init(float* bank, int64_t row_begin, int64_t rows, uint64_t seed, int32_t which, void* stream) {
    AVID_REQUIRE(bank && rows > 0 && row_begin >= 0 && (which == 0 || which == 1), "bank_init: bad arguments");
    int64_t blocks = (rows + 7) / 8;
    if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
    bank_init_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(bank, row_begin, rows, seed, which);
    return check_launch("bank_init_kernel");
}

int avid_rows_l2_normalize(float* x, int64_t rows, void* stream) {
    AVID_REQUIRE(x && rows > 0, "rows_l2_normalize: bad arguments");
    int64_t blocks = (rows + 7) / 8;
    if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
    rows_normalize_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, rows);
    return check_launch("rows_normalize_kernel");
}

}  // extern "C"
