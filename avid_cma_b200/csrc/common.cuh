// Shared helpers for the sm_100a kernels behind include/avid_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "../../include/avid_b200.h"

namespace avid {

constexpr int kNumSMs = 148;      // B200: 2 dies x 74 SMs
constexpr int kD = AVID_EMB_DIM;  // 128

void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return AVID_ECUDA;
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return AVID_OK;
}

#define AVID_REQUIRE(cond, ...)                 \
    do {                                        \
        if (!(cond)) {                          \
            ::avid::set_error(__VA_ARGS__);     \
            return AVID_EINVAL;                 \
        }                                       \
    } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// streaming 16-byte load that does not pollute L1 (rows of the bank are touched once)
__device__ __forceinline__ float4 ld_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

// ---- Philox4x32-10 (Salmon et al. 2011); counter-based, one call -> 4 x 32 random bits ----
struct Philox {
    static constexpr uint32_t kM0 = 0xD2511F53u, kM1 = 0xCD9E8D57u;
    static constexpr uint32_t kW0 = 0x9E3779B9u, kW1 = 0xBB67AE85u;
    __host__ __device__ static inline void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
        uint64_t p0 = (uint64_t)kM0 * c[0];
        uint64_t p1 = (uint64_t)kM1 * c[2];
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c[1] ^ k0;
        uint32_t n1 = lo1;
        uint32_t n2 = hi0 ^ c[3] ^ k1;
        uint32_t n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    }
    __host__ __device__ static inline void generate(uint64_t seed, uint64_t ctr, uint32_t (&out)[4]) {
        uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u};
        uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            round(c, k0, k1);
            k0 += kW0; k1 += kW1;
        }
        out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
    }
};

// uniform integer in [0, range) from 64 random bits (multiply-high; bias < range * 2^-64)
__host__ __device__ inline uint64_t uniform_below(uint32_t r0, uint32_t r1, uint64_t range) {
    uint64_t r = ((uint64_t)r1 << 32) | r0;
#if defined(__CUDA_ARCH__)
    return __umul64hi(r, range);
#else
    return (uint64_t)(((unsigned __int128)r * range) >> 64);
#endif
}

// The negative for instance b, slot k: reference semantics
//   AVID     (avid.py:82-86):        r ~ U[0, N-1);            idx = r + (r >= y)
//   AVID-CMA (avid_cma.py:200-207):  r ~ U[0, N-pos_k);        idx = r + #{j : r >= pos[j] - j}
// the y- / positive-set-dependent half of a draw, from the 64 random bits of its counter
__host__ __device__ inline int64_t finish_negative(uint32_t r0, uint32_t r1, int64_t N, int64_t y, const int32_t* pos_row, int pos_k) {
    if (pos_row == nullptr) {
        int64_t r = (int64_t)uniform_below(r0, r1, (uint64_t)(N - 1));
        return r + (r >= y ? 1 : 0);
    }
    int64_t r = (int64_t)uniform_below(r0, r1, (uint64_t)(N - pos_k));
    int shift = 0;
    for (int j = 0; j < pos_k; ++j) shift += (r >= (int64_t)pos_row[j] - j) ? 1 : 0;
    return r + shift;
}

__host__ __device__ inline int64_t draw_negative(uint64_t seed, uint64_t offset, int b, int k, int K,
                                                 int64_t N, int64_t y, const int32_t* pos_row, int pos_k) {
    uint32_t rnd[4];
    Philox::generate(seed, offset + (uint64_t)b * (uint64_t)K + (uint64_t)k, rnd);
    return finish_negative(rnd[0], rnd[1], N, y, pos_row, pos_k);
}

}  // namespace avid
