// Probe: tcgen05.mma.cta_group::2 (a CTA pair, M = 256) with hand-written PTX: cluster launch, cta_group::2 TMEM allocation, TMA loads
// in both CTAs signalling the LEADER's mbarrier, MMA issued by the leader with B split over the pair (N / 2 rows each), multicast
// commit, epilogue per CTA.  D[256][64] = A[256][64] * B[64][64]^T, bf16 in, fp32 out, exact small integers.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o pair_mma pair_mma.cu -lcuda && ./pair_mma
#include <stdio.h>
#include <vector>
#include "../../avid_cma_b200/csrc/tc_common.cuh"
using namespace avid::tc;

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    // executed by both CTAs of the pair; the transaction bytes update the barrier of CTA 0 (peer bit 24 of the cluster address cleared)
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(da), "l"(db),
                 "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
probe(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, float* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* a_s = smem;                  // [128 rows][64 k] bf16 SW128: this CTA's half of M
    uint8_t* b_s = smem + 16384;          // [32 rows][64 k]: this CTA's half of N
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + 16384 + 4096);
    uint64_t* done = full + 1;
    uint32_t* slot = reinterpret_cast<uint32_t*>(done + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    if (threadIdx.x == 0) {
        mbar_init(full, 1);
        mbar_init(done, 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();                       // both CTAs' barriers are initialised before anything signals them
    tc_fence_after();
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        if (rank == 0) mbar_expect_tx(full, 2 * (16384 + 4096));       // the leader's barrier counts the bytes of both CTAs
        tma_load_2d_pair(a_s, &map_a, full, 0, (int)rank * 128);
        tma_load_2d_pair(b_s, &map_b, full, 0, (int)rank * 32);
        if (rank == 0) {
            mbar_wait(full, 0);
            tc_fence_after();
            const uint64_t da = make_smem_desc_sw128(smem_u32(a_s), 16, 1024), db = make_smem_desc_sw128(smem_u32(b_s), 16, 1024);
            const uint32_t idesc = make_idesc_bf16(256, 64, 0, 0);
            for (int k = 0; k < 4; ++k) umma_bf16_pair(tmem, da + 2 * k, db + 2 * k, idesc, k != 0);
            umma_commit_pair(done, 3);    // arrive on `done` of both CTAs
        }
    }
    mbar_wait(done, 0);
    tc_fence_after();
    uint32_t r[32];
    for (int j = 0; j < 2; ++j) {
        tmem_ld_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16) + j * 32, r);
        tmem_ld_wait();
        for (int c = 0; c < 32; ++c) out[((size_t)rank * 128 + warp * 32 + lane) * 64 + j * 32 + c] = __uint_as_float(r[c]);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64) : "memory");
}

int main() {
    std::vector<__nv_bfloat16> a(256 * 64), b(64 * 64);
    auto av = [](int m, int k) { return (float)((m * 5 + k * 3) % 7 - 3); };
    auto bv = [](int n, int k) { return (float)((n * 3 + k) % 5 - 2); };
    for (int m = 0; m < 256; ++m) for (int k = 0; k < 64; ++k) a[m * 64 + k] = __float2bfloat16(av(m, k));
    for (int n = 0; n < 64; ++n) for (int k = 0; k < 64; ++k) b[n * 64 + k] = __float2bfloat16(bv(n, k));
    __nv_bfloat16 *da, *db;  float* dout;
    cudaMalloc(&da, a.size() * 2);  cudaMalloc(&db, b.size() * 2);  cudaMalloc(&dout, 256 * 64 * 4);
    cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dout, 0xFF, 256 * 64 * 4);
    void* fn = nullptr;  cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    avid::TensorMapApi::EncodeTiled enc = (avid::TensorMapApi::EncodeTiled)fn;
    CUtensorMap ma, mb;
    cuuint64_t strides[1] = {128};  cuuint32_t es[2] = {1, 1};
    cuuint64_t adims[2] = {64, 256};  cuuint32_t abox[2] = {64, 128};
    cuuint64_t bdims[2] = {64, 64};   cuuint32_t bbox[2] = {64, 32};
    CUresult r1 = enc(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, da, adims, strides, abox, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = enc(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, db, bdims, strides, bbox, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d %d\n", (int)r1, (int)r2);
    const int smem = 16384 + 4096 + 1024 + 256;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe<<<2, 128, smem>>>(ma, mb, dout);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<float> out(256 * 64);
    cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int m = 0; m < 256; ++m) for (int n = 0; n < 64; ++n) {
        float want = 0.f;
        for (int k = 0; k < 64; ++k) want += av(m, k) * bv(n, k);
        if (out[m * 64 + n] != want) { if (bad < 6) printf("  m %d n %d got %g want %g\n", m, n, out[m * 64 + n], want); ++bad; }
    }
    printf("PAIR_MMA %s (%d mismatches)\n", bad ? "FAIL" : "OK", bad);
    return 0;
}
