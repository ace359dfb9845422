// Probe: issue rate of tcgen05.mma.cta_group::2 (M = 256 over a CTA pair) against cta_group::1 (M = 128), K-major SW128 operands in
// (zeroed) shared memory.  Variants: A descriptor fixed / moving by whole atoms / moving by single 128-byte rows (halo-style tap shifts);
// B descriptor fixed / walking through a resident filter.  One or 74 clusters.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o pair_rate pair_rate.cu -lcuda && ./pair_rate
#include <stdio.h>
#include "../../avid_cma_b200/csrc/tc_common.cuh"
using namespace avid::tc;

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(da), "l"(db),
                 "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

// pair = 1: cta_group::2, M = 256, N = n (each CTA holds n / 2 rows of B); pair = 0: every CTA on its own, M = 128
// amode: 0 fixed A, 1 A moves by 1 KB atoms, 2 A moves by 128-byte rows (mid-atom starts);  bmode: 0 fixed B, 1 B walks 4 KB tiles
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) probe(int pair, int n, int amode, int bmode, int iters, long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    const uint32_t rank = cluster_ctarank();
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) {
        if (pair) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(256) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            tmem_alloc(&slot, 256);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (threadIdx.x == 0 && (!pair || rank == 0)) {
        const uint32_t a = smem_u32(smem), b = smem_u32(smem) + 96 * 1024;
        const uint64_t da = make_smem_desc_sw128(a, 16, 1024), db = make_smem_desc_sw128(b, 16, 1024);
        const uint32_t idesc = make_idesc_bf16(pair ? 256 : 128, n, 0, 0);
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t ao = amode == 0 ? 0u : amode == 1 ? (uint32_t)((i % 9) * 1024 >> 4) : (uint32_t)((i % 3) * 58 + (i / 3) % 3) * 8u;
            const uint32_t bo = bmode ? (uint32_t)((i % 9) * 8192 >> 4) : 0u;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (pair) umma_bf16_pair(tmem, da + ao + 2 * k, db + bo + 2 * k, idesc, 1);
                else umma_bf16(tmem, da + ao + 2 * k, db + bo + 2 * k, idesc, 1);
            }
        }
        if (pair) umma_commit_pair(&bar, 3); else umma_commit(&bar);
        mbar_wait(&bar, 0);
        if (blockIdx.x == 0) *cycles = clock64() - t0;
    } else if (threadIdx.x == 0) {
        mbar_wait(&bar, 0);
    }
    __syncthreads();
    tc_fence_before();
    cluster_sync();
    if (threadIdx.x < 32) {
        if (pair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
        else tmem_dealloc(tmem, 256);
    }
}

int main() {
    long long* d;
    cudaMalloc(&d, 8);
    const int smem = 202 * 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const char* an[3] = {"A fixed", "A by atoms", "A by rows"};
    for (int grid : {2, 148})
        for (int pair = 0; pair < 2; ++pair)
            for (int n : {64, 128, 256})
                for (int amode = 0; amode < 3; ++amode)
                    for (int bmode = 0; bmode < 2; ++bmode) {
                        if (n != 64 && (amode == 1 || bmode == 1)) continue;
                        const int iters = 2000;
                        probe<<<grid, 128, smem>>>(pair, n, amode, bmode, 10, d);
                        probe<<<grid, 128, smem>>>(pair, n, amode, bmode, iters, d);
                        cudaError_t e = cudaDeviceSynchronize();
                        if (e != cudaSuccess) { printf("pair %d N=%d: %s\n", pair, n, cudaGetErrorString(e)); return 1; }
                        long long c;
                        cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
                        printf("grid %3d  %s  N=%3d  %-10s %-8s: %6.1f cycles per MMA (math per SM %d)\n", grid, pair ? "cta_group::2 M256" : "cta_group::1 M128", n,
                               an[amode], bmode ? "B walks" : "B fixed", (double)c / (iters * 4), n / 2);
                    }
    return 0;
}
