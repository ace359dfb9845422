// Probe: tensor-core operand fetch rate by shared-memory layout.  One thread issues a long chain of M128 x N x K16 bf16 MMAs whose A / B
// descriptors point at (garbage) shared memory in 64-byte-swizzle rows (K = 32 per row) or 128-byte-swizzle rows (K = 64 per row),
// K-major or MN-major; cycles per MMA against the 32 (N = 64) / 64 (N = 128) cycles of pure math.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o umma_fetch umma_fetch.cu -lcuda && ./umma_fetch
#include <stdio.h>
#include "../../avid_cma_b200/csrc/tc_common.cuh"
using namespace avid::tc;

__global__ void __launch_bounds__(128, 1) probe(int mode, int n, int iters, long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(&slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 0) {
        const uint32_t a = smem_u32(smem), b = smem_u32(smem) + 64 * 1024;
        uint64_t da, db;
        uint32_t idesc;
        uint32_t step_a, step_b;     // descriptor advance per K = 16 step (encoded >> 4)
        if (mode == 0) {             // K-major, SW64 rows (K = 32): the round-1 stem forward
            da = make_smem_desc_sw64(a, 16, 512);  db = make_smem_desc_sw64(b, 16, 512);  idesc = make_idesc_bf16(128, n, 0, 0);  step_a = step_b = 2;
        } else if (mode == 1) {      // K-major, SW128 rows (K = 64)
            da = make_smem_desc_sw128(a, 16, 1024);  db = make_smem_desc_sw128(b, 16, 1024);  idesc = make_idesc_bf16(128, n, 0, 0);  step_a = step_b = 2;
        } else if (mode == 2) {      // MN-major A SW64 (atoms 1 KB apart), MN-major B SW128: the round-1 stem wgrad
            da = make_smem_desc_sw64(a, 1024, 512);  db = make_smem_desc_sw128(b, 16384, 1024);  idesc = make_idesc_bf16(128, n, 1, 1);  step_a = 64; step_b = 128;
        } else {                     // MN-major A SW128 (MN groups 1 KB apart), MN-major B SW128
            da = make_smem_desc_sw128(a, 1024, 1024);  db = make_smem_desc_sw128(b, 16384, 1024);  idesc = make_idesc_bf16(128, n, 1, 1);  step_a = 128; step_b = 128;
        }
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(slot, da + (uint32_t)(k * step_a) + (uint32_t)((i & 7) * 1024 >> 4), db + (uint32_t)(k * step_b), idesc, 1);
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        *cycles = clock64() - t0;
    }
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(slot, 256);
}

int main() {
    long long* d;
    cudaMalloc(&d, 8);
    const int smem = 200 * 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const char* names[4] = {"K-major  SW64  A, SW64  B", "K-major  SW128 A, SW128 B", "MN-major SW64  A, SW128 B", "MN-major SW128 A, SW128 B"};
    for (int mode = 0; mode < 4; ++mode)
        for (int n : {64, 128}) {
            const int iters = 2000;
            probe<<<1, 128, smem>>>(mode, n, 10, d);
            probe<<<1, 128, smem>>>(mode, n, iters, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s N=%d: %s\n", names[mode], n, cudaGetErrorString(e)); return 1; }
            long long c;
            cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
            const double per = (double)c / (iters * 4);
            printf("%s  N=%3d: %6.1f cycles per MMA (math %d), operand bytes %d -> %.0f B/clk\n", names[mode], n, per, n / 2, (128 + n) * 32, (128 + n) * 32 / per);
        }
    return 0;
}
