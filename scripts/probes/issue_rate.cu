// Probe: what does the MMA-ISSUING THREAD cost?  The same chain of M128 x N x K16 bf16 MMAs issued (a) from `if (threadIdx.x == 0)` -- ptxas
// cannot prove a single active thread and wraps every UTCHMMA in an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall -- and (b) from a whole
// warp running the loop uniformly with each MMA guarded by elect.sync (operands provably warp-uniform -> uniform registers).
//   nvcc -gencode arch=compute_100a,code=sm_100a -o issue_rate issue_rate.cu -lcuda && ./issue_rate
#include <stdio.h>
#include "../../avid_cma_b200/csrc/tc_common.cuh"
using namespace avid::tc;

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

template <int MODE, int N>
__global__ void __launch_bounds__(128, 1) probe(int iters, long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(&slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const uint32_t a = smem_u32(smem), b = smem_u32(smem) + 64 * 1024;
    constexpr uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
    if (MODE == 0) {
        if (threadIdx.x == 0) {
            const uint32_t tmem = slot;
            const uint64_t da = make_smem_desc_sw128(a, 16, 1024), db = make_smem_desc_sw128(b, 16, 1024);
            const long long t0 = clock64();
            for (int i = 0; i < iters; ++i) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem, da + (uint32_t)(2 * k) + (uint32_t)((i & 7) * 1024 >> 4), db + (uint32_t)(2 * k), idesc, 1);
            }
            umma_commit(&bar);
            mbar_wait(&bar, 0);
            *cycles = clock64() - t0;
        }
    } else {
        if (warp == 0) {
            const uint32_t tmem = __shfl_sync(0xffffffffu, slot, 0);
            const uint64_t da = make_smem_desc_sw128(a, 16, 1024), db = make_smem_desc_sw128(b, 16, 1024);
            const long long t0 = clock64();
            for (int i = 0; i < iters; ++i) {
                if (MODE == 1) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (elect_one()) umma_bf16(tmem, da + (uint32_t)(2 * k) + (uint32_t)((i & 7) * 1024 >> 4), db + (uint32_t)(2 * k), idesc, 1);
                } else if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16(tmem, da + (uint32_t)(2 * k) + (uint32_t)((i & 7) * 1024 >> 4), db + (uint32_t)(2 * k), idesc, 1);
                }
            }
            if (elect_one()) umma_commit(&bar);
            mbar_wait(&bar, 0);
            if (threadIdx.x == 0) *cycles = clock64() - t0;
        }
    }
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(slot, 512);
}

template <int MODE, int N>
void run(long long* d, const char* name) {
    const int smem = 200 * 1024;
    cudaFuncSetAttribute(probe<MODE, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 2000;
    probe<MODE, N><<<1, 128, smem>>>(10, d);
    probe<MODE, N><<<1, 128, smem>>>(iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s N=%d: %s\n", name, N, cudaGetErrorString(e)); return; }
    long long c;
    cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    const double per = (double)c / (iters * 4);
    printf("%-44s N=%3d: %6.1f cycles per MMA (math %d, operand fetch at 128 B/clk %d)\n", name, N, per, N / 2, (128 + N) * 32 / 128);
}

int main() {
    long long* d;
    cudaMalloc(&d, 8);
    run<0, 64>(d, "if (threadIdx.x == 0)");
    run<1, 64>(d, "warp-uniform loop, elect.sync per MMA");
    run<2, 64>(d, "warp-uniform loop, elect.sync per 4 MMAs");
    run<0, 128>(d, "if (threadIdx.x == 0)");
    run<1, 128>(d, "warp-uniform loop, elect.sync per MMA");
    run<2, 128>(d, "warp-uniform loop, elect.sync per 4 MMAs");
    run<0, 256>(d, "if (threadIdx.x == 0)");
    run<1, 256>(d, "warp-uniform loop, elect.sync per MMA");
    run<2, 256>(d, "warp-uniform loop, elect.sync per 4 MMAs");
    return 0;
}
