// Probe: TMA tile::gather4 on a row-major fp32 matrix [N][128] (the memory bank): one instruction fetches 4 arbitrary rows.
// Which box shape does the tensor map need ({128, 1} or {128, 4}), how do the rows land in shared memory, how many bytes complete?
//   nvcc -gencode arch=compute_100a,code=sm_100a -o gather4_tma gather4_tma.cu -lcuda && ./gather4_tma
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap map, int r0, int r1, int r2, int r3, float* out, int* info) {
    __shared__ __align__(1024) float tile[8 * 128];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    for (int i = threadIdx.x; i < 8 * 128; i += blockDim.x) tile[i] = -1.f;
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(4 * 512));
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
                         smem_u32(tile)),
                     "l"(reinterpret_cast<uint64_t>(&map)), "r"(smem_u32(&bar)), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
                     : "memory");
    }
    uint32_t done = 0;
    long long t0 = clock64();
    while (!done && clock64() - t0 < 20000000LL) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(&bar)));
    }
    if (threadIdx.x == 0) *info = (int)done;
    __syncthreads();
    for (int i = threadIdx.x; i < 8 * 128; i += blockDim.x) out[i] = tile[i];
}

int main() {
    const int N = 4096;
    std::vector<float> x((size_t)N * 128);
    for (int r = 0; r < N; ++r) for (int c = 0; c < 128; ++c) x[(size_t)r * 128 + c] = (float)(r * 1000 + c);
    float *dx, *dout;  int* dinfo;
    cudaMalloc(&dx, x.size() * 4);  cudaMalloc(&dout, 8 * 128 * 4);  cudaMalloc(&dinfo, 4);
    cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice);
    typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;  cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    EncodeTiled enc = (EncodeTiled)fn;
    const int rows[4] = {5, 900, 17, 3333};
    for (int boxrows = 1; boxrows <= 4; boxrows += 3) {
        CUtensorMap map;
        cuuint64_t dims[2] = {128, (cuuint64_t)N};  cuuint64_t strides[1] = {512};  cuuint32_t box[2] = {128, (cuuint32_t)boxrows};  cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dx, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("box rows %d: encode rc=%d\n", boxrows, (int)r);
        if (r != CUDA_SUCCESS) continue;
        probe<<<1, 128>>>(map, rows[0], rows[1], rows[2], rows[3], dout, dinfo);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("  kernel: %s\n", cudaGetErrorString(e)); cudaGetLastError(); return 1; }
        std::vector<float> out(8 * 128);  int info = 0;
        cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(&info, dinfo, 4, cudaMemcpyDeviceToHost);
        printf("  barrier completed: %d\n", info);
        for (int slot = 0; slot < 8; ++slot) printf("  smem row %d: [0]=%.0f [1]=%.0f [127]=%.0f\n", slot, out[slot * 128], out[slot * 128 + 1], out[slot * 128 + 127]);
        int ok = 1;
        for (int i = 0; i < 4; ++i) for (int c = 0; c < 128; ++c) ok &= out[i * 128 + c] == (float)(rows[i] * 1000 + c);
        printf("  GATHER4 box rows %d: %s\n", boxrows, ok ? "OK (rows land consecutively, 512 B each)" : "MISMATCH");
    }
    return 0;
}
