// Probe: tiled TMA with SWIZZLE_128B whose innermost box extent is only 64 bytes, second dimension = 2 (a row PAIR, stride = one
// input row), third = wo (16-byte stride, Toeplitz), fourth = pair-row (stride two rows): does the box land as dense 128-byte rows
// [pair-row][wo][j][32] with the 128-byte swizzle on the linear address, and how many bytes does the transaction report?
//   nvcc -gencode arch=compute_100a,code=sm_100a -o pair_tma pair_tma.cu -lcuda && ./pair_tma
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
constexpr int kWo = 8, kP = 4, kBytes = kP * kWo * 2 * 64;      // dense expectation: 4 KB

__global__ void probe(const __grid_constant__ CUtensorMap map, int c2, int c3, int c4, uint32_t expect, uint16_t* out, int* info) {
    __shared__ __align__(1024) uint8_t tile[4 * kBytes];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    for (int i = threadIdx.x; i < 4 * kBytes / 2; i += blockDim.x) reinterpret_cast<uint16_t*>(tile)[i] = 0xFFFF;
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(expect));
        asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
                         smem_u32(tile)),
                     "l"(reinterpret_cast<uint64_t>(&map)), "r"(smem_u32(&bar)), "r"(0), "r"(0), "r"(c2), "r"(c3), "r"(c4)
                     : "memory");
    }
    uint32_t done = 0;
    long long t0 = clock64();
    while (!done && clock64() - t0 < 4000000LL) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(&bar)));
    }
    if (threadIdx.x == 0) *info = (int)done;
    __syncthreads();
    for (int i = threadIdx.x; i < 4 * kBytes / 2; i += blockDim.x) out[i] = reinterpret_cast<uint16_t*>(tile)[i];
}

int main() {
    const int F = 2, HP = 12, WO = 10, WP = 2 * WO + 8, C = 4;
    std::vector<__nv_bfloat16> x((size_t)F * HP * WP * C);
    auto code = [&](int f, int h, int w, int c) { return (float)(((f * 7 + h * 3 + w) * 4 + c) % 251 + 1); };
    for (int f = 0; f < F; ++f) for (int h = 0; h < HP; ++h) for (int w = 0; w < WP; ++w) for (int c = 0; c < C; ++c)
        x[(((size_t)f * HP + h) * WP + w) * C + c] = __float2bfloat16(code(f, h, w, c));
    __nv_bfloat16* dx;  uint16_t* dout;  int* dinfo;
    cudaMalloc(&dx, x.size() * 2);  cudaMalloc(&dout, 4 * kBytes);  cudaMalloc(&dinfo, 4);
    cudaMemcpy(dx, x.data(), x.size() * 2, cudaMemcpyHostToDevice);
    typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;  cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    EncodeTiled enc = (EncodeTiled)fn;
    CUtensorMap map;
    const cuuint64_t row = (cuuint64_t)WP * C * 2;
    cuuint64_t dims[5] = {32, 2, (cuuint64_t)WO, (cuuint64_t)HP / 2, (cuuint64_t)F};
    cuuint64_t strides[4] = {row, 16, 2 * row, row * HP};
    cuuint32_t box[5] = {32, 2, kWo, kP, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, dx, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d\n", (int)r);
    if (r != CUDA_SUCCESS) return 1;
    std::vector<uint16_t> out(4 * kBytes / 2);
    for (uint32_t expect : {(uint32_t)kBytes, (uint32_t)(2 * kBytes)}) {
        const int c2 = 1, c3 = -1, c4 = 1;       // wo0 = 1, first pair-row -1 (zero fill), frame 1
        probe<<<1, 128>>>(map, c2, c3, c4, expect, dout, dinfo);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("kernel: %s\n", cudaGetErrorString(e)); return 1; }
        int info = 0;
        cudaMemcpy(out.data(), dout, out.size() * 2, cudaMemcpyDeviceToHost);
        cudaMemcpy(&info, dinfo, 4, cudaMemcpyDeviceToHost);
        int written = 0, last = -1;
        for (size_t i = 0; i < out.size(); ++i) if (out[i] != 0xFFFF) { ++written; last = (int)i; }
        printf("expect_tx %u: barrier completed %d, %d bytes touched, last touched byte %d\n", expect, info, written * 2, last * 2 + 1);
        // dense-layout hypothesis: 128-byte row q = (p * kWo + wo), halves j, 16-byte chunk swizzled by (q % 8)
        int bad = 0;
        for (int p = 0; p < kP; ++p) for (int wo = 0; wo < kWo; ++wo) for (int j = 0; j < 2; ++j) for (int k = 0; k < 32; ++k) {
            const int qrow = p * kWo + wo;
            const uint32_t logical = qrow * 128 + j * 64 + k * 2;
            const uint32_t phys = logical ^ ((qrow & 7) << 4);
            __nv_bfloat16 v;  *reinterpret_cast<uint16_t*>(&v) = out[phys / 2];
            const int h = 2 * (c3 + p) + j, woo = c2 + wo;
            const bool inb = (c3 + p) >= 0 && (c3 + p) < HP / 2 && woo < WO;
            const float want = inb ? code(c4, h, 2 * woo + k / 4, k % 4) : 0.f;
            if (__bfloat162float(v) != want) { if (bad < 4) printf("  p %d wo %d j %d k %d got %g want %g\n", p, wo, j, k, __bfloat162float(v), want); ++bad; }
        }
        printf("  dense [p][wo][j][32] + SW128 hypothesis: %d mismatches\n", bad);
    }
    return 0;
}
