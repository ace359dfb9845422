// Probe: tiled TMA over an OVERLAPPING-stride ("Toeplitz") view of a W-padded channels-last tensor [n][t][h][wp][4] bf16:
//   dim0 = 32 elements (8 consecutive w pixels x 4 channels), dim1 = wo with a 16-byte stride (= 2 pixels: conv stride 2),
//   dim2 = h (traversed with elementStride 2), dim3 = t, dim4 = n;  box (32, 16, 16/2, 1, 1), SWIZZLE_64B.
// Checks that the driver accepts the map and that smem holds A[pixel = (ho, wo)][k = (kw, c)] with zero fill outside h / t / wo.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o toeplitz_tma toeplitz_tma.cu -lcuda && ./toeplitz_tma
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap map, int c1, int c2, int c3, int c4, uint16_t* out) {
    __shared__ __align__(1024) uint8_t tile[128 * 64];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    for (int i = threadIdx.x; i < 128 * 32; i += blockDim.x) reinterpret_cast<uint16_t*>(tile)[i] = 0xFFFF;
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(128 * 64));
        asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
                         smem_u32(tile)),
                     "l"(reinterpret_cast<uint64_t>(&map)), "r"(smem_u32(&bar)), "r"(0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
                     : "memory");
    }
    uint32_t done = 0;
    while (!done) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(&bar)));
    }
    for (int i = threadIdx.x; i < 128 * 32; i += blockDim.x) out[i] = reinterpret_cast<uint16_t*>(tile)[i];
}

int main() {
    const int N = 2, T = 3, H = 20, WO = 10, WP = 2 * WO + 8, C = 4;
    std::vector<__nv_bfloat16> x((size_t)N * T * H * WP * C);
    // value = small integer code exactly representable in bf16: (h % 16) * 16 + (wp % 16) scaled by channel sign... use hash mod 256
    auto code = [&](int n, int t, int h, int w, int c) { return (float)(((n * 7 + t * 5 + h * 3 + w) * 4 + c) % 251 + 1); };
    for (int n = 0; n < N; ++n) for (int t = 0; t < T; ++t) for (int h = 0; h < H; ++h) for (int w = 0; w < WP; ++w) for (int c = 0; c < C; ++c)
        x[((((size_t)n * T + t) * H + h) * WP + w) * C + c] = __float2bfloat16(code(n, t, h, w, c));
    __nv_bfloat16* dx;  uint16_t* dout;
    cudaMalloc(&dx, x.size() * 2);  cudaMalloc(&dout, 128 * 64);
    cudaMemcpy(dx, x.data(), x.size() * 2, cudaMemcpyHostToDevice);

    typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;  cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    EncodeTiled enc = (EncodeTiled)fn;
    CUtensorMap map;
    cuuint64_t dims[5] = {32, (cuuint64_t)WO, (cuuint64_t)H, (cuuint64_t)T, (cuuint64_t)N};
    cuuint64_t strides[4] = {16, (cuuint64_t)WP * C * 2, (cuuint64_t)H * WP * C * 2, (cuuint64_t)T * H * WP * C * 2};
    cuuint32_t box[5] = {32, 16, 16, 1, 1};
    cuuint32_t estr[5] = {1, 1, 2, 1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, dx, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d\n", (int)r);
    if (r != CUDA_SUCCESS) return 1;
    int cases[4][4] = {{0, 0, 0, 0}, {0, -3, -1, 1}, {0, 9, 2, 1}, {0, 1, 1, 0}};   // (wo0, h0, t0, n)
    std::vector<uint16_t> out(128 * 32);
    int bad_total = 0;
    for (auto& cs : cases) {
        probe<<<1, 128>>>(map, cs[0], cs[1], cs[2], cs[3], dout);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("kernel: %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(out.data(), dout, 128 * 64, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int row = 0; row < 128; ++row) {
            const int j = row / 16, wo = cs[0] + row % 16, h = cs[1] + 2 * j, t = cs[2], n = cs[3];
            for (int k = 0; k < 32; ++k) {
                const uint32_t logical = row * 64 + k * 2;
                const uint32_t phys = logical ^ (((logical >> 7) & 3) << 4);
                __nv_bfloat16 v;  *reinterpret_cast<uint16_t*>(&v) = out[phys / 2];
                const bool inb = wo >= 0 && wo < WO && h >= 0 && h < H && t >= 0 && t < T;
                const float want = inb ? code(n, t, h, 2 * wo + k / 4, k % 4) : 0.f;
                if (__bfloat162float(v) != want) { if (bad < 5) printf("  case(%d,%d,%d,%d) row %d k %d got %g want %g\n", cs[0], cs[1], cs[2], cs[3], row, k, __bfloat162float(v), want); ++bad; }
            }
        }
        printf("case (wo0=%d h0=%d t=%d n=%d): %d mismatches\n", cs[0], cs[1], cs[2], cs[3], bad);
        bad_total += bad;
    }
    printf(bad_total ? "PROBE FAILED\n" : "PROBE OK\n");
    return bad_total != 0;
}
