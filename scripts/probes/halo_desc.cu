// Probe: can a K-major SWIZZLE_128B UMMA operand START INSIDE a 1024-byte swizzle atom (start address = atom base + q * 128 B)?
// That is what a halo tile needs: one strip of pixels [P][64 ch] loaded once by TMA, the 3x3 taps addressed as the same strip
// shifted by (dh * pitch + dw) pixels.  Two encodings of the descriptor are tried for every shift q0:
//   variant 0: base_offset field (bits 49-51) = 0              (works if the hardware swizzles on absolute smem address bits)
//   variant 1: base_offset = (start address >> 7) & 7         (the PTX ISA's rule for a start that is not pattern-aligned)
//   nvcc -gencode arch=compute_100a,code=sm_100a -o halo_desc halo_desc.cu -lcuda && ./halo_desc
#include <stdio.h>
#include <vector>
#include "../../avid_cma_b200/csrc/tc_common.cuh"

using namespace avid::tc;

constexpr int kStrip = 256;        // pixels in the strip (rows of 128 bytes)

__global__ void __launch_bounds__(128, 1)
probe(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, int q0, int variant, float* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* strip = smem;                              // [256][64] bf16, SW128
    uint8_t* wt = smem + kStrip * 128;                  // [64][64] bf16, SW128
    uint64_t* bar = reinterpret_cast<uint64_t*>(wt + 64 * 128);
    uint64_t* done = bar + 1;
    uint32_t* slot = reinterpret_cast<uint32_t*>(done + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_init(done, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(slot, 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, kStrip * 128 + 64 * 128);
        tma_load_2d(strip, &map_x, bar, 0, 0);
        tma_load_2d(wt, &map_w, bar, 0, 0);
        mbar_wait(bar, 0);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(strip) + (uint32_t)q0 * 128u;
        uint64_t da = make_smem_desc_sw128(a_addr, 16, 1024);
        if (variant == 1) da |= (uint64_t)((a_addr >> 7) & 7u) << 49;
        const uint64_t db = make_smem_desc_sw128(smem_u32(wt), 16, 1024);
        constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
        for (int k = 0; k < 4; ++k) umma_bf16(tmem, da + 2 * k, db + 2 * k, idesc, k != 0);
        umma_commit(done);
    }
    mbar_wait(done, 0);
    tc_fence_after();
    uint32_t r[32];
    for (int j = 0; j < 2; ++j) {
        tmem_ld_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16) + j * 32, r);
        tmem_ld_wait();
        for (int c = 0; c < 32; ++c) out[(warp * 32 + lane) * 64 + j * 32 + c] = __uint_as_float(r[c]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 64);
}

int main() {
    std::vector<__nv_bfloat16> x((size_t)kStrip * 64), w(64 * 64);
    auto xv = [](int p, int c) { return (float)((p * 5 + c * 3) % 7 - 3); };
    auto wv = [](int n, int c) { return (float)((n * 3 + c) % 5 - 2); };
    for (int p = 0; p < kStrip; ++p) for (int c = 0; c < 64; ++c) x[p * 64 + c] = __float2bfloat16(xv(p, c));
    for (int n = 0; n < 64; ++n) for (int c = 0; c < 64; ++c) w[n * 64 + c] = __float2bfloat16(wv(n, c));
    __nv_bfloat16 *dx, *dw;  float* dout;
    cudaMalloc(&dx, x.size() * 2);  cudaMalloc(&dw, w.size() * 2);  cudaMalloc(&dout, 128 * 64 * 4);
    cudaMemcpy(dx, x.data(), x.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dw, w.data(), w.size() * 2, cudaMemcpyHostToDevice);
    void* fn = nullptr;  cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    avid::TensorMapApi::EncodeTiled enc = (avid::TensorMapApi::EncodeTiled)fn;
    CUtensorMap mx, mw;
    {
        cuuint64_t dims[2] = {64, kStrip};  cuuint64_t strides[1] = {128};  cuuint32_t box[2] = {64, kStrip};  cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&mx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dx, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        cuuint64_t dims2[2] = {64, 64};  cuuint32_t box2[2] = {64, 64};
        CUresult r2 = enc(&mw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dw, dims2, strides, box2, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode rc=%d %d\n", (int)r, (int)r2);
        if (r || r2) return 1;
    }
    const int smem = kStrip * 128 + 64 * 128 + 1024 + 256;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    std::vector<float> out(128 * 64);
    const int shifts[] = {0, 8, 1, 2, 3, 5, 7, 9, 57, 59, 117};
    int ok_variant[2] = {1, 1};
    for (int variant = 0; variant < 2; ++variant)
        for (int q0 : shifts) {
            probe<<<1, 128, smem>>>(mx, mw, q0, variant, dout);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("kernel: %s\n", cudaGetErrorString(e)); return 1; }
            cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < 64; ++n) {
                    float want = 0.f;
                    for (int c = 0; c < 64; ++c) want += xv(q0 + m, c) * wv(n, c);
                    if (out[m * 64 + n] != want) ++bad;
                }
            printf("variant %d (base_offset %s) shift %3d: %d mismatches\n", variant, variant ? "(addr>>7)&7" : "0", q0, bad);
            if (bad) ok_variant[variant] = 0;
        }
    printf("HALO_DESC variant0 %s variant1 %s\n", ok_variant[0] ? "OK" : "FAIL", ok_variant[1] ? "OK" : "FAIL");
    return 0;
}
