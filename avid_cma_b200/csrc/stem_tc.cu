// tcgen05 kernels for the two stem convolutions (video 3x7x7 / audio 7x7, stride 2 in h and w, Cin = 3 / 1): the layers
// whose K dimension (kt*kh*kw*Cin = 441 / 49) has no 64-channel blocks for the generic im2col kernel (conv_tc.cu).
//
// Operand A without an im2col buffer ("Toeplitz view"): the clip is stored channels-last with Cin padded to 4 and the row
// padded in w ([n][t][h][wp][4] bf16, avid_stem_pack).  One output pixel's 7 horizontal taps x 4 channels are then 28 of 32
// CONTIGUOUS elements starting at padded pixel 2*wo, and consecutive output pixels start 2 pixels = 16 bytes apart -- so a
// tiled tensor map whose dim-1 stride (16 B) is smaller than its dim-0 extent (64 B) delivers A[pixel][k = (kw, c)] rows
// directly (TMA only requires 16-byte-multiple strides; verified on B200 by scripts/probes/toeplitz_tma.cu).  The vertical
// stride 2 is the map's elementStride, vertical / temporal padding is TMA zero fill.
//
// Row reuse: a tile is an 8 (ho) x 16 (wo) output patch.  For one temporal tap and one row parity the kernel loads 11
// input rows ONCE ([row][wo][32] = 1 KB per row, 64-byte swizzle); the A operand of vertical tap kh is the 128-row window
// starting at row kh/2 of the slot of parity kh%2 -- 4 (or 3) taps share each load.  GEMM K per (kt, kh) "chunk" is 32.
//
// forward : D[128 pixels][64 co] += A_chunk[128][32] * W_chunk[64 co][32]^T; all filter chunks stay resident in shared
//           memory, persistent CTAs, accumulator double-buffered in TMEM so the epilogue of tile i overlaps tile i+1.
// wgrad   : dW[(kh-group, kw, c)][co] += A^T dZ with both operands MN-major (the pixel index is the smem row index); the 4
//           vertical taps of one parity form the 128 accumulator rows (atoms 1 KB = one input row apart), one accumulator
//           per (kt, parity) lives in TMEM for the whole kernel; pixels are split over persistent CTAs, fp32 atomics at the end.
#include "common.cuh"
#include "tc_common.cuh"

namespace avid {
using namespace tc;

constexpr int kStemThreads = 192;            // warp 0: TMA, warp 1: TMEM alloc + MMA issue, warps 2-5: epilogue
constexpr int kStemCo = 64;
constexpr int kTileH = 8, kTileW = 16;       // output patch of one tile: 128 pixels = UMMA M (forward) / one k-block (wgrad)
constexpr int kRowBytes = kTileW * 64;       // one input row of the Toeplitz view: 16 wo x 32 elements x 2 B
constexpr int kSlotRows = 11;                // rows per (kt, parity, plane) unit: tap offsets kh/2 in 0..3 plus 8 output rows
constexpr int kSlotBytes = kSlotRows * kRowBytes;
constexpr int kChunkBBytes = kStemCo * 64;   // [64 co][32 k] bf16
constexpr int kDzBytes = 128 * 128;          // [128 pixels][64 co] bf16
constexpr int kMaxStages = 14;
constexpr int kStemStgBytes = 4 * 2 * 64 * 4;      // per-CTA reduction of the BatchNorm sums: [4 lane quarters][2 sums][64 channels]
constexpr int kSmemLimit = 232448 - 1024;    // 227 KB minus alignment slack

struct StemParams {
    int n, to, ho, wo;      // output extents
    int kt, kh, kw, ci;     // taps; ci = real input channels (<= 4)
    int st, pt, ph;         // temporal stride / padding, vertical padding (vertical and horizontal strides are 2)
    int tiles_h, tiles_w, num_tiles;
    int x3, stages;
};

__device__ __forceinline__ void tile_coords(const StemParams& p, int tile, int& n_i, int& t_o, int& h0, int& w0) {
    const int tw = tile % p.tiles_w;  tile /= p.tiles_w;
    const int th = tile % p.tiles_h;  tile /= p.tiles_h;
    t_o = tile % p.to;
    n_i = tile / p.to;
    h0 = th * kTileH;
    w0 = tw * kTileW;
}

// KH > 0: the vertical tap count is a compile-time constant (7 for both towers' stems) so the MMA issue loop unrolls into
// straight-line code -- with N = 64 an MMA is only 32 tensor-pipe cycles, and a single thread issuing through runtime loops
// and descriptor re-encoding (~100 cycles per MMA, measured) was the bottleneck of this kernel.  KH == 0: generic path.
template <int KH, bool X3>
__global__ void __launch_bounds__(kStemThreads, 1)
stem_forward_kernel(const __grid_constant__ CUtensorMap map_x_hi, const __grid_constant__ CUtensorMap map_x_lo,
                    const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo, const StemParams p,
                    float* __restrict__ out, double* __restrict__ stats) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment by pointer arithmetic on the __shared__ array (a pointer -> integer -> pointer round trip would make
    // every later access a generic LD / ST instead of LDS / STS)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int chunks = p.kt * p.kh, planes = p.x3 ? 2 : 1;
    uint8_t* w_smem = smem;                                          // [chunk][plane][64 co][32 k]: hi | lo of a chunk = one 128-row operand
    uint8_t* ring = smem + planes * chunks * kChunkBBytes;           // [stage][11 rows][16 wo][32 k]
    float* s_stage = reinterpret_cast<float*>(ring + p.stages * kSlotBytes);       // [4 lane quarters][2 sums][64]: BatchNorm sums of the CTA
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + p.stages * kSlotBytes + kStemStgBytes);
    uint64_t* empty_bar = full_bar + kMaxStages;
    uint64_t* w_bar = empty_bar + kMaxStages;
    uint64_t* tmem_full = w_bar + 1;      // [2]
    uint64_t* tmem_empty = tmem_full + 2; // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;      // provably warp-uniform
    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&map_x_hi);
        prefetch_tensormap(&map_w_hi);
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(w_bar, 1);
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tmem_full[b], 1);
            mbar_init(&tmem_empty[b], 4);     // one arrive per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, X3 ? 256 : 128);     // two accumulators of 64 (bf16x3: 128 = hi*hi | hi*lo) columns
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    // producer and MMA issuer: whole warps walk the loops, one elected lane issues (uniform-register operands, no per-instruction
    // ELECT / R2UR waterfall -- see conv_tc_kernel)
    if (warp == 0) {
        // ===== TMA producer: the filter once, then (kt, parity, plane) row slots of every tile =====
        if (elect_one()) {
            mbar_expect_tx(w_bar, (uint32_t)(planes * chunks * kChunkBBytes));
            for (int pl = 0; pl < planes; ++pl)
                for (int c = 0; c < chunks; ++c)
                    tma_load_2d(w_smem + (c * planes + pl) * kChunkBBytes, pl ? &map_w_lo : &map_w_hi, w_bar, c * 32, 0);
        }
        __syncwarp();
        int stage = 0, phase = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            int n_i, t_o, h0, w0;
            tile_coords(p, tile, n_i, t_o, h0, w0);
            for (int a = 0; a < p.kt; ++a)
                for (int par = 0; par < 2 && par < p.kh; ++par)
                    for (int pl = 0; pl < planes; ++pl) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        if (elect_one()) {
                            mbar_expect_tx(&full_bar[stage], kSlotRows * kRowBytes);
                            tma_load_5d(ring + stage * kSlotBytes, pl ? &map_x_lo : &map_x_hi, &full_bar[stage], 0, w0, 2 * h0 - p.ph + par,
                                        t_o * p.st - p.pt + a, n_i);
                        }
                        __syncwarp();
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        constexpr uint32_t idesc = make_idesc_bf16(128, kStemCo, 0, 0);
        // bf16x3: x_hi x [w_hi ; w_lo] is ONE MMA of N = 128 (the planes of a chunk are adjacent in shared memory), its column
        // halves are summed in the epilogue; x_lo x w_hi adds into the first half: 14 KB instead of 18 KB of operand fetch per k-step
        constexpr uint32_t idesc2 = make_idesc_bf16(128, 2 * kStemCo, 0, 0);
        constexpr int kPl = X3 ? 2 : 1;
        mbar_wait(w_bar, 0);
        // descriptors differ only in the 14-bit start-address field (bytes >> 4): encode once, then add offsets
        const uint64_t desc0 = make_smem_desc_sw64(0, 16, 512);
        const uint64_t w_desc = desc0 + (smem_u32(w_smem) >> 4);
        const uint64_t ring_desc = desc0 + (smem_u32(ring) >> 4);
        const int kh_n = KH > 0 ? KH : p.kh;
        int stage = 0, phase = 0, it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            mbar_wait(&tmem_empty[buf], ((it >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t acc = tmem_base + buf * (kPl * kStemCo);
            uint32_t accumulate = 0;
            for (int a = 0; a < p.kt; ++a) {
                const uint64_t w_a = w_desc + (uint32_t)((a * kh_n * kPl * kChunkBBytes) >> 4);
#pragma unroll
                for (int par = 0; par < 2; ++par) {
                    if (par >= kh_n) break;
#pragma unroll
                    for (int pl = 0; pl < (X3 ? 2 : 1); ++pl) {
                        mbar_wait(&full_bar[stage], phase);
                        tc_fence_after();
                        const uint64_t slot = ring_desc + (uint32_t)((stage * kSlotBytes) >> 4);
                        if (elect_one()) {
#pragma unroll
                            for (int kh = par; kh < (KH > 0 ? KH : 8); kh += 2) {
                                if (KH == 0 && kh >= kh_n) break;
#pragma unroll
                                for (int ks = 0; ks < 2; ++ks) {
                                    const uint64_t da = slot + (uint32_t)(((kh >> 1) * kRowBytes + ks * 32) >> 4);
                                    const uint64_t db_hi = w_a + (uint32_t)((kh * kPl * kChunkBBytes + ks * 32) >> 4);
                                    if (pl == 0) umma_bf16(acc, da, db_hi, X3 ? idesc2 : idesc, accumulate | (uint32_t)(kh != par || ks != 0));
                                    else umma_bf16(acc, da, db_hi, idesc, 1);
                                }
                            }
                            umma_commit(&empty_bar[stage]);
                        }
                        __syncwarp();
                        if (pl == 0) accumulate = 1;
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
            if (elect_one()) umma_commit(&tmem_full[buf]);
            __syncwarp();
        }
    } else if (warp >= 2) {
        // ===== epilogue: TMEM -> registers -> per-warp shared-memory transpose -> coalesced global rows + BatchNorm statistics =====
        // tcgen05.ld hands every thread one output pixel (row of 64 channels).  Round 1 stored those rows straight from that layout:
        // every STG.128 of a warp touched 32 different 256-byte rows with 16 bytes each (512 partial-sector writes per warp and
        // tile), and ncu showed the epilogue warps stalled on the store queue for 40 % of all samples with the tensor pipe 46 %
        // active -- the stores, not the MMAs, paced the kernel.  Now each warp transposes 32 x 16 half-chunks through a padded
        // staging tile and writes 64-byte segments (4 lanes per row, 8 rows per instruction: full sectors, 4x fewer requests), like
        // conv_tc_kernel; the per-channel sums stay in registers across the CTA's tiles and leave once, as fp64 atomics.
        // Round 2b: the transpose through shared memory is gone as well -- the kernel is bound by shared-memory bandwidth (ncu: MMA operand
        // reads 62 % + epilogue LSU traffic 23 % + TMA writes of the data pipe), and tcgen05.ld.16x256b already hands the 4 lanes of a row
        // 32 contiguous bytes (one full sector) per 8-column group: 8 rows per store instruction as before, no staging tile.
        const int q = warp & 3;
        const int r8 = lane >> 2, c2 = (lane & 3) * 2;
        const int r = q * 32 + lane;
        float run1[16], run2[16];          // running sums of this lane's 16 columns: [32-column half][column group j][pair element]
#pragma unroll
        for (int t = 0; t < 16; ++t) run1[t] = run2[t] = 0.f;
        int it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            int n_i, t_o, h0, w0;
            tile_coords(p, tile, n_i, t_o, h0, w0);
            const int ho = h0 + r / kTileW, wo = w0 + r % kTileW;
            const uint32_t my_row = (ho < p.ho && wo < p.wo) ? (uint32_t)(((n_i * p.to + t_o) * p.ho + ho) * p.wo + wo) : ~0u;     // output pixel (< 2^31: host check)
            uint32_t rows4[4];                      // the rows this lane serves: r8, r8 + 8, r8 + 16, r8 + 24 of the quarter (~0: no row)
#pragma unroll
            for (int i = 0; i < 4; ++i) rows4[i] = __shfl_sync(0xffffffffu, my_row, i * 8 + r8);
            mbar_wait(&tmem_full[buf], (it >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + buf * ((X3 ? 2 : 1) * kStemCo) + ((uint32_t)(q * 32) << 16);
#pragma unroll
            for (int hc = 0; hc < 2; ++hc) {        // 32-column halves of the 64 output channels
                uint32_t v[2][16];                  // [16-row half][4 j + 2 (row + 8) + pair element]
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) tmem_ld_16x256b_x4(taddr + ((uint32_t)(hf * 16) << 16) + hc * 32, v[hf]);
                if (X3) {
                    uint32_t v2[2][16];
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) tmem_ld_16x256b_x4(taddr + ((uint32_t)(hf * 16) << 16) + kStemCo + hc * 32, v2[hf]);   // the hi*lo half
                    tmem_ld_wait();
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf)
#pragma unroll
                        for (int u = 0; u < 16; ++u) v[hf][u] = __float_as_uint(__uint_as_float(v[hf][u]) + __uint_as_float(v2[hf][u]));
                } else {
                    tmem_ld_wait();
                }
                if (hc == 1) {          // the accumulator is in registers: the next tile may start
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty[buf]);
                }
#pragma unroll
                for (int hf = 0; hf < 2; ++hf)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int rs = 0; rs < 2; ++rs) {
                            const uint32_t row = rows4[hf * 2 + rs];
                            if (row == ~0u) continue;
                            const float o0 = __uint_as_float(v[hf][4 * j + 2 * rs]), o1 = __uint_as_float(v[hf][4 * j + 2 * rs + 1]);
                            *reinterpret_cast<float2*>(out + (size_t)row * kStemCo + hc * 32 + 8 * j + c2) = make_float2(o0, o1);
                            if (stats) {
                                run1[hc * 8 + 2 * j] += o0;      run2[hc * 8 + 2 * j] = fmaf(o0, o0, run2[hc * 8 + 2 * j]);
                                run1[hc * 8 + 2 * j + 1] += o1;  run2[hc * 8 + 2 * j + 1] = fmaf(o1, o1, run2[hc * 8 + 2 * j + 1]);
                            }
                        }
            }
        }
        if (stats) {
            // lanes with equal (lane & 3) hold partial sums of the same columns: reduce over the 8 row groups of the warp, over the
            // 4 lane quarters through shared memory, then ONE fp64 atomic per column and CTA
            float* const s_red = s_stage;           // [4 quarters][2 sums][64]
            asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
            for (int t = 0; t < 16; ++t) {
                float a = run1[t], b = run2[t];
#pragma unroll
                for (int o = 4; o <= 16; o <<= 1) {         // over the 8 lanes (rows) that own the same columns
                    a += __shfl_xor_sync(0xffffffffu, a, o);
                    b += __shfl_xor_sync(0xffffffffu, b, o);
                }
                if (lane < 4) {
                    const int col = (t >> 3) * 32 + 8 * ((t & 7) >> 1) + c2 + (t & 1);
                    s_red[(q * 2 + 0) * kStemCo + col] = a;
                    s_red[(q * 2 + 1) * kStemCo + col] = b;
                }
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const int t = threadIdx.x - 64;      // 0..127 = (which, channel)
            const int which = t >> 6, ch = t & 63;
            atomicAdd(stats + which * kStemCo + ch, (double)s_red[(0 * 2 + which) * kStemCo + ch] + (double)s_red[(1 * 2 + which) * kStemCo + ch] +
                                                        (double)s_red[(2 * 2 + which) * kStemCo + ch] + (double)s_red[(3 * 2 + which) * kStemCo + ch]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, X3 ? 256 : 128);
}

// ---- filter gradient -------------------------------------------------------------------------------------------------
__device__ __forceinline__ void red_add_v4f(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <bool X3>
__global__ void __launch_bounds__(kStemThreads, 1)
stem_wgrad_kernel(const __grid_constant__ CUtensorMap map_x_hi, const __grid_constant__ CUtensorMap map_x_lo,
                  const __grid_constant__ CUtensorMap map_d_hi, const __grid_constant__ CUtensorMap map_d_lo, const StemParams p,
                  uint32_t tmem_cols, float* __restrict__ dfilt) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment by pointer arithmetic on the __shared__ array (a pointer -> integer -> pointer round trip would make
    // every later access a generic LD / ST instead of LDS / STS)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int planes = p.x3 ? 2 : 1;
    uint8_t* dz_smem = smem;                                   // [2 buffers][plane][128 pixels][64 co]
    uint8_t* ring = smem + 2 * planes * kDzBytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + p.stages * kSlotBytes);
    uint64_t* empty_bar = full_bar + kMaxStages;
    uint64_t* dz_full = empty_bar + kMaxStages;   // [2]
    uint64_t* dz_empty = dz_full + 2;             // [2]
    uint64_t* accum_bar = dz_empty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;      // provably warp-uniform
    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&map_x_hi);
        prefetch_tensormap(&map_d_hi);
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&dz_full[b], 1);
            mbar_init(&dz_empty[b], 1);
        }
        mbar_init(accum_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    // The temporal taps are split over the CTAs (CTA i works on tap i % kt of the tiles i / kt, i / kt + gridDim.x / kt, ...): a CTA
    // then needs only two accumulators (row parities), which leaves TMEM room for the wide bf16x3 form x_hi x [dZ_hi | dZ_lo]
    // (N = 128, the dZ planes are 16 KB apart = the descriptor's leading-dimension offset) + x_lo x dZ_hi: 14 KB instead of
    // 18 KB of operand fetch per k-step of this fetch-bound kernel.  The kt CTAs of a tile group walk the same tiles, so the
    // dZ tiles they share come from L2.
    const int a_only = blockIdx.x % p.kt, cta = blockIdx.x / p.kt, ctas = gridDim.x / p.kt;
    constexpr int kAccW = X3 ? 2 * kStemCo : kStemCo;      // accumulator columns per (tap, parity)
    const bool has_work = cta < p.num_tiles;

    if (warp == 0) {
        // ===== TMA producer (whole warp, one elected lane issues: see conv_tc_kernel) =====
        int stage = 0, phase = 0, it = 0;
        for (int tile = cta; tile < p.num_tiles; tile += ctas, ++it) {
            int n_i, t_o, h0, w0;
            tile_coords(p, tile, n_i, t_o, h0, w0);
            const int buf = it & 1;
            mbar_wait(&dz_empty[buf], ((it >> 1) & 1) ^ 1);
            if (elect_one()) {
                mbar_expect_tx(&dz_full[buf], (uint32_t)(planes * kDzBytes));
                for (int pl = 0; pl < planes; ++pl)
                    tma_load_5d(dz_smem + (buf * planes + pl) * kDzBytes, pl ? &map_d_lo : &map_d_hi, &dz_full[buf], 0, w0, h0, t_o, n_i);
            }
            __syncwarp();
            for (int a = a_only; a <= a_only; ++a)
                for (int par = 0; par < 2 && par < p.kh; ++par)
                    for (int pl = 0; pl < planes; ++pl) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        if (elect_one()) {
                            mbar_expect_tx(&full_bar[stage], kSlotRows * kRowBytes);
                            tma_load_5d(ring + stage * kSlotBytes, pl ? &map_x_lo : &map_x_hi, &full_bar[stage], 0, w0, 2 * h0 - p.ph + par,
                                        t_o * p.st - p.pt + a, n_i);
                        }
                        __syncwarp();
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: A = x slot (MN-major, 64-byte swizzle, 4 tap atoms one row apart), B = dZ (MN-major, 128-byte swizzle) =====
        constexpr uint32_t idesc = make_idesc_bf16(128, kStemCo, 1, 1);
        constexpr uint32_t idesc2 = make_idesc_bf16(128, 2 * kStemCo, 1, 1);
        const uint64_t ring_desc = make_smem_desc_sw64(0, kRowBytes, 512) + (smem_u32(ring) >> 4);
        const uint64_t dz_desc = make_smem_desc_sw128(0, X3 ? kDzBytes : 16, 1024) + (smem_u32(dz_smem) >> 4);
        int stage = 0, phase = 0, it = 0;
        for (int tile = cta; tile < p.num_tiles; tile += ctas, ++it) {
            const int buf = it & 1;
            mbar_wait(&dz_full[buf], (it >> 1) & 1);
            tc_fence_after();
            const uint64_t dz_hi = dz_desc + (uint32_t)((buf * planes * kDzBytes) >> 4);
            for (int a = a_only; a <= a_only; ++a)
                for (int par = 0; par < 2 && par < p.kh; ++par) {
                    const uint32_t acc = tmem_base + par * kAccW;
#pragma unroll
                    for (int pl = 0; pl < (X3 ? 2 : 1); ++pl) {
                        mbar_wait(&full_bar[stage], phase);
                        tc_fence_after();
                        const uint64_t slot = ring_desc + (uint32_t)((stage * kSlotBytes) >> 4);
                        if (elect_one()) {
#pragma unroll
                            for (int ks = 0; ks < 8; ++ks) {          // 16 pixels (one patch row) per MMA
                                const uint64_t da = slot + (uint32_t)((ks * kRowBytes) >> 4);
                                const uint64_t db_hi = dz_hi + (uint32_t)((ks * 2048) >> 4);
                                if (pl == 0) {
                                    umma_bf16(acc, da, db_hi, X3 ? idesc2 : idesc, (it | ks) != 0);      // x_hi x [dZ_hi | dZ_lo]
                                } else {
                                    umma_bf16(acc, da, db_hi, idesc, 1);
                                }
                            }
                            umma_commit(&empty_bar[stage]);
                        }
                        __syncwarp();
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    }
                }
            if (elect_one()) {
                umma_commit(&dz_empty[buf]);
                if (tile + ctas >= p.num_tiles) umma_commit(accum_bar);
            }
            __syncwarp();
        }
    } else if (warp >= 2 && has_work) {
        // ===== epilogue: accumulator row = (tap kh = parity + 2*(row/32), kw = (row%32)/4, c = row%4) -> atomics into dW[tap][c][co] =====
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int kwi = (r & 31) >> 2, c = r & 3;
        for (int a = a_only; a <= a_only; ++a)
            for (int par = 0; par < 2 && par < p.kh; ++par) {
                const int kh = par + 2 * (r >> 5);
                const bool valid = kh < p.kh && kwi < p.kw && c < p.ci;
                float* dst = dfilt + ((size_t)(((a * p.kh + kh) * p.kw + kwi) * 4 + c)) * kStemCo;
                const uint32_t taddr = tmem_base + par * kAccW + ((uint32_t)(q * 32) << 16);
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(taddr + j * 32, v);
                    if (X3) {
                        uint32_t v2[32];
                        tmem_ld_32x32b_x32(taddr + kStemCo + j * 32, v2);      // the x_hi x dZ_lo half
                        tmem_ld_wait();
#pragma unroll
                        for (int u = 0; u < 32; ++u) v[u] = __float_as_uint(__uint_as_float(v[u]) + __uint_as_float(v2[u]));
                    }
                    tmem_ld_wait();
                    if (valid) {
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            red_add_v4f(dst + j * 32 + 4 * u, __uint_as_float(v[4 * u]), __uint_as_float(v[4 * u + 1]), __uint_as_float(v[4 * u + 2]),
                                        __uint_as_float(v[4 * u + 3]));
                    }
                }
            }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

// ---- packing kernels -------------------------------------------------------------------------------------------------
// x [n][c][t*h][w] fp32 -> planes [n][t*h][wp][4] bf16 with the pixel w at padded position w + pad_left, zeros elsewhere
__global__ void __launch_bounds__(256) stem_pack_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                                        int n, int c, int64_t th, int w, int wp, int pad_left) {
    const int64_t total = (int64_t)n * th * wp;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int wq = (int)(i % wp);
        const int64_t row = i / wp;             // n * th + (t*h index)
        const int64_t img = row / th, s = row - img * th;
        const int wr = wq - pad_left;
        __nv_bfloat16 h[4], l[4];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            const float v = (ch < c && wr >= 0 && wr < w) ? __ldg(x + ((img * c + ch) * th + s) * w + wr) : 0.f;
            h[ch] = __float2bfloat16_rn(v);
            l[ch] = __float2bfloat16_rn(v - __bfloat162float(h[ch]));
        }
        reinterpret_cast<uint2*>(hi)[i] = *reinterpret_cast<uint2*>(h);
        if (lo) reinterpret_cast<uint2*>(lo)[i] = *reinterpret_cast<uint2*>(l);
    }
}

// PyTorch filter [co][ci][kt][kh][kw] fp32 -> planes [co][(kt*kh) chunks][32 = (kw padded to 8) x 4 channels] bf16
__global__ void stem_filter_pack_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int co, int ci,
                                        int chunks, int kw) {
    const int total = co * chunks * 32;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int e = i & 31, chunk = (i >> 5) % chunks, o = (i >> 5) / chunks;
        const int kwi = e >> 2, c = e & 3;
        const float v = (kwi < kw && c < ci) ? __ldg(w + (((size_t)o * ci + c) * chunks + chunk) * kw + kwi) : 0.f;
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        hi[i] = h;
        if (lo) lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
}

// ---- host ------------------------------------------------------------------------------------------------------------
static int encode_tiled_nd(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box,
                           const cuuint32_t* estr, CUtensorMapSwizzle swizzle) {
    const TensorMapApi& api = tensor_map_api();
    if (!api.ok) { set_error("stem_tc: cuTensorMapEncode* driver entry points unavailable"); return AVID_ECUDA; }
    CUresult r = api.tiled(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("stem_tc: cuTensorMapEncodeTiled failed (%d)", (int)r); return AVID_ECUDA; }
    return AVID_OK;
}

static int stem_check(const avid_conv_shape_t* s, int wp, StemParams* p) {
    AVID_REQUIRE(s, "stem_tc: NULL shape");
    AVID_REQUIRE(s->sh == 2 && s->sw == 2, "stem_tc: vertical and horizontal strides must be 2 (got %d, %d)", s->sh, s->sw);
    AVID_REQUIRE(s->kw >= 1 && s->kw <= 8 && s->kh >= 1 && s->kh <= 8 && s->kt >= 1 && s->kt <= 4, "stem_tc: filter %dx%dx%d not supported", s->kt, s->kh, s->kw);
    AVID_REQUIRE(s->ci >= 1 && s->ci <= 4 && s->co == kStemCo, "stem_tc: needs ci <= 4 and co == 64 (got %d, %d)", s->ci, s->co);
    AVID_REQUIRE(wp == 2 * s->wo + 8, "stem_tc: padded row length must be 2*wo + 8 = %d (got %d)", 2 * s->wo + 8, wp);
    AVID_REQUIRE(s->wi + s->pw <= wp && s->n > 0 && s->to > 0 && s->ho > 0 && s->wo > 0, "stem_tc: bad geometry");
    p->n = s->n;  p->to = s->to;  p->ho = s->ho;  p->wo = s->wo;
    p->kt = s->kt;  p->kh = s->kh;  p->kw = s->kw;  p->ci = s->ci;
    p->st = s->st;  p->pt = s->pt;  p->ph = s->ph;
    p->tiles_h = (s->ho + kTileH - 1) / kTileH;
    p->tiles_w = (s->wo + kTileW - 1) / kTileW;
    const int64_t tiles = (int64_t)s->n * s->to * p->tiles_h * p->tiles_w;
    AVID_REQUIRE(tiles < ((int64_t)1 << 31), "stem_tc: too many tiles");
    p->num_tiles = (int)tiles;
    return AVID_OK;
}

static int encode_stem_x(CUtensorMap* map, const void* base, const avid_conv_shape_t* s, int wp) {
    const cuuint64_t row = (cuuint64_t)wp * 8;   // bytes per padded input row (4 channels x bf16)
    cuuint64_t dims[5] = {32, (cuuint64_t)s->wo, (cuuint64_t)s->hi, (cuuint64_t)s->ti, (cuuint64_t)s->n};
    cuuint64_t strides[4] = {16, row, row * s->hi, row * s->hi * s->ti};
    cuuint32_t box[5] = {32, kTileW, 2 * kSlotRows, 1, 1};
    cuuint32_t estr[5] = {1, 1, 2, 1, 1};
    return encode_tiled_nd(map, base, 5, dims, strides, box, estr, CU_TENSOR_MAP_SWIZZLE_64B);
}

int stem_forward_run(const avid_conv_shape_t* s, const void* x_hi, const void* x_lo, int wp, const void* w_hi, const void* w_lo, float* out,
                     double* stats, cudaStream_t st) {
    StemParams p;
    int rc = stem_check(s, wp, &p);
    if (rc) return rc;
    AVID_REQUIRE(x_hi && w_hi && out && (x_lo == nullptr) == (w_lo == nullptr), "stem_forward_tc: NULL pointer / mismatched lo planes");
    AVID_REQUIRE((int64_t)s->n * s->to * s->ho * s->wo < ((int64_t)1 << 31), "stem_forward_tc: more than 2^31 output pixels");
    p.x3 = x_lo != nullptr;
    const int planes = p.x3 ? 2 : 1, chunks = p.kt * p.kh;
    const int w_bytes = planes * chunks * kChunkBBytes;
    p.stages = (kSmemLimit - 512 - kStemStgBytes - w_bytes) / kSlotBytes;
    if (p.stages > kMaxStages) p.stages = kMaxStages;
    AVID_REQUIRE(p.stages >= 2, "stem_forward_tc: filter does not fit in shared memory");
    const int smem = 1024 + w_bytes + p.stages * kSlotBytes + kStemStgBytes + 512;
    CUtensorMap mx[2], mw[2];
    if ((rc = encode_stem_x(&mx[0], x_hi, s, wp))) return rc;
    mx[1] = mx[0];
    if (p.x3 && (rc = encode_stem_x(&mx[1], x_lo, s, wp))) return rc;
    cuuint64_t wdims[2] = {(cuuint64_t)chunks * 32, kStemCo};
    cuuint64_t wstr[1] = {(cuuint64_t)chunks * 64};
    cuuint32_t wbox[2] = {32, kStemCo}, west[2] = {1, 1};
    if ((rc = encode_tiled_nd(&mw[0], w_hi, 2, wdims, wstr, wbox, west, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
    mw[1] = mw[0];
    if (p.x3 && (rc = encode_tiled_nd(&mw[1], w_lo, 2, wdims, wstr, wbox, west, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
    typedef void (*Kernel)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, StemParams, float*, double*);
    const Kernel kernel = p.kh == 7 ? (p.x3 ? (Kernel)stem_forward_kernel<7, true> : (Kernel)stem_forward_kernel<7, false>)
                                    : (p.x3 ? (Kernel)stem_forward_kernel<0, true> : (Kernel)stem_forward_kernel<0, false>);
    static bool configured = false;
    if (!configured) {
        for (Kernel k : {(Kernel)stem_forward_kernel<7, true>, (Kernel)stem_forward_kernel<7, false>, (Kernel)stem_forward_kernel<0, true>,
                         (Kernel)stem_forward_kernel<0, false>}) {
            cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
            if (e != cudaSuccess) { set_error("stem_forward_tc: smem attribute: %s", cudaGetErrorString(e)); return AVID_ECUDA; }
        }
        configured = true;
    }
    const int grid = p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs;
    kernel<<<grid, kStemThreads, smem, st>>>(mx[0], mx[1], mw[0], mw[1], p, out, stats);
    return check_launch("stem_forward_kernel");
}

int stem_wgrad_run(const avid_conv_shape_t* s, const void* x_hi, const void* x_lo, int wp, const void* d_hi, const void* d_lo, float* dfilt,
                   cudaStream_t st) {
    StemParams p;
    int rc = stem_check(s, wp, &p);
    if (rc) return rc;
    AVID_REQUIRE(x_hi && d_hi && dfilt && (x_lo == nullptr) == (d_lo == nullptr), "stem_wgrad_tc: NULL pointer / mismatched lo planes");
    p.x3 = x_lo != nullptr;
    const int planes = p.x3 ? 2 : 1;
    p.stages = (kSmemLimit - 512 - 2 * planes * kDzBytes) / kSlotBytes;
    if (p.stages > kMaxStages) p.stages = kMaxStages;
    const int smem = 1024 + 2 * planes * kDzBytes + p.stages * kSlotBytes + 512;
    uint32_t cols = 32;
    while ((int)cols < 2 * (p.x3 ? 2 : 1) * kStemCo) cols <<= 1;       // two row-parity accumulators per CTA (the kt taps are split over CTAs)
    CUtensorMap mx[2], md[2];
    if ((rc = encode_stem_x(&mx[0], x_hi, s, wp))) return rc;
    mx[1] = mx[0];
    if (p.x3 && (rc = encode_stem_x(&mx[1], x_lo, s, wp))) return rc;
    cuuint64_t ddims[5] = {kStemCo, (cuuint64_t)s->wo, (cuuint64_t)s->ho, (cuuint64_t)s->to, (cuuint64_t)s->n};
    cuuint64_t dstr[4] = {128, (cuuint64_t)s->wo * 128, (cuuint64_t)s->ho * s->wo * 128, (cuuint64_t)s->to * s->ho * s->wo * 128};
    cuuint32_t dbox[5] = {kStemCo, kTileW, kTileH, 1, 1}, dest[5] = {1, 1, 1, 1, 1};
    if ((rc = encode_tiled_nd(&md[0], d_hi, 5, ddims, dstr, dbox, dest, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    md[1] = md[0];
    if (p.x3 && (rc = encode_tiled_nd(&md[1], d_lo, 5, ddims, dstr, dbox, dest, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(stem_wgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(stem_wgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
        if (e != cudaSuccess) { set_error("stem_wgrad_tc: smem attribute: %s", cudaGetErrorString(e)); return AVID_ECUDA; }
        configured = true;
    }
    const int groups = kNumSMs / p.kt;                    // CTAs per temporal tap
    const int grid = (p.num_tiles < groups ? p.num_tiles : groups) * p.kt;
    if (p.x3) stem_wgrad_kernel<true><<<grid, kStemThreads, smem, st>>>(mx[0], mx[1], md[0], md[1], p, cols, dfilt);
    else stem_wgrad_kernel<false><<<grid, kStemThreads, smem, st>>>(mx[0], mx[1], md[0], md[1], p, cols, dfilt);
    return check_launch("stem_wgrad_kernel");
}

}  // namespace avid

using namespace avid;

extern "C" {

int avid_stem_pack(const float* x, void* hi, void* lo, int32_t n, int32_t c, int32_t t, int32_t h, int32_t w, int32_t wp, int32_t pad_left,
                   void* stream) {
    AVID_REQUIRE(x && hi && n > 0 && c >= 1 && c <= 4 && t > 0 && h > 0 && w > 0 && wp >= w + pad_left && pad_left >= 0 && wp % 2 == 0,
                 "stem_pack: bad arguments");
    const int64_t total = (int64_t)n * t * h * wp;
    int64_t blocks = (total + 255) / 256;
    if (blocks > 16 * kNumSMs) blocks = 16 * kNumSMs;
    stem_pack_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, static_cast<__nv_bfloat16*>(hi), static_cast<__nv_bfloat16*>(lo), n, c,
                                                                                      (int64_t)t * h, w, wp, pad_left);
    return check_launch("stem_pack_kernel");
}

int avid_stem_filter_pack(const float* w_oihw, void* hi, void* lo, int32_t co, int32_t ci, int32_t kt, int32_t kh, int32_t kw, void* stream) {
    AVID_REQUIRE(w_oihw && hi && co > 0 && ci >= 1 && ci <= 4 && kt >= 1 && kh >= 1 && kw >= 1 && kw <= 8, "stem_filter_pack: bad arguments");
    const int total = co * kt * kh * 32;
    stem_filter_pack_kernel<<<(total + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(w_oihw, static_cast<__nv_bfloat16*>(hi),
                                                                                               static_cast<__nv_bfloat16*>(lo), co, ci, kt * kh, kw);
    return check_launch("stem_filter_pack_kernel");
}

int avid_stem_forward_tc(const avid_conv_shape_t* s, const void* x_hi, const void* x_lo, int32_t wp, const void* filt_hi, const void* filt_lo,
                         float* out, double* bn_stats, void* stream) {
    return stem_forward_run(s, x_hi, x_lo, wp, filt_hi, filt_lo, out, bn_stats, static_cast<cudaStream_t>(stream));
}

int avid_stem_wgrad_tc(const avid_conv_shape_t* s, const void* x_hi, const void* x_lo, int32_t wp, const void* dout_hi, const void* dout_lo,
                       float* dfilt, void* stream) {
    return stem_wgrad_run(s, x_hi, x_lo, wp, dout_hi, dout_lo, dfilt, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
