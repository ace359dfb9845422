// tcgen05 implicit-GEMM convolution for the encoders (include/avid_b200.h, avid_conv_*_tc).
//
// forward / stride-1 input gradient:  D[m, n] = sum_{tap, c} A[pix(m, tap), c] * B[tap][n][c]
//   * no im2col buffer: the A tile (128 output pixels x 64 channels of one filter tap) is fetched by ONE
//     TMA im2col-mode load straight from the channels-last bf16 activation tensor; zero padding, conv
//     stride and the wrap from one row / frame / clip to the next are done by the TMA unit.
//   * B tile (BN output channels x 64 input channels of the tap, K-major) by a tiled TMA load.
//   * both land in 128-byte-swizzled shared memory and feed tcgen05.mma (M=128, N=BN, K=16, bf16 in,
//     fp32 accumulate in TMEM) issued by a single thread; an mbarrier ring (TMA -> MMA -> TMA) of
//     kStages stages keeps the tensor pipe fed.
//   * "bf16x3" mode: operands are split as x = hi + lo (two bf16 planes); the kernel accumulates
//     hi*hi + hi*lo + lo*hi, i.e. ~16 significand bits per operand -- this is the mode that meets the
//     1e-3 parity bar through 40+ train-mode BatchNorms (SURVEY.md §7 "Precision").  "bf16" mode
//     accumulates hi*hi only.
//   * epilogue: 4 warps read the accumulator with tcgen05.ld (one output pixel per thread), add the
//     optional residual and store fp32 rows.
//
// filter gradient: D[ci, co] = sum_m X[pix(m, tap), ci] * dZ[m, co] per tap, reduction over pixels:
//   both operands are MN-major (the reduction index is the row index of the channels-last tensors);
//   X tiles by im2col TMA, dZ tiles by tiled TMA, split over pixel ranges, fp32 atomics into dW.
#include <mutex>
#include <stdlib.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace avid {

const TensorMapApi& tensor_map_api() {
    static TensorMapApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && fn)
            api.tiled = reinterpret_cast<TensorMapApi::EncodeTiled>(fn);
        fn = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &q) == cudaSuccess && fn)
            api.im2col = reinterpret_cast<TensorMapApi::EncodeIm2col>(fn);
        api.ok = api.tiled && api.im2col;
    });
    return api;
}

using namespace tc;

// conv_pair.cu: 64 -> 64 channel 1x3x3 stride-1 layers on CTA pairs (halo tile + resident filter); returns AVID_EUNSUPPORTED when the
// geometry is not its case
bool conv_pair_supported(const avid_conv_shape_t* s);
int conv_pair_run(const avid_conv_shape_t* s, int dgrad, const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, const float* addend,
                  float* out, double* stats, const BnBwdFuse& fuse, cudaStream_t st);

constexpr int kBM = 128;        // output pixels per CTA (UMMA M)
constexpr int kBK = 64;         // channels per k-block: 64 bf16 = one 128-byte swizzle row
constexpr int kTcThreads = 192; // warp 0: TMA producer, warp 1: TMEM alloc + MMA issuer, warps 2-5: epilogue

constexpr int kMaxTaps = 32;
constexpr int kStgLd = 20;      // row pitch (floats) of the 32 x 16 epilogue staging tiles: 16-byte aligned, conflict-free for float4 rows
constexpr int kConvThreads = 320; // conv_tc_kernel: warp 0 TMA producer, warp 1 TMEM alloc + MMA issuer, warps 2-9 epilogue

constexpr int kMaxCls = 8;      // stride-parity classes of one launch (2 x 2 x 2)

// One stride-parity class of a strided input gradient (the only class of every other launch): which destination pixels it
// enumerates, where their taps read the source, which filter taps it uses.
struct TcClass {
    int M;                          // destination pixels of the class, enumerated ((n * tq + t) * hq + h) * wq + w
    int tq, hq, wq;                 // extents of that enumeration
    FastDiv dtq, dhq, dwq;
    int bt, bh, bw;                 // source coordinate read by tap offset 0 for destination pixel 0 (forward: -padding)
    int tap0, ntaps;                // its entries of TcConvParams::taps
    int rt, rh, rw;                 // residues: enumerated pixel (n, t, h, w) -> ((n * Td + t * ot + rt) * Hd + h * oh + rh) * Wd + w * ow + rw
};

struct TcConvParams {
    int ncls;
    TcClass cls[kMaxCls];
    int mtiles;                     // 128-pixel tiles of the largest class
    FastDiv dncls, dnblocks;        // divisions of the tile index
    FastDiv dadd_t, dadd_h, dadd_w;
    int cd;                         // destination channels (GEMM N total)
    int cs;                         // source channels (GEMM K per tap)
    int st, sh, sw;                 // source traversal strides (forward conv stride; 1 for dgrad)
    uint32_t taps[kMaxTaps];        // off_w | off_h << 8 | off_t << 16 | filter tap << 24 (offsets relative to bt/bh/bw of the class)
    int x3;                         // 1: hi/lo planes, 3 MMAs per k-step; 0: hi only
    int strided_out;                // destination rows are scattered to the pixels of their class
    int Td, Hd, Wd, ot, oh, ow;
    // subsampled addend (add_w != 0): the addend tensor is [n, add_T, add_H, add_W, cd] and belongs to the destination pixels whose
    // coordinates are multiples of (add_t, add_h, add_w) -- the input gradient of a strided 1x1x1 residual convolution, zero elsewhere
    int add_t, add_h, add_w, add_T, add_H, add_W;
    int debug;                      // AVID_TC_DEBUG probe bits (scripts/probe_conv.py; 0 in production): 1 no global stores, 2 no
                                    // statistics, 4 epilogue only hands the accumulator back, 8 no MMAs, 64 no per-tile atomics
};

struct alignas(64) TcConvMaps {
    CUtensorMap a[kMaxCls][2];      // im2col maps of the source planes (hi, lo) per class: the box corners depend on the class
    CUtensorMap b[2];               // filter planes
};

// tile -> (class, 128-pixel tile of the class, channel block); the channel block is the fastest index (all tiles of a CTA share it),
// then the class: the classes of a strided input gradient read the SAME source pixels through different taps, so tiles that run
// at the same time share them in L2 (one launch per class read the source once per class: 4 x 205 MB for the conv3x entry)
struct TcTile {
    int cls, m0, n0;
    bool valid;
};
template <int BN>
__device__ __forceinline__ TcTile decode_tile(const TcConvParams& p, int tile, int nblocks) {
    TcTile t;
    int nb;
    const int rest = fdivmod(tile, p.dnblocks, nb);
    t.n0 = nb * BN;
    t.m0 = fdivmod(rest, p.dncls, t.cls) * kBM;
    t.valid = t.m0 < p.cls[t.cls].M;
    (void)nblocks;
    return t;
}

template <int BN>
struct TcSmem {
    static constexpr int kABytes = kBM * kBK * 2;   // 16 KB
    static constexpr int kBBytes = BN * kBK * 2;
    static constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;
    static constexpr int kStages = BN <= 64 ? 4 : 3;
    static constexpr int kStatBytes = BN * 16;                  // [BN] float4 BatchNorm constants of a fused BatchNorm backward
    static constexpr int kStgBytes = 8 * 32 * kStgLd * 4;       // [8 epilogue warps][32 rows][kStgLd] floats: transpose staging
    static constexpr int kBytes = kStages * kStageBytes + kStatBytes + kStgBytes + 1024 /*alignment slack*/ + 256 /*barriers*/;
};

// Persistent: CTA i works on tiles i, i + gridDim.x, ...; a tile is (128 destination pixels) x (BN destination channels), tiles
// of the same pixels are adjacent so concurrently running CTAs share the A tile in L2.  The accumulator is double-buffered in
// TMEM (2 x BN columns): the epilogue of tile k (TMEM -> registers -> global, BatchNorm statistics) overlaps the TMA / MMA of
// tile k + 1.
template <int BN>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_tc_kernel(const __grid_constant__ TcConvMaps maps, const __grid_constant__ TcConvParams p,
               const float* __restrict__ addend, float* __restrict__ out, double* __restrict__ stats, const BnBwdFuse fuse) {
    using S = TcSmem<BN>;
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment by pointer arithmetic on the __shared__ array (a pointer -> integer -> pointer round trip would make
    // every later access a generic LD / ST instead of LDS / STS)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float4* s_par = reinterpret_cast<float4*>(smem + S::kStages * S::kStageBytes);
    float* s_stage = reinterpret_cast<float*>(smem + S::kStages * S::kStageBytes + S::kStatBytes);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kStages * S::kStageBytes + S::kStatBytes + S::kStgBytes);
    uint64_t* empty_bar = full_bar + S::kStages;
    uint64_t* tmem_full = empty_bar + S::kStages;     // [2]
    uint64_t* tmem_empty = tmem_full + 2;             // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;      // provably warp-uniform
    const int cblocks = p.cs / kBK;
    const int nblocks = p.cd / BN;
    const int num_tiles = p.mtiles * p.ncls * nblocks;

    if (warp == 0 && lane == 0) {
        for (int c = 0; c < p.ncls; ++c) {
            prefetch_tensormap(&maps.a[c][0]);
            if (p.x3) prefetch_tensormap(&maps.a[c][1]);
        }
        prefetch_tensormap(&maps.b[0]);
        if (p.x3) prefetch_tensormap(&maps.b[1]);
        for (int s = 0; s < S::kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tmem_full[b], 1);
            mbar_init(&tmem_empty[b], 8);      // one arrive per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 4 * BN);      // two accumulator buffers x (hi | lo filter-plane halves)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    // The producer and the MMA issuer are WHOLE warps walking their loops (waits included) with one elected lane issuing: the operands
    // of UTMALDG / UTCHMMA are then provably warp-uniform and live in uniform registers.  Issued from `if (lane == 0)` ptxas wraps
    // every such instruction in an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall (it cannot prove a single active thread), which made
    // the issuing thread -- not the tensor pipe -- the pace of the short-K layers (scripts/probes/issue_rate.cu, probe_pair.py).
    if (warp == 0) {
        // ===== TMA producer =====
        const uint32_t tx = (uint32_t)(S::kABytes + S::kBBytes) * (p.x3 ? 2u : 1u);
        int stage = 0, phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const TcTile tl = decode_tile<BN>(p, tile, nblocks);
            if (!tl.valid) continue;
            const TcClass& cl = p.cls[tl.cls];
            const CUtensorMap* const map_a_hi = &maps.a[tl.cls][0];
            const CUtensorMap* const map_a_lo = &maps.a[tl.cls][1];
            const int n0 = tl.n0;
            int w_o, h_o, t_o;
            const int n_i = fdivmod(fdivmod(fdivmod(tl.m0, cl.dwq, w_o), cl.dhq, h_o), cl.dtq, t_o);
            const int bw = w_o * p.sw + cl.bw, bh = h_o * p.sh + cl.bh, bt = t_o * p.st + cl.bt;   // source pixel read by tap offset 0
            const int nkb = cl.ntaps * cblocks;
            for (int kb = 0; kb < nkb; ++kb) {
                const int tap = kb / cblocks, c0 = (kb - tap * cblocks) * kBK;
                const uint32_t tp = p.taps[cl.tap0 + tap];
                const uint16_t c = tp & 0xFF, b = (tp >> 8) & 0xFF, a = (tp >> 16) & 0xFF;
                const int ftap = tp >> 24;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* st = smem + stage * S::kStageBytes;
                if (elect_one()) {
                    mbar_expect_tx(&full_bar[stage], tx);
                    tma_load_im2col_5d(st, map_a_hi, &full_bar[stage], c0, bw, bh, bt, n_i, c, b, a);
                    tma_load_2d(st + 2 * S::kABytes, &maps.b[0], &full_bar[stage], c0, ftap * p.cd + n0);
                    if (p.x3) {
                        tma_load_im2col_5d(st + S::kABytes, map_a_lo, &full_bar[stage], c0, bw, bh, bt, n_i, c, b, a);
                        tma_load_2d(st + 2 * S::kABytes + S::kBBytes, &maps.b[1], &full_bar[stage], c0, ftap * p.cd + n0);
                    }
                }
                __syncwarp();
                if (++stage == S::kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        constexpr uint32_t idesc = make_idesc_bf16(kBM, BN, 0, 0);
        // bf16x3 as TWO MMAs per k-step instead of three: the hi and lo filter planes are adjacent K-major tiles, i.e. ONE
        // operand of 2*BN rows, so A_hi x [B_hi ; B_lo] is a single N = 2*BN MMA whose two column halves (hi*hi | hi*lo) the
        // epilogue adds; A_lo x B_hi (N = BN) accumulates into the first half.  An M128 MMA fetches (128 + N) x 32 B of
        // operands at ~64 B/clk (measured), so fewer, wider MMAs cut the operand traffic per product by 22 % (BN = 64).
        constexpr uint32_t idesc2 = make_idesc_bf16(kBM, 2 * BN, 0, 0);
        // K-major SW128 tiles: 8-row groups are 1024 bytes apart; a K=16 slice is 32 bytes into the swizzle row.  Descriptors
        // differ only in the start-address field (bytes >> 4): encode once and add offsets, so that the single issuing thread
        // spends a couple of instructions per MMA (an N = 64 MMA is only 32 tensor-pipe cycles).
        const uint64_t desc0 = make_smem_desc_sw128(smem_u32(smem), 16, 1024);
        const bool x3 = p.x3 != 0;
        int stage = 0, phase = 0, it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const TcTile tl = decode_tile<BN>(p, tile, nblocks);
            if (!tl.valid) continue;
            const int nkb = p.cls[tl.cls].ntaps * cblocks;
            const int buf = it & 1;
            mbar_wait(&tmem_empty[buf], ((it >> 1) & 1) ^ 1);       // the epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t acc = tmem_base + buf * 2 * BN;
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint64_t a_hi = desc0 + (uint32_t)((stage * S::kStageBytes) >> 4);
                const uint64_t b_hi = a_hi + (uint32_t)((2 * S::kABytes) >> 4);
                if (elect_one()) {
                    if (p.debug & 8) {
                    } else if (x3) {
#pragma unroll
                        for (int k = 0; k < kBK / 16; ++k) {
                            umma_bf16(acc, a_hi + 2 * k, b_hi + 2 * k, idesc2, (kb | k) != 0);                               // hi*hi | hi*lo
                            umma_bf16(acc, a_hi + (uint32_t)(S::kABytes >> 4) + 2 * k, b_hi + 2 * k, idesc, 1);              // + lo*hi
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < kBK / 16; ++k) umma_bf16(acc, a_hi + 2 * k, b_hi + 2 * k, idesc, (kb | k) != 0);
                    }
                    umma_commit(&empty_bar[stage]);     // frees the smem stage once these MMAs have read it
                    if (kb == nkb - 1) umma_commit(&tmem_full[buf]);           // accumulator complete
                }
                __syncwarp();
                if (++stage == S::kStages) { stage = 0; phase ^= 1; }
            }
            ++it;
        }
    } else if (warp >= 2) {
        // ===== epilogue: TMEM -> registers -> per-warp shared-memory transpose -> coalesced global rows, optional BatchNorm statistics =====
        // EIGHT warps: the per-tile chain (tcgen05.ld -> staging -> stores) is latency-bound with one warp per scheduler, and with
        // short contractions (temporal taps, 64 channels) it -- not the MMAs -- paced the kernel.  Warp w may access TMEM lanes
        // 32 * (w % 4) .. + 31, so each lane quarter is served by two warps that split the tile's 32-column chunks.
        // tcgen05.ld hands every thread one output pixel (row); storing rows from that layout would make each warp store touch 32
        // different lines.  Each warp therefore transposes 32 x 16 half-chunks through a padded staging tile and accesses global
        // memory with 4 lanes per row (64 contiguous bytes = 2 full sectors per row, 8 rows per instruction).  The residual addend
        // and the z tile of the fused BatchNorm backward use the same layout.  Per-channel sums are kept in registers across the
        // CTA's tiles (all tiles of a CTA have the same channel block: gridDim.x is a multiple of the channel-block count) and
        // leave the CTA once, as fp64 atomics.
        constexpr int kChunks = BN / 64;            // 32-column chunks per warp
        const int q = warp & 3;                     // TMEM lane quarter this warp may access
        const int hsel = (warp - 2) >> 2;           // which half of the tile's columns
        float* const stg = s_stage + (warp - 2) * (32 * kStgLd);
        const int r8 = lane >> 2, c4 = (lane & 3) * 4;
        const int n0 = (blockIdx.x % nblocks) * BN; // constant per CTA
        double* const acc_out = stats ? stats : fuse.sums;       // which per-channel sums this launch accumulates, if any
        if (fuse.z) {       // BatchNorm constants of this CTA's channels
            const int t = threadIdx.x - 64;
            if (t < BN) s_par[t] = make_float4(__ldg(fuse.mean + n0 + t), __ldg(fuse.invstd + n0 + t), __ldg(fuse.gamma + n0 + t), __ldg(fuse.beta + n0 + t));
            asm volatile("bar.sync 1, 256;" ::: "memory");
        }
        float run1[kChunks][2][4], run2[kChunks][2][4];      // running column sums: [chunk][half][4 columns of this lane]
#pragma unroll
        for (int a = 0; a < kChunks; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b)
#pragma unroll
                for (int t = 0; t < 4; ++t) run1[a][b][t] = run2[a][b][t] = 0.f;
        int it = 0;
        // element offsets of the destination row this lane reads from TMEM in `tile` and of its addend row (~0: none)
        struct RowPair { unsigned long long row, arow; };
        auto row_of = [&](int tile) -> RowPair {
            RowPair rp{~0ull, ~0ull};
            if (tile >= num_tiles) return rp;
            const TcTile tl = decode_tile<BN>(p, tile, nblocks);
            if (!tl.valid) return rp;
            const TcClass& cl = p.cls[tl.cls];
            const int m = tl.m0 + q * 32 + lane;
            if (m >= cl.M) return rp;
            size_t pix = (size_t)m;
            if (p.strided_out || p.add_w) {        // a stride-parity class of a strided input gradient: scatter rows to their pixels
                int w_o, h_o, t_o;
                const int n_i = fdivmod(fdivmod(fdivmod(m, cl.dwq, w_o), cl.dhq, h_o), cl.dtq, t_o);
                const int td = t_o * p.ot + cl.rt, hd = h_o * p.oh + cl.rh, wd = w_o * p.ow + cl.rw;
                pix = (((size_t)n_i * p.Td + td) * p.Hd + hd) * p.Wd + wd;
                if (p.add_w) {
                    int rt_, rh_, rw_;
                    const int ta = fdivmod(td, p.dadd_t, rt_), ha = fdivmod(hd, p.dadd_h, rh_), wa = fdivmod(wd, p.dadd_w, rw_);
                    if ((rt_ | rh_ | rw_) == 0) rp.arow = ((((size_t)n_i * p.add_T + ta) * p.add_H + ha) * p.add_W + wa) * p.cd + n0;
                }
            }
            rp.row = (unsigned long long)(pix * p.cd + n0);
            if (!p.add_w) rp.arow = rp.row;
            return rp;
        };
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const RowPair mine = row_of(tile);
            const unsigned long long my_row = mine.row;
            if (!decode_tile<BN>(p, tile, nblocks).valid) continue;
            const int buf = it & 1;
            if (addend || fuse.z) {
                // the rows of the CTA's NEXT tile on their way to L2 (a 32-column chunk of a row is one 128-byte line)
                const RowPair nxt = row_of(tile + gridDim.x);
#pragma unroll
                for (int jj = 0; jj < kChunks; ++jj) {
                    if (addend && nxt.arow != ~0ull) asm volatile("prefetch.global.L2 [%0];" ::"l"(addend + nxt.arow + (hsel * kChunks + jj) * 32));
                    if (fuse.z && nxt.row != ~0ull) asm volatile("prefetch.global.L2 [%0];" ::"l"(fuse.z + nxt.row + (hsel * kChunks + jj) * 32));
                }
            }
            unsigned long long rows4[4], rows4a[4]; // element offsets of the rows this lane serves in the coalesced layout (~0: no row) / their addend rows
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                rows4[i] = __shfl_sync(0xffffffffu, my_row, i * 8 + r8);
                rows4a[i] = __shfl_sync(0xffffffffu, mine.arow, i * 8 + r8);
            }
            // BN = 64 (two half-chunks per warp, registers to spare): the residual addend and the z tile of the fused BatchNorm
            // backward of the WHOLE tile are requested before the accumulator wait -- their addresses depend on the tile index
            // only.  Requested per half-chunk behind the wait, the strided conv3x-entry input gradient (both operands, 85 tiles per
            // CTA) spent ~5 us per tile in two serial [DRAM round trip + 300 instructions] periods per warp: 421 us for 1.4 GB.
            constexpr bool kPre = BN == 64;
            float4 ad_pre[kPre ? 2 : 1][4], zz_pre[kPre ? 2 : 1][4];
            if (kPre) {
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int col = hsel * 32 + hh * 16;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        ad_pre[hh][i] = (addend && rows4a[i] != ~0ull) ? __ldg(reinterpret_cast<const float4*>(addend + rows4a[i] + col + c4)) : make_float4(0.f, 0.f, 0.f, 0.f);
                        zz_pre[hh][i] = (fuse.z && rows4[i] != ~0ull) ? __ldg(reinterpret_cast<const float4*>(fuse.z + rows4[i] + col + c4)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
            }
            mbar_wait_sleep(&tmem_full[buf], (it >> 1) & 1, 128);
            tc_fence_after();
            if (p.debug & 4) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[buf]);
                ++it;
                continue;
            }
#pragma unroll
            for (int hc = 0; hc < 2 * kChunks; ++hc) {      // 16-column half-chunks of this warp
                const int jj = hc >> 1, hh = hc & 1;
                const int col = (hsel * kChunks + jj) * 32 + hh * 16;       // first column of the half-chunk inside the tile
                uint32_t r[16], r2[16];
                const uint32_t taddr = tmem_base + buf * 2 * BN + ((uint32_t)(q * 32) << 16) + col;
                tmem_ld_32x32b_x16(taddr, r);
                if (p.x3) tmem_ld_32x32b_x16(taddr + BN, r2);       // the hi*lo half of the bf16x3 accumulator
                // global operands of this half-chunk are requested while the TMEM loads are in flight
                float4 ad[4], zz[4];
                if (kPre) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        ad[i] = ad_pre[kPre ? hh : 0][i];
                        zz[i] = zz_pre[kPre ? hh : 0][i];
                    }
                } else {
                    if (addend) {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            ad[i] = rows4a[i] != ~0ull ? __ldg(reinterpret_cast<const float4*>(addend + rows4a[i] + col + c4)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    if (fuse.z) {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            zz[i] = rows4[i] != ~0ull ? __ldg(reinterpret_cast<const float4*>(fuse.z + rows4[i] + col + c4)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
                tmem_ld_wait();
                if (hc == 2 * kChunks - 1) {        // this warp's part of the accumulator is in registers: hand the TMEM buffer back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty[buf]);
                }
                float4* const srow = reinterpret_cast<float4*>(stg + lane * kStgLd);
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    float4 o = make_float4(__uint_as_float(r[4 * v]), __uint_as_float(r[4 * v + 1]), __uint_as_float(r[4 * v + 2]), __uint_as_float(r[4 * v + 3]));
                    if (p.x3) {
                        o.x += __uint_as_float(r2[4 * v]); o.y += __uint_as_float(r2[4 * v + 1]);
                        o.z += __uint_as_float(r2[4 * v + 2]); o.w += __uint_as_float(r2[4 * v + 3]);
                    }
                    srow[v] = o;
                }
                __syncwarp();
                float4 par[4];
                if (fuse.z) {
#pragma unroll
                    for (int t = 0; t < 4; ++t) par[t] = s_par[col + c4 + t];      // mean, invstd, gamma, beta
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float4 o = *reinterpret_cast<const float4*>(stg + (i * 8 + r8) * kStgLd + c4);
                    if (rows4[i] == ~0ull) continue;
                    if (addend) { o.x += ad[i].x; o.y += ad[i].y; o.z += ad[i].z; o.w += ad[i].w; }
                    if (!(p.debug & 1)) *reinterpret_cast<float4*>(out + rows4[i] + col + c4) = o;
                    if (acc_out && !(p.debug & 2)) {
                        // per-channel sums:
                        //   forward (stats):      sum(o), sum(o^2) of the stored output -> BatchNorm statistics
                        //   input gradient (fuse): sum(g), sum(g * xhat) with g = o * relu'(bn(z)) -> the next BatchNorm backward
                        const float ov[4] = {o.x, o.y, o.z, o.w};
                        if (fuse.z) {
                            const float zv[4] = {zz[i].x, zz[i].y, zz[i].z, zz[i].w};
#pragma unroll
                            for (int t = 0; t < 4; ++t) {
                                const float xh = (zv[t] - par[t].x) * par[t].y;
                                const float g = fmaf(xh, par[t].z, par[t].w) > 0.f ? ov[t] : 0.f;
                                run1[jj][hh][t] += g;
                                run2[jj][hh][t] = fmaf(g, xh, run2[jj][hh][t]);
                            }
                        } else {
#pragma unroll
                            for (int t = 0; t < 4; ++t) {
                                run1[jj][hh][t] += ov[t];
                                run2[jj][hh][t] = fmaf(ov[t], ov[t], run2[jj][hh][t]);
                            }
                        }
                    }
                }
                __syncwarp();                       // the staging tile is rewritten by the next half-chunk
            }
            ++it;
        }
        if (acc_out && !(p.debug & 64)) {
            // lanes with equal (lane & 3) hold partial sums of the same 4 columns: reduce over the 8 row groups of the warp, over the
            // 4 lane quarters through shared memory (the staging tiles are free now), then ONE fp64 atomic per column and CTA --
            // all CTAs finish together, and atomics on the same address serialise in L2
            float* const s_red = s_stage;           // [4 quarters][2 sums][BN]
            asm volatile("bar.sync 1, 256;" ::: "memory");      // every epilogue warp is done with its staging tile
#pragma unroll
            for (int jj = 0; jj < kChunks; ++jj)
#pragma unroll
                for (int hh = 0; hh < 2; ++hh)
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        float a = run1[jj][hh][t], b = run2[jj][hh][t];
#pragma unroll
                        for (int o = 4; o <= 16; o <<= 1) {
                            a += __shfl_xor_sync(0xffffffffu, a, o);
                            b += __shfl_xor_sync(0xffffffffu, b, o);
                        }
                        if (lane < 4) {
                            const int col = (hsel * kChunks + jj) * 32 + hh * 16 + c4 + t;
                            s_red[(q * 2 + 0) * BN + col] = a;
                            s_red[(q * 2 + 1) * BN + col] = b;
                        }
                    }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const int t = threadIdx.x - 64;         // 0..255 = (which sum, column)
            if (t < 2 * BN) {
                const int which = t / BN, col = t - which * BN;
                const float tot = s_red[(0 * 2 + which) * BN + col] + s_red[(1 * 2 + which) * BN + col] + s_red[(2 * 2 + which) * BN + col] +
                                  s_red[(3 * 2 + which) * BN + col];
                atomicAdd(acc_out + (size_t)which * p.cd + n0 + col, (double)tot);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 4 * BN);
}

// ---- filter gradient ---------------------------------------------------------------------------------
// dW[tap][ci][co] = sum_m X[pix(m, tap), ci] * dZ[m, co].  GEMM rows = 128 = two "groups" (tap, 64-channel block)
// of the flattened (taps x ci/64) index space, columns = BN output channels, reduction = destination pixels in
// k-blocks of 64.  Both operands are MN-major SW128 tiles [64 pixels][64 channels] (8 KB, exactly what one TMA
// box delivers): groups of 64 channels are 8 KB apart (LBO), groups of 8 pixels 1 KB apart (SBO), a K = 16 step
// advances 2 KB.
constexpr int kWgKB = 64;                    // pixels per k-block
constexpr int kSubTile = kWgKB * 64 * 2;     // 8 KB: [64 pixels][64 channels] bf16

struct TcWgradParams {
    int M;                  // destination pixels (rows of dZ)
    int td, hd, wd;         // destination extents
    int cs, cd;             // ci, co
    int kt, kh, kw, st, sh, sw, pt, ph, pw;
    int groups;             // taps * cs / 64
    int kb_per_split;       // k-blocks per split
    int x3;
};

// G = (tap, 64-channel) groups per CTA: 2 (one accumulator of 128 rows), or 3 for layers with exactly three groups (the 64-channel
// temporal convolutions): a second accumulator takes the third group (its rows 64..127 multiply whatever follows the group in
// shared memory and are never stored), so the CTA loads dZ ONCE per k-block instead of two CTAs loading it twice -- that
// launch is HBM-bound (6.7 TB/s) and was reading 1.54 GB for 0.82 GB of operands.
template <int BN, int G = 2>
struct TcWgSmem {
    static constexpr int kABytes = G * kSubTile;            // G (tap, 64-channel) groups
    static constexpr int kBBytes = (BN / 64) * kSubTile;
    static constexpr int kStageBytes = 2 * (kABytes + kBBytes);             // [A hi | A lo | B hi | B lo]
    static constexpr int kStages = BN <= 64 ? (G == 3 ? 3 : 4) : (BN <= 128 ? 3 : 2);
    // bf16x3 with BN <= 128: X_hi x [dZ_hi | dZ_lo] is ONE MMA of N = 2 * BN (the lo plane's 64-channel groups follow the hi
    // plane's at the same 8 KB pitch), its two column halves are summed in the epilogue; X_lo x dZ_hi adds into the first half.
    static constexpr bool kWide = BN <= 128;
    static constexpr int kAccCols = kWide ? 2 * BN : BN;                    // columns of one accumulator
    static constexpr int kTmemCols = (G == 3 ? 2 : 1) * kAccCols;
    static constexpr int kBytes = kStages * kStageBytes + 1024 + 256;
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int BN, int G>
__global__ void __launch_bounds__(kTcThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_x_hi, const __grid_constant__ CUtensorMap map_x_lo,
                const __grid_constant__ CUtensorMap map_d_hi, const __grid_constant__ CUtensorMap map_d_lo, const TcWgradParams p,
                float* __restrict__ dfilt) {
    using S = TcWgSmem<BN, G>;
    constexpr int kAcc = G == 3 ? 2 : 1;                 // accumulators of 128 rows
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kStages * S::kStageBytes);
    uint64_t* empty_bar = full_bar + S::kStages;
    uint64_t* accum_bar = empty_bar + S::kStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const int g0 = blockIdx.x * G;                       // first (tap, channel-block) group of this CTA
    const int n0 = blockIdx.y * BN;
    const int cblocks = p.cs / 64;
    const int total_kb = (p.M + kWgKB - 1) / kWgKB;
    const int kb0 = blockIdx.z * p.kb_per_split;
    const int nkb = min(p.kb_per_split, total_kb - kb0);

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&map_x_hi);
        prefetch_tensormap(&map_d_hi);
        for (int s = 0; s < S::kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(accum_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, S::kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    if (warp == 0) {
        // ===== TMA producer (whole warp, one elected lane issues: see conv_tc_kernel) =====
        int gtap[G], gc0[G], ga[G], gb[G], gc[G];
#pragma unroll
        for (int i = 0; i < G; ++i) {
            const int g = min(g0 + i, p.groups - 1);     // an odd tail group is loaded twice and not stored
            gtap[i] = g / cblocks;
            gc0[i] = (g - gtap[i] * cblocks) * 64;
            ga[i] = gtap[i] / (p.kh * p.kw);
            const int r = gtap[i] - ga[i] * p.kh * p.kw;
            gb[i] = r / p.kw;
            gc[i] = r - gb[i] * p.kw;
        }
        const uint32_t tx = (uint32_t)(S::kABytes + S::kBBytes) * (p.x3 ? 2u : 1u);
        int stage = 0, phase = 0;
        // Temporal filters: walk the k-blocks frame-fastest (same 64 pixels of consecutive frames back to back), so that the three
        // reads of an input tile (taps t-1, t, t+1) are 1-2 k-blocks apart and hit L2.  In pixel order they were a whole frame of
        // the CTA's range apart -- with 148 CTAs streaming concurrently that is ~1 GB of other traffic, and X came from DRAM
        // three times (1.54 GB read for 0.82 GB of operands, HBM-bound).  Needs k-blocks that do not straddle frames.
        const int hw_blocks = (p.hd * p.wd) / kWgKB;
        const bool frame_fastest = p.kt > 1 && (p.hd * p.wd) % kWgKB == 0 && p.td > 1;
        for (int kb = 0; kb < nkb; ++kb) {
            int blk = kb0 + kb;
            if (frame_fastest) {
                const int per_clip = p.td * hw_blocks;
                const int n_c = blk / per_clip, rem = blk - n_c * per_clip;
                const int hb = rem / p.td, t_f = rem - hb * p.td;
                blk = (n_c * p.td + t_f) * hw_blocks + hb;
            }
            int m = blk * kWgKB;
            const int blk_row = m;
            const int w_o = m % p.wd;  m /= p.wd;
            const int h_o = m % p.hd;  m /= p.hd;
            const int t_o = m % p.td;
            const int n_i = m / p.td;
            const int bw = w_o * p.sw - p.pw, bh = h_o * p.sh - p.ph, bt = t_o * p.st - p.pt;
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* st = smem + stage * S::kStageBytes;
            if (elect_one()) {
                mbar_expect_tx(&full_bar[stage], tx);
                const int planes = p.x3 ? 2 : 1;
                for (int pl = 0; pl < planes; ++pl) {
                    const CUtensorMap* mx = pl ? &map_x_lo : &map_x_hi;
                    const CUtensorMap* md = pl ? &map_d_lo : &map_d_hi;
                    uint8_t* a = st + pl * S::kABytes;
                    uint8_t* b = st + 2 * S::kABytes + pl * S::kBBytes;
#pragma unroll
                    for (int i = 0; i < G; ++i)
                        tma_load_im2col_5d(a + i * kSubTile, mx, &full_bar[stage], gc0[i], bw, bh, bt, n_i, (uint16_t)gc[i], (uint16_t)gb[i], (uint16_t)ga[i]);
#pragma unroll
                    for (int j = 0; j < BN / 64; ++j)
                        tma_load_2d(b + j * kSubTile, md, &full_bar[stage], n0 + j * 64, blk_row);
                }
            }
            __syncwarp();
            if (++stage == S::kStages) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        constexpr uint32_t idesc = make_idesc_bf16(128, BN, 1, 1);
        constexpr uint32_t idesc2 = make_idesc_bf16(128, S::kWide ? 2 * BN : BN, 1, 1);
        const uint64_t desc0 = make_smem_desc_sw128(smem_u32(smem), kSubTile, 1024);
        const bool x3 = p.x3 != 0;
        int stage = 0, phase = 0;
        for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint64_t a_hi = desc0 + (uint32_t)((stage * S::kStageBytes) >> 4);
            const uint64_t b_hi = a_hi + (uint32_t)((2 * S::kABytes) >> 4);
            constexpr uint32_t kLoA = (uint32_t)(S::kABytes >> 4), kLoB = (uint32_t)(S::kBBytes >> 4);      // hi -> lo plane of an operand
            if (elect_one()) {
#pragma unroll
            for (int acc_i = 0; acc_i < kAcc; ++acc_i) {
                const uint32_t acc = tmem_base + acc_i * S::kAccCols;
                const uint64_t a_g = a_hi + (uint32_t)((acc_i * 2 * kSubTile) >> 4);       // groups 0,1 | group 2 (+ 8 KB of don't-care rows)
                if (x3) {
#pragma unroll
                    for (int k = 0; k < kWgKB / 16; ++k) {          // a K = 16 step advances 2 KB = 128 encoded units
                        if (S::kWide) {
                            umma_bf16(acc, a_g + 128 * k, b_hi + 128 * k, idesc2, (kb | k) != 0);       // hi*hi | hi*lo
                            umma_bf16(acc, a_g + kLoA + 128 * k, b_hi + 128 * k, idesc, 1);             // + lo*hi
                        } else {
                            umma_bf16(acc, a_g + 128 * k, b_hi + 128 * k, idesc, (kb | k) != 0);
                            umma_bf16(acc, a_g + 128 * k, b_hi + kLoB + 128 * k, idesc, 1);
                            umma_bf16(acc, a_g + kLoA + 128 * k, b_hi + 128 * k, idesc, 1);
                        }
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < kWgKB / 16; ++k) umma_bf16(acc, a_g + 128 * k, b_hi + 128 * k, idesc, (kb | k) != 0);
                }
            }
            umma_commit(&empty_bar[stage]);
            if (kb == nkb - 1) umma_commit(accum_bar);
            }
            __syncwarp();
            if (++stage == S::kStages) { stage = 0; phase ^= 1; }
        }
    } else if (warp >= 2 && nkb > 0) {
        // ===== epilogue: accumulator row = (group, channel) -> fp32 atomics into dW[tap][ci][co] =====
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        const int q = warp & 3;
        const int row = q * 32 + lane;              // 0..127
#pragma unroll
        for (int acc_i = 0; acc_i < kAcc; ++acc_i) {
        const int g = g0 + 2 * acc_i + (row >> 6);
        const bool valid = g < p.groups && g < g0 + G;
        const int tap = g / cblocks, ci = (g - tap * cblocks) * 64 + (row & 63);
        float* dst = dfilt + ((size_t)tap * p.cs + ci) * p.cd + n0;
        const uint32_t tacc = tmem_base + acc_i * S::kAccCols + ((uint32_t)(q * 32) << 16);
#pragma unroll
        for (int j = 0; j < BN / 32; ++j) {
            uint32_t r[32], r2[32];
            tmem_ld_32x32b_x32(tacc + j * 32, r);
            const bool wide = S::kWide && p.x3;
            if (wide) tmem_ld_32x32b_x32(tacc + BN + j * 32, r2);     // the hi*lo half
            tmem_ld_wait();
            if (valid) {
#pragma unroll
                for (int v = 0; v < 8; ++v) {
                    float o[4];
#pragma unroll
                    for (int t = 0; t < 4; ++t) o[t] = wide ? __uint_as_float(r[4 * v + t]) + __uint_as_float(r2[4 * v + t]) : __uint_as_float(r[4 * v + t]);
                    red_add_v4(dst + j * 32 + 4 * v, o[0], o[1], o[2], o[3]);
                }
            }
        }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, S::kTmemCols);
}

// fp32 -> (bf16 hi, bf16 lo) planes: hi = rn(x), lo = rn(x - hi)
__global__ void __launch_bounds__(256) split_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                                         int64_t n4) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
        const float f[4] = {v.x, v.y, v.z, v.w};
        __nv_bfloat16 h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            h[j] = __float2bfloat16_rn(f[j]);
            l[j] = __float2bfloat16_rn(f[j] - __bfloat162float(h[j]));
        }
        reinterpret_cast<uint2*>(hi)[i] = *reinterpret_cast<uint2*>(h);
        if (lo) reinterpret_cast<uint2*>(lo)[i] = *reinterpret_cast<uint2*>(l);
    }
}

// ---- host side ------------------------------------------------------------------------------------
static int encode_im2col(CUtensorMap* map, const void* base, int n, int t, int h, int w, int c, const int lower[3], const int upper[3],
                         const int stride[3], int pixels) {
    const TensorMapApi& api = tensor_map_api();
    if (!api.ok) { set_error("conv_tc: cuTensorMapEncode* driver entry points unavailable"); return AVID_ECUDA; }
    cuuint64_t dims[5] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)t, (cuuint64_t)n};
    cuuint64_t strides[4] = {(cuuint64_t)c * 2, (cuuint64_t)w * c * 2, (cuuint64_t)h * w * c * 2, (cuuint64_t)t * h * w * c * 2};
    cuuint32_t estr[5] = {1, (cuuint32_t)stride[0], (cuuint32_t)stride[1], (cuuint32_t)stride[2], 1};
    CUresult r = api.im2col(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, lower, upper, kBK, (cuuint32_t)pixels,
                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv_tc: cuTensorMapEncodeIm2col failed (%d)", (int)r); return AVID_ECUDA; }
    // driver <= 13.1 mis-encodes im2col maps of tensors smaller than 128 KiB (same work-around as CUTLASS)
    int drv = 0;
    cudaDriverGetVersion(&drv);
    if (drv <= 13010 && (size_t)n * t * h * w * c * 2 < 131072) reinterpret_cast<uint64_t*>(map)[1] &= ~(1ull << 21);
    return AVID_OK;
}

static int encode_tiled_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols) {
    const TensorMapApi& api = tensor_map_api();
    if (!api.ok) { set_error("conv_tc: cuTensorMapEncode* driver entry points unavailable"); return AVID_ECUDA; }
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = api.tiled(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv_tc: cuTensorMapEncodeTiled failed (%d)", (int)r); return AVID_ECUDA; }
    return AVID_OK;
}

template <int BN>
static int launch_conv_tc(const TcConvMaps& maps, const TcConvParams& p, const float* addend, float* out, double* stats, const BnBwdFuse& fuse,
                          cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem<BN>::kBytes);
        if (e != cudaSuccess) { set_error("conv_tc: smem attribute: %s", cudaGetErrorString(e)); return AVID_ECUDA; }
        configured = true;
    }
    const int nblocks = p.cd / BN;
    const int tiles = p.mtiles * p.ncls * nblocks;
    // a multiple of the channel-block count, so that every tile of a CTA has the same channel block (running statistics)
    const int grid = tiles < kNumSMs ? tiles : (kNumSMs / nblocks) * nblocks;
    AVID_REQUIRE(grid > 0, "conv_tc: %d channel blocks exceed the SM count", nblocks);
    conv_tc_kernel<BN><<<grid, kConvThreads, TcSmem<BN>::kBytes, st>>>(maps, p, addend, out, stats, fuse);
    return check_launch("conv_tc_kernel");
}

// One dimension of one launch: which destination coordinates it enumerates, which source coordinates / filter taps they read.
struct DimPlan {
    int cnt = 0;            // enumerated destination coordinates
    int base = 0;           // source coordinate of tap offset 0 for destination coordinate 0 (== im2col lower corner)
    int upper = 0;          // im2col upper corner
    int trav = 1;           // source traversal stride
    int ntap = 0;
    int off[8], ftap[8];    // source offset (>= 0, relative to base) and filter tap index
    int ostride = 1, r = 0; // destination coordinate = enumerated * ostride + r
};

// forward: destination = conv output, source = input
static DimPlan plan_forward(int dst, int k, int s, int pad) {
    DimPlan d;
    d.cnt = dst;  d.base = -pad;  d.upper = pad - (k - 1);  d.trav = s;  d.ntap = k;
    for (int j = 0; j < k; ++j) { d.off[j] = j;  d.ftap[j] = j; }
    return d;
}
// input gradient, destination coordinates i = q * s + r:  din[i] = sum_b dout[(i + pad - b) / s] w[b] over the taps b with
// (i + pad - b) divisible by s, i.e. a stride-1 correlation of dout with the taps of residue class r
static DimPlan plan_dgrad(int dst, int src, int k, int s, int pad, int r) {
    DimPlan d;
    d.ostride = s;  d.r = r;  d.trav = 1;
    d.cnt = r < dst ? (dst - r + s - 1) / s : 0;
    int lo = 1 << 30;
    for (int b = 0; b < k; ++b) {
        const int x = r + pad - b;
        if (((x % s) + s) % s != 0) continue;
        const int off = (x - (((x % s) + s) % s)) / s;      // exact: x divisible by s
        d.off[d.ntap] = off;  d.ftap[d.ntap] = b;  ++d.ntap;
        if (off < lo) lo = off;
    }
    if (d.ntap == 0) return d;
    for (int j = 0; j < d.ntap; ++j) d.off[j] -= lo;
    d.base = lo;
    d.upper = lo + d.cnt - src;
    return d;
}

// dgrad == 0: out[n,to,ho,wo,co] = conv(in, filt);  a_* = input planes [n,ti,hi,wi,ci], b_* = filter planes [taps][co][ci]
// dgrad == 1: out[n,ti,hi,wi,ci] = conv_transpose(dout, filt); a_* = dout planes, b_* = filter planes [taps][ci][co].  A strided
//             input gradient is st * sh * sw stride-parity classes, each a stride-1 correlation with its own taps and im2col
//             box; all classes run in ONE launch (tile -> class, see decode_tile).
int conv_tc_run(const avid_conv_shape_t* s, int dgrad, const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo,
                const float* addend, const int32_t* addend_stride, float* out, double* stats, const BnBwdFuse& fuse, cudaStream_t st) {
    AVID_REQUIRE(s && a_hi && b_hi && out, "conv_tc: NULL pointer");
    AVID_REQUIRE(!(stats && fuse.z), "conv_tc: forward statistics and the fused BatchNorm backward reduction are exclusive");
    AVID_REQUIRE(!fuse.z || (fuse.mean && fuse.invstd && fuse.gamma && fuse.beta && fuse.sums), "conv_tc: incomplete BatchNorm fusion arguments");
    AVID_REQUIRE((a_lo == nullptr) == (b_lo == nullptr), "conv_tc: give both lo planes (bf16x3) or neither (bf16)");
    AVID_REQUIRE(s->kt >= 1 && s->kh >= 1 && s->kw >= 1 && s->kt <= 8 && s->kh <= 8 && s->kw <= 8 && s->kt * s->kh * s->kw <= kMaxTaps,
                 "conv_tc: filter %dx%dx%d not supported (at most 8 per dimension, %d taps)", s->kt, s->kh, s->kw, kMaxTaps);
    AVID_REQUIRE(s->st >= 1 && s->sh >= 1 && s->sw >= 1, "conv_tc: bad stride");
    const int cs = dgrad ? s->co : s->ci, cd = dgrad ? s->ci : s->co;
    AVID_REQUIRE(cs % kBK == 0, "conv_tc: source channels (%d) must be a multiple of 64", cs);
    AVID_REQUIRE(cd % 64 == 0, "conv_tc: destination channels (%d) must be a multiple of 64", cd);
    const int src[3] = {dgrad ? s->wo : s->wi, dgrad ? s->ho : s->hi, dgrad ? s->to : s->ti};   // (w, h, t) order, like the tensor map
    const int dst[3] = {dgrad ? s->wi : s->wo, dgrad ? s->hi : s->ho, dgrad ? s->ti : s->to};
    const int kk[3] = {s->kw, s->kh, s->kt}, ss[3] = {s->sw, s->sh, s->st}, pp[3] = {s->pw, s->ph, s->pt};
    const int64_t dst_pixels = (int64_t)s->n * dst[0] * dst[1] * dst[2];
    AVID_REQUIRE(dst_pixels > 0 && dst_pixels < ((int64_t)1 << 31) - 256, "conv_tc: bad pixel count");
    const int bn = cd % 128 == 0 ? 128 : 64;
    const int taps_total = s->kt * s->kh * s->kw;
    const bool x3 = a_lo != nullptr;
    int rc;
    // the 64 -> 64 channel 1x3x3 stride-1 layers (forward and input gradient) run on CTA pairs with a halo tile and a resident filter
    const bool sub_addend = addend && addend_stride && (addend_stride[0] > 1 || addend_stride[1] > 1 || addend_stride[2] > 1);
    AVID_REQUIRE(!sub_addend || (dgrad && addend_stride[0] >= 1 && addend_stride[1] >= 1 && addend_stride[2] >= 1),
                 "conv_tc: a subsampled addend belongs to an input gradient and needs positive strides");
    if (!sub_addend) {
        rc = conv_pair_run(s, dgrad, a_hi, a_lo, b_hi, b_lo, addend, out, stats, fuse, st);
        if (rc != AVID_EUNSUPPORTED) return rc;
    }
    TcConvMaps maps;
    if ((rc = encode_tiled_2d(&maps.b[0], b_hi, (uint64_t)taps_total * cd, cs, bn, kBK))) return rc;
    maps.b[1] = maps.b[0];
    if (x3 && (rc = encode_tiled_2d(&maps.b[1], b_lo, (uint64_t)taps_total * cd, cs, bn, kBK))) return rc;

    const int classes[3] = {dgrad ? ss[0] : 1, dgrad ? ss[1] : 1, dgrad ? ss[2] : 1};
    AVID_REQUIRE(classes[0] * classes[1] * classes[2] <= kMaxCls, "conv_tc: stride %dx%dx%d has more than %d parity classes", ss[2], ss[1], ss[0], kMaxCls);
    // classes whose residue meets no filter tap (filter smaller than the stride) receive no gradient: zero fill first
    bool any_empty = false;
    if (dgrad)
        for (int d = 0; d < 3; ++d)
            for (int r = 0; r < classes[d]; ++r) any_empty |= plan_dgrad(dst[d], src[d], kk[d], ss[d], pp[d], r).ntap == 0;
    if (any_empty) {
        AVID_REQUIRE(addend == nullptr, "conv_tc: an addend is not supported when the filter is smaller than the stride");
        cudaError_t e = cudaMemsetAsync(out, 0, (size_t)dst_pixels * cd * sizeof(float), st);
        if (e != cudaSuccess) { set_error("conv_tc: memset: %s", cudaGetErrorString(e)); return AVID_ECUDA; }
    }
    TcConvParams p;
    p.ncls = 0;
    p.mtiles = 0;
    p.cd = cd;  p.cs = cs;
    p.x3 = x3;
    {
        const char* dbg = getenv("AVID_TC_DEBUG");
        p.debug = dbg ? atoi(dbg) : 0;
    }
    p.strided_out = dgrad && (ss[0] > 1 || ss[1] > 1 || ss[2] > 1);
    p.Wd = dst[0];  p.Hd = dst[1];  p.Td = dst[2];
    p.ow = classes[0];  p.oh = classes[1];  p.ot = classes[2];
    p.sw = dgrad ? 1 : ss[0];  p.sh = dgrad ? 1 : ss[1];  p.st = dgrad ? 1 : ss[2];
    p.add_t = p.add_h = p.add_w = 0;
    p.add_T = p.add_H = p.add_W = 0;
    p.dadd_t = p.dadd_h = p.dadd_w = make_fastdiv(1);
    if (sub_addend) {      // (t, h, w) order like the shape struct
        p.add_t = addend_stride[0];  p.add_h = addend_stride[1];  p.add_w = addend_stride[2];
        p.add_T = (dst[2] + p.add_t - 1) / p.add_t;  p.add_H = (dst[1] + p.add_h - 1) / p.add_h;  p.add_W = (dst[0] + p.add_w - 1) / p.add_w;
        p.dadd_t = make_fastdiv(p.add_t);  p.dadd_h = make_fastdiv(p.add_h);  p.dadd_w = make_fastdiv(p.add_w);
    }
    int ntaps_all = 0;
    for (int rt = 0; rt < classes[2]; ++rt)
        for (int rh = 0; rh < classes[1]; ++rh)
            for (int rw = 0; rw < classes[0]; ++rw) {
                const int rr[3] = {rw, rh, rt};
                DimPlan d[3];
                bool empty = false;
                for (int i = 0; i < 3; ++i) {
                    d[i] = dgrad ? plan_dgrad(dst[i], src[i], kk[i], ss[i], pp[i], rr[i]) : plan_forward(dst[i], kk[i], ss[i], pp[i]);
                    empty |= d[i].ntap == 0 || d[i].cnt == 0;
                }
                if (empty) continue;
                TcClass& c = p.cls[p.ncls];
                c.wq = d[0].cnt;  c.hq = d[1].cnt;  c.tq = d[2].cnt;
                c.M = s->n * c.tq * c.hq * c.wq;
                c.dwq = make_fastdiv(c.wq);  c.dhq = make_fastdiv(c.hq);  c.dtq = make_fastdiv(c.tq);
                c.bw = d[0].base;  c.bh = d[1].base;  c.bt = d[2].base;
                c.rw = d[0].r;  c.rh = d[1].r;  c.rt = d[2].r;
                c.tap0 = ntaps_all;
                c.ntaps = d[2].ntap * d[1].ntap * d[0].ntap;
                AVID_REQUIRE(ntaps_all + c.ntaps <= kMaxTaps, "conv_tc: more than %d taps", kMaxTaps);
                for (int a = 0; a < d[2].ntap; ++a)
                    for (int b = 0; b < d[1].ntap; ++b)
                        for (int e = 0; e < d[0].ntap; ++e) {
                            const uint32_t ftap = (uint32_t)((d[2].ftap[a] * s->kh + d[1].ftap[b]) * s->kw + d[0].ftap[e]);
                            p.taps[ntaps_all++] = (uint32_t)d[0].off[e] | ((uint32_t)d[1].off[b] << 8) | ((uint32_t)d[2].off[a] << 16) | (ftap << 24);
                        }
                const int lower[3] = {d[0].base, d[1].base, d[2].base};
                const int upper[3] = {d[0].upper, d[1].upper, d[2].upper};
                const int stride[3] = {d[0].trav, d[1].trav, d[2].trav};
                for (int i = 0; i < 3; ++i)
                    AVID_REQUIRE(lower[i] >= -16 && lower[i] <= 15 && upper[i] >= -16 && upper[i] <= 15 &&
                                     (src[i] + upper[i] - lower[i] - 1) / stride[i] + 1 == d[i].cnt,
                                 "conv_tc: geometry not expressible as an im2col box (dim %d: lower %d upper %d)", i, lower[i], upper[i]);
                if ((rc = encode_im2col(&maps.a[p.ncls][0], a_hi, s->n, src[2], src[1], src[0], cs, lower, upper, stride, kBM))) return rc;
                maps.a[p.ncls][1] = maps.a[p.ncls][0];
                if (x3 && (rc = encode_im2col(&maps.a[p.ncls][1], a_lo, s->n, src[2], src[1], src[0], cs, lower, upper, stride, kBM))) return rc;
                const int mt = (c.M + kBM - 1) / kBM;
                if (mt > p.mtiles) p.mtiles = mt;
                ++p.ncls;
            }
    if (p.ncls == 0) return AVID_OK;
    p.dncls = make_fastdiv(p.ncls);
    p.dnblocks = make_fastdiv(cd / bn);
    for (int c = p.ncls; c < kMaxCls; ++c) {      // unused slots: defined bytes in the parameter block
        p.cls[c] = p.cls[0];
        maps.a[c][0] = maps.a[0][0];
        maps.a[c][1] = maps.a[0][1];
    }
    for (int t = ntaps_all; t < kMaxTaps; ++t) p.taps[t] = 0;
    return bn == 128 ? launch_conv_tc<128>(maps, p, addend, out, stats, fuse, st) : launch_conv_tc<64>(maps, p, addend, out, stats, fuse, st);
}

template <int BN, int G>
static int launch_wgrad_tc(const CUtensorMap* maps, const TcWgradParams& p, int splits, float* dfilt, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel<BN, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcWgSmem<BN, G>::kBytes);
        if (e != cudaSuccess) { set_error("wgrad_tc: smem attribute: %s", cudaGetErrorString(e)); return AVID_ECUDA; }
        configured = true;
    }
    dim3 grid((p.groups + G - 1) / G, p.cd / BN, splits);
    wgrad_tc_kernel<BN, G><<<grid, kTcThreads, TcWgSmem<BN, G>::kBytes, st>>>(maps[0], maps[1], maps[2], maps[3], p, dfilt);
    return check_launch("wgrad_tc_kernel");
}

// dfilt[taps][ci][co] += sum over pixels; x_* planes [n,ti,hi,wi,ci], d_* planes [n,to,ho,wo,co]; dfilt zeroed by the caller
int wgrad_tc_run(const avid_conv_shape_t* s, const void* x_hi, const void* x_lo, const void* d_hi, const void* d_lo, float* dfilt, cudaStream_t st) {
    AVID_REQUIRE(s && x_hi && d_hi && dfilt, "wgrad_tc: NULL pointer");
    AVID_REQUIRE((x_lo == nullptr) == (d_lo == nullptr), "wgrad_tc: give both lo planes (bf16x3) or neither (bf16)");
    AVID_REQUIRE(s->ci % 64 == 0 && s->co % 64 == 0, "wgrad_tc: ci (%d) and co (%d) must be multiples of 64", s->ci, s->co);
    TcWgradParams p;
    p.td = s->to; p.hd = s->ho; p.wd = s->wo;
    p.cs = s->ci; p.cd = s->co;
    p.kt = s->kt; p.kh = s->kh; p.kw = s->kw;
    p.st = s->st; p.sh = s->sh; p.sw = s->sw;
    p.pt = s->pt; p.ph = s->ph; p.pw = s->pw;
    p.x3 = x_lo != nullptr;
    const int64_t M = (int64_t)s->n * s->to * s->ho * s->wo;
    AVID_REQUIRE(M > 0 && M < ((int64_t)1 << 31) - 256, "wgrad_tc: bad pixel count");
    p.M = (int)M;
    const int taps = s->kt * s->kh * s->kw;
    p.groups = taps * (s->ci / 64);
    const int bn = s->co % 256 == 0 ? 256 : (s->co % 128 == 0 ? 128 : 64);
    // 64 output channels and a multiple of three (tap, 64-channel) groups -- the 64-channel temporal layers (3 groups) and the 64 -> 64
    // channel 3x3 layers (9 groups: conv2x spatial, audio block1): a CTA takes THREE groups in two accumulators (see TcWgSmem).  With
    // two groups per CTA the 3x3 layers ran 5 CTAs per pixel range (the fifth with a dummy group), each loading its own dZ tile:
    // 5 x 48 KB of operands per 64 pixels against 3 x 64 KB, on a launch that is bound by the L2 -> shared-memory stream.
    const bool three = p.groups % 3 == 0 && bn == 64;
    const int tiles = three ? p.groups / 3 : ((p.groups + 1) / 2) * (s->co / bn);
    const int total_kb = (p.M + kWgKB - 1) / kWgKB;
    // one CTA per SM (the stages fill shared memory): pick the pixel split that minimises waves x (k-blocks per CTA + the fixed
    // prologue / atomic-epilogue cost, ~8 k-blocks) -- a grid of 300 CTAs on 148 SMs would run a third wave for 4 CTAs.
    int splits = 1;
    {
        const int max_splits = total_kb >= 4 ? (total_kb + 3) / 4 : 1;       // at least 4 k-blocks per CTA
        long best = -1;
        for (int sp = 1; sp <= max_splits && (long)sp * tiles <= 4L * kNumSMs; ++sp) {
            const long waves = ((long)sp * tiles + kNumSMs - 1) / kNumSMs;
            const long cost = waves * ((total_kb + sp - 1) / sp + 8);
            if (best < 0 || cost < best) { best = cost; splits = sp; }
        }
    }
    p.kb_per_split = (total_kb + splits - 1) / splits;
    splits = (total_kb + p.kb_per_split - 1) / p.kb_per_split;
    const int lower[3] = {-s->pw, -s->ph, -s->pt};
    const int upper[3] = {s->pw - (s->kw - 1), s->ph - (s->kh - 1), s->pt - (s->kt - 1)};
    const int stride[3] = {s->sw, s->sh, s->st};
    CUtensorMap maps[4];
    int rc;
    if ((rc = encode_im2col(&maps[0], x_hi, s->n, s->ti, s->hi, s->wi, s->ci, lower, upper, stride, kWgKB))) return rc;
    if ((rc = encode_tiled_2d(&maps[2], d_hi, (uint64_t)M, s->co, kWgKB, 64))) return rc;
    maps[1] = maps[0];
    maps[3] = maps[2];
    if (p.x3) {
        if ((rc = encode_im2col(&maps[1], x_lo, s->n, s->ti, s->hi, s->wi, s->ci, lower, upper, stride, kWgKB))) return rc;
        if ((rc = encode_tiled_2d(&maps[3], d_lo, (uint64_t)M, s->co, kWgKB, 64))) return rc;
    }
    if (bn == 256) return launch_wgrad_tc<256, 2>(maps, p, splits, dfilt, st);
    if (bn == 128) return launch_wgrad_tc<128, 2>(maps, p, splits, dfilt, st);
    if (three) return launch_wgrad_tc<64, 3>(maps, p, splits, dfilt, st);
    return launch_wgrad_tc<64, 2>(maps, p, splits, dfilt, st);
}

}  // namespace avid

using namespace avid;

extern "C" {

int avid_conv_wgrad_tc(const avid_conv_shape_t* s, const void* in_hi, const void* in_lo, const void* dout_hi, const void* dout_lo,
                       float* dfilt, void* stream) {
    return wgrad_tc_run(s, in_hi, in_lo, dout_hi, dout_lo, dfilt, static_cast<cudaStream_t>(stream));
}

int avid_split_bf16(const float* x, void* hi, void* lo, int64_t n, void* stream) {
    AVID_REQUIRE(x && hi && n > 0 && n % 4 == 0, "split_bf16: n=%lld must be a positive multiple of 4", (long long)n);
    int64_t blocks = (n / 4 + 255) / 256;
    if (blocks > 16 * kNumSMs) blocks = 16 * kNumSMs;
    split_bf16_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, static_cast<__nv_bfloat16*>(hi), static_cast<__nv_bfloat16*>(lo), n / 4);
    return check_launch("split_bf16_kernel");
}

int avid_conv_tc_plan(const avid_conv_shape_t* s, int32_t dgrad, int64_t* out) {
    AVID_REQUIRE(s && out, "conv_tc_plan: NULL pointer");
    AVID_REQUIRE(s->st >= 1 && s->sh >= 1 && s->sw >= 1 && s->kt >= 1 && s->kh >= 1 && s->kw >= 1, "conv_tc_plan: bad filter / stride");
    const int src[3] = {dgrad ? s->wo : s->wi, dgrad ? s->ho : s->hi, dgrad ? s->to : s->ti};
    const int dst[3] = {dgrad ? s->wi : s->wo, dgrad ? s->hi : s->ho, dgrad ? s->ti : s->to};
    const int kk[3] = {s->kw, s->kh, s->kt}, ss[3] = {s->sw, s->sh, s->st}, pp[3] = {s->pw, s->ph, s->pt};
    const int classes[3] = {dgrad ? ss[0] : 1, dgrad ? ss[1] : 1, dgrad ? ss[2] : 1};
    int64_t ncls = 0, mtiles = 0, taps = 0, pixels = 0, div_ok = 1;
    for (int rt = 0; rt < classes[2]; ++rt)
        for (int rh = 0; rh < classes[1]; ++rh)
            for (int rw = 0; rw < classes[0]; ++rw) {
                const int rr[3] = {rw, rh, rt};
                DimPlan d[3];
                bool empty = false;
                for (int i = 0; i < 3; ++i) {
                    d[i] = dgrad ? plan_dgrad(dst[i], src[i], kk[i], ss[i], pp[i], rr[i]) : plan_forward(dst[i], kk[i], ss[i], pp[i]);
                    empty |= d[i].ntap == 0 || d[i].cnt == 0;
                }
                if (empty) continue;
                const int64_t M = (int64_t)s->n * d[2].cnt * d[1].cnt * d[0].cnt;
                ++ncls;
                taps += (int64_t)d[2].ntap * d[1].ntap * d[0].ntap;
                pixels += M;
                if ((M + kBM - 1) / kBM > mtiles) mtiles = (M + kBM - 1) / kBM;
                // the multiply-high divisions of the tile -> pixel decode on the class extents: first / last indices and a stride through the range
                for (int i = 0; i < 3; ++i) {
                    const FastDiv f = make_fastdiv(d[i].cnt);
                    const int64_t step = M / 997 + 1;
                    for (int64_t n = 0; n < M; n += step)
                        div_ok &= (int64_t)(((uint64_t)(uint32_t)n * f.magic) >> f.shift) == n / d[i].cnt;
                    const int64_t last = M - 1;
                    div_ok &= (int64_t)(((uint64_t)(uint32_t)last * f.magic) >> f.shift) == last / d[i].cnt;
                }
            }
    out[0] = ncls;  out[1] = mtiles;  out[2] = taps;  out[3] = pixels;  out[4] = div_ok;
    return AVID_OK;
}

int avid_conv_tc_uses_cta_pairs(const avid_conv_shape_t* s, int32_t dgrad) {
    (void)dgrad;      // the pair kernel takes both directions of the layers it takes
    return s && s->ci % 64 == 0 && s->co % 64 == 0 && conv_pair_supported(s) ? 1 : 0;
}

int avid_conv_forward_tc(const avid_conv_shape_t* s, const void* in_hi, const void* in_lo, const void* filt_hi, const void* filt_lo,
                         const float* addend, float* out, double* bn_stats, void* stream) {
    return conv_tc_run(s, 0, in_hi, in_lo, filt_hi, filt_lo, addend, nullptr, out, bn_stats, BnBwdFuse{}, static_cast<cudaStream_t>(stream));
}

int avid_conv_dgrad_tc_sub(const avid_conv_shape_t* s, const void* dout_hi, const void* dout_lo, const void* filt_hi, const void* filt_lo,
                           const float* addend, const int32_t* addend_stride, float* din, const avid_bn_backward_fuse_t* fuse, void* stream) {
    BnBwdFuse f;
    if (fuse) {
        f.z = fuse->z;  f.mean = fuse->mean;  f.invstd = fuse->invstd;  f.gamma = fuse->gamma;  f.beta = fuse->beta;  f.sums = fuse->sums;
        AVID_REQUIRE(f.z, "conv_dgrad_tc: fuse->z is NULL");
    }
    // a class no filter tap reaches is zero-filled, not computed: its rows would be missing from the fused sums
    AVID_REQUIRE(!fuse || (s && s->kt >= s->st && s->kh >= s->sh && s->kw >= s->sw), "conv_dgrad_tc: BatchNorm fusion needs filter >= stride");
    return conv_tc_run(s, 1, dout_hi, dout_lo, filt_hi, filt_lo, addend, addend_stride, din, nullptr, f, static_cast<cudaStream_t>(stream));
}

int avid_conv_dgrad_tc(const avid_conv_shape_t* s, const void* dout_hi, const void* dout_lo, const void* filt_hi, const void* filt_lo,
                       const float* addend, float* din, const avid_bn_backward_fuse_t* fuse, void* stream) {
    return avid_conv_dgrad_tc_sub(s, dout_hi, dout_lo, filt_hi, filt_lo, addend, nullptr, din, fuse, stream);
}

}  // extern "C"
