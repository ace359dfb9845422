// 64 -> 64 channel 1x3x3 stride-1 convolutions (forward and input gradient) on CTA PAIRS: tcgen05.mma.cta_group::2, halo tile,
// shared-memory-resident filter.  These are the layers of the 64-channel stage (video conv2x: 8 x 56 x 56 pixels per clip, audio
// block1): 41 % of the video FLOPs at the smallest N the tensor core sees.
//
// Why (round-2 measurements, scripts/probes/umma_fetch.cu + ncu): an SS-mode MMA reads its operands from shared memory at 128 B/clk,
// whatever the swizzle mode; M128 x N64 x K16 reads 6 KB for 32 cycles of math, so the 64-channel kernels are SHARED-MEMORY-bandwidth
// bound, and the TMA writes of their operands share that bandwidth.  conv_tc_kernel<64> writes 48 KB per 64-channel k-block (each
// activation tile once PER TAP through im2col TMA, the filter tile once per pixel tile) and reads 56 KB: 104 KB per 384 cycles of
// math = the measured 47-51 % tensor-pipe activity.  Here
//   * the activation tile is written ONCE per plane: a strip of R rows x (W + 2) pixels x 64 channels (zero padding by TMA) whose
//     128 MMA rows are 128 CONSECUTIVE strip pixels; tap (dh, dw) is the same strip shifted by dh * (W + 2) + dw pixels, i.e. an
//     operand descriptor that starts dh * (W + 2) + dw rows of 128 bytes later -- inside a 1024-byte swizzle atom (verified:
//     scripts/probes/halo_desc.cu, the swizzle is a function of the absolute shared-memory address);
//   * the filter stays resident: a CTA pair splits the N = 64 filter rows, 32 rows x 9 taps x 2 planes = 72 KB per CTA, loaded once
//     per launch (scripts/probes/pair_mma.cu shows the cta_group::2 protocol);
//   * bf16x3 as TWO MMAs per k-step: A_hi x [B_hi ; B_lo] is ONE M256 x N128 MMA (the hi and lo planes of a tap are adjacent in each
//     CTA's filter half, so the pair's N = 128 columns come out as [hi*hi | hi*lo] of channels 0..31 from CTA 0 and of channels
//     32..63 from CTA 1), A_lo x B_hi an M256 x N64 MMA into a third 64-column accumulator; the epilogue adds the three pieces.
//     Ordered in two phases per tile (all taps of the hi plane, then all taps of the lo plane) so that three plane slots pipeline
//     the loads.  An SM's shared memory serves A (4 KB), its own half of B and the half the peer reads, at 128 B/clk: 8 KB = 64
//     cycles for the wide MMA (64 of math), 6 KB = 48 for the narrow one (32 of math); three N = 64 MMAs took 3 x 48.
// The two CTAs of a pair work on the same strip position of two consecutive frames, so one A descriptor serves both.  Columns 0 and
// W + 1 of the strip are padding: their MMA rows are computed and dropped (W / (W + 2) efficiency), as are the rows past the end of a
// frame.  Measured (config-2 conv2x layer, 118 GFLOP): 269 us = 440 TFLOP/s algorithmic; the 3 x 1.06 x 118 GFLOP of executed MMAs run at
// 1.40 PFLOP/s, the measured sustained bf16 peak of the power-capped GPU (ncu: tensor pipe 87 % active at the lower clock it runs at).
#include <stdio.h>
#include <stdlib.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace avid {
using namespace tc;

constexpr int kPairThreads = 320;     // warp 0: TMA producer, warp 1: TMEM alloc + MMA issuer (leader CTA), warps 2-9: epilogue
constexpr int kPairSlots = 3;         // activation plane slots (2 when three do not fit: wide frames such as the audio tower's 100 x 129)
constexpr int kPairBN = 64;
constexpr int kPairAcc = 192;          // TMEM columns of one accumulator buffer: [hi*hi | hi*lo] x 2 channel halves, then lo*hi
constexpr int kPairBBytes = 9 * 2 * 32 * 128;          // [tap][plane][32 filter rows][64 k] bf16: 72 KB per CTA
constexpr int kPairStgBytes = 4 * 2 * kPairBN * 4;     // per-CTA reduction of the BatchNorm sums: [4 lane quarters][2 sums][64 columns]

struct PairParams {
    int frames, T, H, W, Wp;          // N * T frames of H x W pixels; strip pitch Wp = W + 2
    int tiles_per_frame, num_pairs;   // 128-pixel strip tiles per frame; (frame pairs) x (tiles per frame)
    FastDiv d_tpf, d_wp;              // divisions by tiles_per_frame / Wp in the epilogue's tile -> pixel decode (tc_common.cuh)
    int slots;                        // plane slots in the ring: 3, or 2 for wide frames
    int slot_bytes;                   // pitch of the plane slots: R rows x Wp pixels x 128 B rounded up to 1 KB
    int slot_bytes_tx;                // bytes one strip load delivers (R * Wp * 128)
    uint32_t taps[9];                 // (dw + 1) | (dh + 1) << 8 | filter tap << 24
    int tap_shift8[9];                // ((dh - 1) * Wp + (dw - 1)) * 8: descriptor units (16 B) from the tile's first strip pixel to tap t's
    int x3;
    int debug;                        // AVID_PAIR_DEBUG probe bits (scripts/probe_pair.py; 0 in production): 1 no epilogue stores, 8 no activation
                                      // loads, 256 report the launch geometry
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA loads of a CTA pair: executed by both CTAs, the transaction bytes update the barrier of CTA 0 (peer bit of the address cleared)
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_load_5d_pair(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
                     smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
                 : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(da), "l"(db),
                 "r"(idesc), "r"(accumulate)
                 : "memory");
}
// arrive (once all MMAs issued so far have completed) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"((uint16_t)3)
                 : "memory");
}
// arrive on the barrier at this offset in CTA 0 of the pair.  Default (.release.cta) semantics: the TMEM reads it publishes are ordered by
// tcgen05.wait::ld + tcgen05.fence::before_thread_sync; `.release.cluster` compiled to MEMBAR.ALL.GPU and made every tile's hand-back
// wait for the previous tile's global stores (ncu: 2.1 membar stalls per issued instruction, +100 us per launch)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, 0;\n\tmbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar))
                 : "memory");
}

template <bool X3>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPairThreads, 1)
conv_pair_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                 const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo, const PairParams p,
                 const float* __restrict__ addend, float* __restrict__ out, double* __restrict__ stats, const BnBwdFuse fuse) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* b_smem = smem;                                        // resident filter half: [tap][plane][32 rows][64 k]
    uint8_t* ring = smem + kPairBBytes;                            // [3 slots] activation plane strips
    float4* s_par = reinterpret_cast<float4*>(ring + p.slots * p.slot_bytes);
    float* s_stage = reinterpret_cast<float*>(ring + p.slots * p.slot_bytes + kPairBN * 16);
    uint64_t* a_full = reinterpret_cast<uint64_t*>(ring + p.slots * p.slot_bytes + kPairBN * 16 + kPairStgBytes);
    uint64_t* a_empty = a_full + kPairSlots;
    uint64_t* w_full = a_empty + kPairSlots;
    uint64_t* tmem_full = w_full + 1;       // [2]
    uint64_t* tmem_empty = tmem_full + 2;   // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;      // provably warp-uniform
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    constexpr int planes = X3 ? 2 : 1;

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&map_a_hi);
        prefetch_tensormap(&map_b_hi);
        for (int s = 0; s < kPairSlots; ++s) {
            mbar_init(&a_full[s], 1);
            mbar_init(&a_empty[s], 1);
        }
        mbar_init(w_full, 1);
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tmem_full[b], 1);
            mbar_init(&tmem_empty[b], 16);     // one arrive per epilogue warp of BOTH CTAs (on the leader's barrier)
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                     // both CTAs' barriers exist before anything signals them
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    if (warp == 0) {
        // ===== TMA producer (both CTAs; whole warp, one elected lane issues): this CTA's half of the filter once, then its plane strips =====
        if (elect_one()) {
            if (rank == 0) mbar_expect_tx(w_full, 2u * (uint32_t)(9 * planes * 4096));
            for (int t = 0; t < 9; ++t) {
                const int ftap = p.taps[t] >> 24;
                for (int pl = 0; pl < planes; ++pl)
                    tma_load_2d_pair(b_smem + (t * 2 + pl) * 4096, pl ? &map_b_lo : &map_b_hi, w_full, 0, ftap * kPairBN + (int)rank * 32);
            }
        }
        __syncwarp();
        int n_load = 0;      // plane loads issued so far: load j goes to slot j % 3
        for (int pr = cluster_id; pr < p.num_pairs && !(p.debug & 8); pr += num_clusters) {
            const int g = pr / p.tiles_per_frame, i = pr - g * p.tiles_per_frame;
            int f = 2 * g + (int)rank;
            if (f >= p.frames) f = p.frames - 1;             // odd frame count: the last pair's second CTA recomputes a tile and stores nothing
            const int r0 = (128 * i) / p.Wp;
            for (int pl = 0; pl < planes; ++pl, ++n_load) {
                const int slot = n_load % p.slots;
                mbar_wait(&a_empty[slot], ((n_load / p.slots) & 1) ^ 1);
                if (elect_one()) {
                    if (rank == 0) mbar_expect_tx(&a_full[slot], 2u * (uint32_t)p.slot_bytes_tx);
                    tma_load_5d_pair(ring + slot * p.slot_bytes, pl ? &map_a_lo : &map_a_hi, &a_full[slot], 0, -1, r0 - 1, f % p.T, f / p.T);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1 && rank == 0) {
        // ===== MMA issuer (leader CTA).  The WHOLE warp walks the loop (waits included) and one elected lane issues: every operand of
        //       the MMAs is then provably warp-uniform and lives in uniform registers -- issued from `if (lane == 0)` each UTCHMMA sat in
        //       an ELECT / R2UR.BROADCAST waterfall over spilled 64-bit descriptors (119 cycles per MMA measured for 32 of math) =====
        constexpr uint32_t idesc = make_idesc_bf16(256, kPairBN, 0, 0), idesc2 = make_idesc_bf16(256, 2 * kPairBN, 0, 0);
        constexpr uint32_t desc_hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);      // SBO 1024 B, version 1, SWIZZLE_128B
        const uint32_t ring_lo = ((smem_u32(ring) >> 4) & 0x3FFF) | (1u << 16);            // start address, LBO 16 B
        const uint32_t b_lo = ((smem_u32(b_smem) >> 4) & 0x3FFF) | (1u << 16);
        mbar_wait(w_full, 0);
        tc_fence_after();
        int n_use = 0, it = 0;
        for (int pr = cluster_id; pr < p.num_pairs; pr += num_clusters, ++it) {
            const int i = pr % p.tiles_per_frame;
            const int s0 = 128 * i;
            const int l0 = s0 - (s0 / p.Wp - 1) * p.Wp;           // strip pixel of MMA row 0 inside the slot (tap (0, 0))
            const int buf = it & 1;
            mbar_wait(&tmem_empty[buf], ((it >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t acc = tmem_base + buf * kPairAcc;
#pragma unroll
            for (int pl = 0; pl < planes; ++pl, ++n_use) {
                const int slot = n_use % p.slots;
                if (!(p.debug & 8)) mbar_wait(&a_full[slot], (n_use / p.slots) & 1);
                tc_fence_after();
                const uint32_t a_lo32 = ring_lo + (uint32_t)((slot * p.slot_bytes) >> 4) + (uint32_t)(l0 * 8);      // 128 bytes per strip pixel = 8 units
                if (elect_one()) {
#pragma unroll
                    for (int t = 0; t < 9; ++t) {
                        const uint32_t at = a_lo32 + (uint32_t)p.tap_shift8[t];
                        const uint32_t bt = b_lo + (uint32_t)(t * ((2 * 4096) >> 4));
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t da = ((uint64_t)desc_hi << 32) | (at + 2 * k), db = ((uint64_t)desc_hi << 32) | (bt + 2 * k);
                            if (X3 && pl == 0) umma_bf16_pair(acc, da, db, idesc2, (t | k) != 0);           // hi * [hi ; lo] -> columns 0..127
                            else umma_bf16_pair(acc + 128, da, db, idesc, (t | k) != 0);                    // lo * hi (or the single plane) -> 128..191
                        }
                    }
                    umma_commit_pair(&a_empty[slot]);                              // both producers may refill the slot once these MMAs have read it
                    if (pl == planes - 1) umma_commit_pair(&tmem_full[buf]);       // accumulator complete: both CTAs' epilogues
                }
                __syncwarp();
            }
        }
    } else if (warp >= 2) {
        // ===== epilogue (both CTAs): the same transposing epilogue as conv_tc_kernel -- TMEM -> registers -> per-warp staging -> coalesced
        //       64-byte segments, optional addend, BatchNorm statistics / fused BatchNorm-backward sums kept in registers =====
        // No shared-memory transpose: tcgen05.ld.16x256b hands lane l the column pairs 8 j + 2 (l % 4) + {0, 1} of rows l / 4 and l / 4 + 8, so the four
        // lanes of a row cover 32 contiguous bytes (one full sector) per column group and a warp store touches 8 rows -- like the
        // staged float4 stores did -- while the staging tile's 64 KB per tile of shared-memory traffic (the kernel is bound by the
        // 128 B/clk the MMAs read operands at) and its two warp syncs per half-chunk are gone.
        const int q = warp & 3;                     // TMEM lane quarter this warp may access
        const int hsel = (warp - 2) >> 2;           // which 32 columns of the 64
        const int r8 = lane >> 2, c2 = (lane & 3) * 2;
        double* const acc_out = stats ? stats : fuse.sums;
        if (fuse.z) {
            const int t = threadIdx.x - 64;
            if (t < kPairBN) s_par[t] = make_float4(__ldg(fuse.mean + t), __ldg(fuse.invstd + t), __ldg(fuse.gamma + t), __ldg(fuse.beta + t));
            asm volatile("bar.sync 1, 256;" ::: "memory");
        }
        float run1[8], run2[8];                     // running sums of this lane's 8 columns: [column group j][pair element]
#pragma unroll
        for (int t = 0; t < 8; ++t) run1[t] = run2[t] = 0.f;
        int it = 0;
        for (int pr = cluster_id; pr < p.num_pairs; pr += num_clusters, ++it) {
            const int buf = it & 1;
            // output pixel of accumulator row q * 32 + lane of pair tile `t` (the host checks frames * H * W < 2^31); ~0: none
            auto row_of = [&](int t) -> uint32_t {
                int i, c;
                const int g = fdivmod(t, p.d_tpf, i);
                const int f = 2 * g + (int)rank;
                const int s = 128 * i + q * 32 + lane;      // strip pixel
                const int h = fdivmod(s, p.d_wp, c);
                const bool valid = t < p.num_pairs && f < p.frames && h < p.H && c >= 1 && c <= p.W;
                return valid ? (uint32_t)((f * p.H + h) * p.W + (c - 1)) : ~0u;
            };
            const uint32_t my_row = row_of(pr);
            if (addend || fuse.z) {      // the rows of the NEXT tile on their way to L2 (this warp's 32 columns of a row are one 128-byte line)
                const uint32_t nxt = row_of(pr + num_clusters);
                if (nxt != ~0u) {
                    if (addend) asm volatile("prefetch.global.L2 [%0];" ::"l"(addend + (size_t)nxt * kPairBN + hsel * 32));
                    if (fuse.z) asm volatile("prefetch.global.L2 [%0];" ::"l"(fuse.z + (size_t)nxt * kPairBN + hsel * 32));
                }
            }
            uint32_t rows4[4];                              // this lane's rows: r8, r8 + 8, r8 + 16, r8 + 24 of the quarter
#pragma unroll
            for (int j = 0; j < 4; ++j) rows4[j] = __shfl_sync(0xffffffffu, my_row, j * 8 + r8);
            // residual addend / z of the fused BatchNorm backward: this lane's 8 column pairs of the first 16-row half are requested before
            // the wait for the accumulator, those of the second half before the first half is processed (HBM latency off the per-tile chain)
            const int col0 = hsel * 32 + c2;
            float2 ad[2][8], zz[2][8];                      // [half][2 j + row select]
            auto request = [&](int hf) {
                if (addend) {
#pragma unroll
                    for (int v = 0; v < 8; ++v) {
                        const uint32_t row = rows4[hf * 2 + (v & 1)];
                        ad[hf][v] = row != ~0u ? __ldg(reinterpret_cast<const float2*>(addend + (size_t)row * kPairBN + col0 + 8 * (v >> 1))) : make_float2(0.f, 0.f);
                    }
                }
                if (fuse.z) {
#pragma unroll
                    for (int v = 0; v < 8; ++v) {
                        const uint32_t row = rows4[hf * 2 + (v & 1)];
                        zz[hf][v] = row != ~0u ? __ldg(reinterpret_cast<const float2*>(fuse.z + (size_t)row * kPairBN + col0 + 8 * (v >> 1))) : make_float2(0.f, 0.f);
                    }
                }
            };
            request(0);
            mbar_wait_sleep(&tmem_full[buf], (it >> 1) & 1, 128);
            tc_fence_after();
            // all accumulator pieces of this warp's 32 rows x 32 columns go to registers first and the TMEM buffer is handed back at once
            uint32_t ra[2][16];                             // [16-row half][4 j + 2 (row + 8) + pair element]
            {
                const uint32_t tacc = tmem_base + buf * kPairAcc + ((uint32_t)(q * 32) << 16);
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) tmem_ld_16x256b_x4(tacc + ((uint32_t)(hf * 16) << 16) + 128 + hsel * 32, ra[hf]);     // lo * hi (or the single plane)
                if (X3) {
                    uint32_t r1[2][16];
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) tmem_ld_16x256b_x4(tacc + ((uint32_t)(hf * 16) << 16) + hsel * 64 + 32, r1[hf]);  // hi * lo
                    tmem_ld_wait();
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf)
#pragma unroll
                        for (int v = 0; v < 16; ++v) ra[hf][v] = __float_as_uint(__uint_as_float(r1[hf][v]) + __uint_as_float(ra[hf][v]));
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) tmem_ld_16x256b_x4(tacc + ((uint32_t)(hf * 16) << 16) + hsel * 64, r1[hf]);       // hi * hi
                    tmem_ld_wait();
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf)
#pragma unroll
                        for (int v = 0; v < 16; ++v) ra[hf][v] = __float_as_uint(__uint_as_float(r1[hf][v]) + __uint_as_float(ra[hf][v]));
                } else {
                    tmem_ld_wait();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_leader(&tmem_empty[buf]);
            }
            request(1);
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float4 par[2];
                    if (fuse.z) { par[0] = s_par[col0 + 8 * j];  par[1] = s_par[col0 + 8 * j + 1]; }
#pragma unroll
                    for (int rs = 0; rs < 2; ++rs) {
                        const uint32_t row = rows4[hf * 2 + rs];
                        if (row == ~0u || (p.debug & 1)) continue;
                        float o[2] = {__uint_as_float(ra[hf][4 * j + 2 * rs]), __uint_as_float(ra[hf][4 * j + 2 * rs + 1])};
                        if (addend) { o[0] += ad[hf][2 * j + rs].x;  o[1] += ad[hf][2 * j + rs].y; }
                        *reinterpret_cast<float2*>(out + (size_t)row * kPairBN + col0 + 8 * j) = make_float2(o[0], o[1]);
                        if (acc_out) {
                            if (fuse.z) {
                                const float zv[2] = {zz[hf][2 * j + rs].x, zz[hf][2 * j + rs].y};
#pragma unroll
                                for (int e = 0; e < 2; ++e) {
                                    const float xh = (zv[e] - par[e].x) * par[e].y;
                                    const float gg = fmaf(xh, par[e].z, par[e].w) > 0.f ? o[e] : 0.f;
                                    run1[2 * j + e] += gg;
                                    run2[2 * j + e] = fmaf(gg, xh, run2[2 * j + e]);
                                }
                            } else {
#pragma unroll
                                for (int e = 0; e < 2; ++e) {
                                    run1[2 * j + e] += o[e];
                                    run2[2 * j + e] = fmaf(o[e], o[e], run2[2 * j + e]);
                                }
                            }
                        }
                    }
                }
            }
        }
        if (acc_out) {
            float* const s_red = s_stage;           // [4 quarters][2 sums][64]
            asm volatile("bar.sync 1, 256;" ::: "memory");
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                float a = run1[t], b = run2[t];
#pragma unroll
                for (int o = 4; o <= 16; o <<= 1) {         // over the 8 lanes (rows) that own the same columns
                    a += __shfl_xor_sync(0xffffffffu, a, o);
                    b += __shfl_xor_sync(0xffffffffu, b, o);
                }
                if (lane < 4) {
                    const int col = hsel * 32 + 8 * (t >> 1) + c2 + (t & 1);
                    s_red[(q * 2 + 0) * kPairBN + col] = a;
                    s_red[(q * 2 + 1) * kPairBN + col] = b;
                }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const int t = threadIdx.x - 64;         // 0..255: (which sum, column) for t < 128
            if (t < 2 * kPairBN) {
                const int which = t / kPairBN, col = t - which * kPairBN;
                const float tot = s_red[(0 * 2 + which) * kPairBN + col] + s_red[(1 * 2 + which) * kPairBN + col] + s_red[(2 * 2 + which) * kPairBN + col] +
                                  s_red[(3 * 2 + which) * kPairBN + col];
                atomicAdd(acc_out + (size_t)which * kPairBN + col, (double)tot);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                     // no CTA frees its half of the pair's TMEM while the other may still use it
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

// the layers this kernel takes: 1 x 3 x 3, stride 1, padding (0, 1, 1), 64 -> 64 channels, both bf16 planes or hi only
static bool conv_pair_plan(const avid_conv_shape_t* s, PairParams& p) {
    static const bool enabled = [] { const char* e = getenv("AVID_CONV_PAIR"); return !(e && atoi(e) == 0); }();
    if (!enabled) return false;
    if (!(s->kt == 1 && s->kh == 3 && s->kw == 3 && s->st == 1 && s->sh == 1 && s->sw == 1 && s->pt == 0 && s->ph == 1 && s->pw == 1 && s->ci == 64 &&
          s->co == 64))
        return false;
    p.frames = s->n * s->ti;  p.T = s->ti;  p.H = s->hi;  p.W = s->wi;  p.Wp = s->wi + 2;
    if (p.Wp > 256 || p.frames < 2 || (int64_t)p.frames * p.H * p.W >= (int64_t)1 << 31) return false;
    const int rows = (p.Wp + 126) / p.Wp + 3;              // input rows a 128-pixel strip tile can touch
    const int slot_tx = rows * p.Wp * 128;
    p.slot_bytes = (slot_tx + 1023) & ~1023;
    constexpr int kFixed = 1024 + kPairBBytes + kPairBN * 16 + kPairStgBytes + 256;       // everything but the ring
    p.slots = kFixed + kPairSlots * p.slot_bytes <= 232448 ? kPairSlots : 2;
    if (kFixed + p.slots * p.slot_bytes > 232448 || rows > 256) return false;
    p.slot_bytes_tx = slot_tx;
    p.tiles_per_frame = (p.H * p.Wp + 127) / 128;
    p.num_pairs = ((p.frames + 1) / 2) * p.tiles_per_frame;
    p.d_tpf = make_fastdiv(p.tiles_per_frame);  p.d_wp = make_fastdiv(p.Wp);
    return true;
}

bool conv_pair_supported(const avid_conv_shape_t* s) {
    PairParams p;
    return conv_pair_plan(s, p);
}

int conv_pair_run(const avid_conv_shape_t* s, int dgrad, const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, const float* addend,
                  float* out, double* stats, const BnBwdFuse& fuse, cudaStream_t st) {
    PairParams p;
    if (!conv_pair_plan(s, p)) return AVID_EUNSUPPORTED;
    const int rows = p.slot_bytes_tx / (p.Wp * 128);
    p.x3 = a_lo != nullptr;
    { const char* dbg = getenv("AVID_PAIR_DEBUG"); p.debug = dbg ? atoi(dbg) : 0; }
    for (int dh = 0; dh < 3; ++dh)
        for (int dw = 0; dw < 3; ++dw) {
            // forward: output (h, w) reads input (h + dh - 1, w + dw - 1) with filter tap (dh, dw); input gradient: din(h, w) reads
            // dout(h + dh - 1, w + dw - 1) with the mirrored tap (2 - dh, 2 - dw)
            const int ftap = dgrad ? (2 - dh) * 3 + (2 - dw) : dh * 3 + dw;
            p.taps[dh * 3 + dw] = (uint32_t)dw | ((uint32_t)dh << 8) | ((uint32_t)ftap << 24);
            p.tap_shift8[dh * 3 + dw] = ((dh - 1) * p.Wp + (dw - 1)) * 8;
        }
    const TensorMapApi& api = tensor_map_api();
    if (!api.ok) { set_error("conv_pair: cuTensorMapEncode* driver entry points unavailable"); return AVID_ECUDA; }
    CUtensorMap maps[4];
    const void* planes_a[2] = {a_hi, a_lo ? a_lo : a_hi};
    const void* planes_b[2] = {b_hi, b_lo ? b_lo : b_hi};
    for (int pl = 0; pl < 2; ++pl) {
        // activations [N][T][H][W][64] bf16: one box = the whole strip of a tile [rows][Wp][64 channels], zero-filled outside the frame
        cuuint64_t dims[5] = {64, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.T, (cuuint64_t)s->n};
        cuuint64_t strides[4] = {128, (cuuint64_t)p.W * 128, (cuuint64_t)p.H * p.W * 128, (cuuint64_t)p.T * p.H * p.W * 128};
        cuuint32_t box[5] = {64, (cuuint32_t)p.Wp, (cuuint32_t)rows, 1, 1};
        cuuint32_t es[5] = {1, 1, 1, 1, 1};
        CUresult r = api.tiled(&maps[pl], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(planes_a[pl]), dims, strides, box, es,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("conv_pair: cuTensorMapEncodeTiled (activations) failed (%d)", (int)r); return AVID_ECUDA; }
        // filter planes [9 taps * 64 rows][64 k]: boxes of 32 rows (this CTA's half of N)
        cuuint64_t bdims[2] = {64, 9 * 64};
        cuuint64_t bstr[1] = {128};
        cuuint32_t bbox[2] = {64, 32}, bes[2] = {1, 1};
        r = api.tiled(&maps[2 + pl], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(planes_b[pl]), bdims, bstr, bbox, bes, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("conv_pair: cuTensorMapEncodeTiled (filter) failed (%d)", (int)r); return AVID_ECUDA; }
    }
    const int smem = 1024 + kPairBBytes + p.slots * p.slot_bytes + kPairBN * 16 + kPairStgBytes + 256;
    auto kern = p.x3 ? conv_pair_kernel<true> : conv_pair_kernel<false>;
    static bool configured[2] = {false, false};
    if (!configured[p.x3]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
        if (e != cudaSuccess) { set_error("conv_pair: smem attribute: %s", cudaGetErrorString(e)); return AVID_ECUDA; }
        configured[p.x3] = true;
    }
    if (smem > 232448) return AVID_EUNSUPPORTED;
    int clusters = kNumSMs / 2;
    if (clusters > p.num_pairs) clusters = p.num_pairs;
    if (p.debug & 256) printf("conv_pair: %d clusters, smem %d, slot %d, %d tiles per frame\n", clusters, smem, p.slot_bytes, p.tiles_per_frame);
    kern<<<2 * clusters, kPairThreads, smem, st>>>(maps[0], maps[1], maps[2], maps[3], p, addend, out, stats, fuse);
    return check_launch("conv_pair_kernel");
}

}  // namespace avid
