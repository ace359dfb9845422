// CMA positive mining on the tensor cores (include/avid_b200.h, avid_cma_topk_scan_tc / _rescore / _certify).
//
// CMASampler.sample_instance (criterions/avid_cma.py:42-73) ranks all N candidates of every query by
// combine(<V_c, v_q>, <A_c, a_q>) in fp32 and keeps the top pos_k + 1.  The ranking only has to be exact among the best few
// dozen candidates, so the N x N similarity work is done ONCE in fp16 on tcgen05 (fp32 accumulation in TMEM) to produce an
// approximate top-64 per query, and the 64 survivors are then re-scored in exact fp32:
//
//   scan_tc   persistent CTAs, one 128-query tile at a time.  The query tile (both modalities, K = 128 as two 64-wide
//             K-major SW128 blocks) stays resident in shared memory; candidate tiles of 128 rows stream through a TMA ring,
//             two accumulators (video | audio, 128 columns each) are double-buffered in TMEM so the top-k epilogue of
//             candidate tile i overlaps the MMAs of tile i + 1.  Epilogue: one thread per query row reads its 128
//             similarities with tcgen05.ld, combines the modalities (min / max) and keeps the running top-64
//             (slot-major list in shared memory, threshold in a register; replace-the-minimum insertion).
//   rescore   one warp per query: exact fp32 dot products against the (<= 64) listed candidate rows of the current shard.
//   certify   every candidate outside a full list has approximate similarity <= a_min (the list minimum) and therefore exact
//             similarity <= a_min + eps, eps = 2^-10 being a rigorous bound of the fp16 input-rounding error of a dot product of
//             unit vectors.  If the (pos_k+1)-th best EXACT score of the list exceeds a_min + eps the exact top-(pos_k+1) is
//             proven to be inside the list; the list is then reordered (exact score descending, index ascending) for
//             avid_cma_topk_finish.  Queries without a certificate are reported to the caller, which re-mines them with the
//             fp32 kernel (csrc/cma.cu).  With 240k random unit rows the gap between rank 33 and rank 64 is ~20 eps.
#include <cuda_fp16.h>
#include <math.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace avid {
using namespace tc;

constexpr int kCmaSlots = 64;
constexpr int kCmaD = 128;
constexpr int kCmaBlk = 128 * 64 * 2;        // [128 rows][64 k] fp16 = 16 KB, one SW128 K-major operand block
constexpr int kCmaStages = 4;
constexpr int kCmaThreads = 192;             // warp 0: TMA, warp 1: TMEM alloc + MMA issue, warps 2-5: top-k epilogue
constexpr int kCmaSmem = 4 * kCmaBlk + kCmaStages * kCmaBlk + 2 * kCmaSlots * 128 * 4 + 32 * 128 * 4 + 256 + 1024;

__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {      // .kind::f16, A = B = F16 (format 0), D = F32, K-major
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct CmaTcParams {
    int64_t num_queries, num_cand, cand_begin;
    int mode;                 // 0 consensus (min), 1 union (max), 2 video, 3 audio
    float* top_val;           // [num_queries][64] approximate similarities
    int* top_idx;             // [num_queries][64] candidate rows (global indices), -1 = empty
    float* top_exact;         // [num_queries][64] exact similarities (written by rescore); reset for replaced entries
};

__global__ void __launch_bounds__(kCmaThreads, 1)
cma_scan_tc_kernel(const __grid_constant__ CUtensorMap map_qv, const __grid_constant__ CUtensorMap map_qa,
                   const __grid_constant__ CUtensorMap map_cv, const __grid_constant__ CUtensorMap map_ca, const CmaTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* q_smem = smem;                                  // [modality][k-block] 16 KB each
    uint8_t* ring = smem + 4 * kCmaBlk;                      // [stage] 16 KB
    float* lv = reinterpret_cast<float*>(ring + kCmaStages * kCmaBlk);      // [64 slots][128 query rows]
    int* li = reinterpret_cast<int*>(lv + kCmaSlots * 128);
    float* chunk = reinterpret_cast<float*>(li + kCmaSlots * 128);          // [32 columns][128 query rows]: chunks with list hits
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(chunk + 32 * 128);
    uint64_t* empty_bar = full_bar + kCmaStages;
    uint64_t* q_full = empty_bar + kCmaStages;
    uint64_t* q_free = q_full + 1;
    uint64_t* tmem_full = q_free + 1;         // [2]
    uint64_t* tmem_empty = tmem_full + 2;     // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;      // provably warp-uniform
    const bool use_v = p.mode != 3, use_a = p.mode != 2;
    const int nmod = (use_v ? 1 : 0) + (use_a ? 1 : 0);
    const int num_qtiles = (int)((p.num_queries + 127) / 128);
    const int num_ctiles = (int)((p.num_cand + 127) / 128);

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&map_qv);
        prefetch_tensormap(&map_qa);
        prefetch_tensormap(&map_cv);
        prefetch_tensormap(&map_ca);
        for (int s = 0; s < kCmaStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(q_full, 1);
        mbar_init(q_free, 1);
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tmem_full[b], 1);
            mbar_init(&tmem_empty[b], 4);      // one arrive per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);     // 2 buffers x (video | audio) x 128 columns
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    // producer and MMA issuer: whole warps walk the loops, one elected lane issues (uniform-register operands, no per-instruction
    // ELECT / R2UR waterfall -- see conv_tc_kernel)
    if (warp == 0) {
        // ===== TMA producer =====
        int stage = 0, phase = 0, it = 0;
        for (int qt = blockIdx.x; qt < num_qtiles; qt += gridDim.x, ++it) {
            if (it > 0) mbar_wait(q_free, (it - 1) & 1);            // the MMAs of the previous query tile have read it
            if (elect_one()) {
                mbar_expect_tx(q_full, (uint32_t)(nmod * 2 * kCmaBlk));
                for (int m = 0; m < 2; ++m) {
                    if (!(m ? use_a : use_v)) continue;
                    for (int kb = 0; kb < 2; ++kb) tma_load_2d(q_smem + (m * 2 + kb) * kCmaBlk, m ? &map_qa : &map_qv, q_full, kb * 64, qt * 128);
                }
            }
            __syncwarp();
            for (int ct = 0; ct < num_ctiles; ++ct)
                for (int m = 0; m < 2; ++m) {
                    if (!(m ? use_a : use_v)) continue;
                    for (int kb = 0; kb < 2; ++kb) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        if (elect_one()) {
                            mbar_expect_tx(&full_bar[stage], (uint32_t)kCmaBlk);
                            tma_load_2d(ring + stage * kCmaBlk, m ? &map_ca : &map_cv, &full_bar[stage], kb * 64, ct * 128);
                        }
                        __syncwarp();
                        if (++stage == kCmaStages) { stage = 0; phase ^= 1; }
                    }
                }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: D[query][candidate] += Q[query][k] * C[candidate][k]^T, M = N = 128, K = 16 per instruction =====
        constexpr uint32_t idesc = make_idesc_f16(128, 128);
        const uint64_t desc0 = make_smem_desc_sw128(0, 16, 1024);
        const uint64_t q_desc = desc0 + (smem_u32(q_smem) >> 4);
        const uint64_t r_desc = desc0 + (smem_u32(ring) >> 4);
        int stage = 0, phase = 0, it = 0;
        uint32_t cnt = 0;
        for (int qt = blockIdx.x; qt < num_qtiles; qt += gridDim.x, ++it) {
            mbar_wait(q_full, it & 1);
            tc_fence_after();
            for (int ct = 0; ct < num_ctiles; ++ct, ++cnt) {
                const uint32_t buf = cnt & 1;
                mbar_wait(&tmem_empty[buf], ((cnt >> 1) & 1) ^ 1);      // the epilogue has drained this accumulator pair
                tc_fence_after();
                int slot = 0;
                for (int m = 0; m < 2; ++m) {
                    if (!(m ? use_a : use_v)) continue;
                    const uint32_t acc = tmem_base + buf * 256 + slot * 128;
                    ++slot;
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb) {
                        mbar_wait(&full_bar[stage], phase);
                        tc_fence_after();
                        const uint64_t a = q_desc + (uint32_t)(((m * 2 + kb) * kCmaBlk) >> 4);
                        const uint64_t b = r_desc + (uint32_t)((stage * kCmaBlk) >> 4);
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) umma_bf16(acc, a + 2 * k, b + 2 * k, idesc, (kb | k) != 0);
                            umma_commit(&empty_bar[stage]);
                        }
                        __syncwarp();
                        if (++stage == kCmaStages) { stage = 0; phase ^= 1; }
                    }
                }
                if (elect_one()) {
                    umma_commit(&tmem_full[buf]);
                    if (ct == num_ctiles - 1) umma_commit(q_free);
                }
                __syncwarp();
            }
        }
    } else if (warp >= 2) {
        // ===== epilogue: one thread per query row keeps that query's running top-64 =====
        const int q = warp & 3;
        const int r = q * 32 + lane;
        uint32_t cnt = 0;
        for (int qt = blockIdx.x; qt < num_qtiles; qt += gridDim.x) {
            const int64_t qrow = (int64_t)qt * 128 + r;
            const bool valid_q = qrow < p.num_queries;
            // the list minimum is tracked per group of 8 slots (minimum and its slot in registers): replacing the minimum rescans
            // 8 slots instead of 64 -- with ~64 (1 + ln(N / 64)) replacements per query and 32 independent lists per warp the
            // rescans, not the 128 compares per tile, dominate the epilogue
            float gmin[8];
            int gpos[8];
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                gmin[g] = INFINITY;
                gpos[g] = g * 8;
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int s = g * 8 + e;
                    float v = -INFINITY;
                    int ix = -1;
                    if (valid_q) {
                        v = p.top_val[qrow * kCmaSlots + s];
                        ix = p.top_idx[qrow * kCmaSlots + s];
                    }
                    lv[s * 128 + r] = v;
                    li[s * 128 + r] = ix;
                    if (v < gmin[g]) { gmin[g] = v; gpos[g] = s; }
                }
            }
            float thr = gmin[0];
            int min_g = 0;
#pragma unroll
            for (int g = 1; g < 8; ++g)
                if (gmin[g] < thr) { thr = gmin[g]; min_g = g; }
            if (!valid_q) thr = INFINITY;            // rows past the end never insert
            for (int ct = 0; ct < num_ctiles; ++ct, ++cnt) {
                const uint32_t buf = cnt & 1;
                mbar_wait_sleep(&tmem_full[buf], (cnt >> 1) & 1, 64);
                tc_fence_after();
                const uint32_t t0 = tmem_base + buf * 256 + ((uint32_t)(q * 32) << 16);
                const int c_base = ct * 128;
#pragma unroll 1
                for (int j = 0; j < 4; ++j) {
                    uint32_t sv[32], sa[32];
                    tmem_ld_32x32b_x32(t0 + j * 32, sv);
                    if (nmod == 2) tmem_ld_32x32b_x32(t0 + 128 + j * 32, sa);
                    tmem_ld_wait();
                    if (j == 3) {                    // both accumulators are in registers: the next-but-one tile may overwrite them
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tmem_empty[buf]);
                    }
                    // fast path: combine the modalities and compare the chunk maximum with the list minimum; a chunk in which no
                    // lane of the warp has a hit (the common case once the lists have warmed up) costs ~2 instructions per column
                    float mx = -INFINITY;
                    const int c0 = c_base + j * 32;
                    if (nmod == 2) {
                        if (p.mode == 0) {
#pragma unroll
                            for (int v = 0; v < 32; ++v) sv[v] = __float_as_uint(fminf(__uint_as_float(sv[v]), __uint_as_float(sa[v])));
                        } else {
#pragma unroll
                            for (int v = 0; v < 32; ++v) sv[v] = __float_as_uint(fmaxf(__uint_as_float(sv[v]), __uint_as_float(sa[v])));
                        }
                    }
                    if (c0 + 32 > (int)p.num_cand) {        // ragged last tile: rows past the end are zero-filled by TMA, not candidates
#pragma unroll
                        for (int v = 0; v < 32; ++v)
                            if (c0 + v >= (int)p.num_cand) sv[v] = __float_as_uint(-INFINITY);
                    }
#pragma unroll
                    for (int v = 0; v < 32; ++v) mx = fmaxf(mx, __uint_as_float(sv[v]));
                    if (!__any_sync(0xffffffffu, mx > thr)) continue;
                    // slow path (ONE copy of the list code): the hit lanes park their chunk in shared memory and walk their hits
                    uint32_t mask = 0;
                    if (mx > thr) {
#pragma unroll
                        for (int v = 0; v < 32; ++v) {
                            chunk[v * 128 + r] = __uint_as_float(sv[v]);
                            if (__uint_as_float(sv[v]) > thr) mask |= 1u << v;
                        }
                    }
                    while (mask) {
                        const int v = __ffs(mask) - 1;
                        mask &= mask - 1;
                        const float s = chunk[v * 128 + r];
                        if (!(s > thr)) continue;        // the minimum has risen since the mask was built
                        int slot = 0;
#pragma unroll
                        for (int g = 0; g < 8; ++g)
                            if (g == min_g) slot = gpos[g];
                        lv[slot * 128 + r] = s;
                        li[slot * 128 + r] = (int)p.cand_begin + c0 + v;
                        // new minimum of the group that held the list minimum
                        float x[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) x[e] = lv[(min_g * 8 + e) * 128 + r];
                        float m = x[0];
                        int mp = 0;
#pragma unroll
                        for (int e = 1; e < 8; ++e)
                            if (x[e] < m) { m = x[e]; mp = e; }
#pragma unroll
                        for (int g = 0; g < 8; ++g)
                            if (g == min_g) { gmin[g] = m; gpos[g] = g * 8 + mp; }
                        thr = gmin[0];
                        min_g = 0;
#pragma unroll
                        for (int g = 1; g < 8; ++g)
                            if (gmin[g] < thr) { thr = gmin[g]; min_g = g; }
                    }
                }
            }
            if (valid_q)
                for (int s = 0; s < kCmaSlots; ++s) {
                    const int ix = li[s * 128 + r];
                    if (ix != p.top_idx[qrow * kCmaSlots + s]) p.top_exact[qrow * kCmaSlots + s] = -INFINITY;      // a new entry: not re-scored yet
                    p.top_val[qrow * kCmaSlots + s] = lv[s * 128 + r];
                    p.top_idx[qrow * kCmaSlots + s] = ix;
                }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

__global__ void __launch_bounds__(256) cma_to_half_kernel(const float* __restrict__ x, __half* __restrict__ out, int64_t n4) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
        __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
        uint2 o;
        o.x = *reinterpret_cast<uint32_t*>(&a);
        o.y = *reinterpret_cast<uint32_t*>(&b);
        reinterpret_cast<uint2*>(out)[i] = o;
    }
}

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// one warp per query: exact fp32 similarities of the listed candidates that lie in [cand_begin, cand_begin + num_cand)
__global__ void __launch_bounds__(256) cma_rescore_kernel(const float* __restrict__ q_video, const float* __restrict__ q_audio, int64_t num_queries,
                                                          const float* __restrict__ c_video, const float* __restrict__ c_audio, int64_t cand_begin,
                                                          int64_t num_cand, int mode, const int* __restrict__ top_idx, float* __restrict__ top_exact) {
    const int64_t qrow = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (qrow >= num_queries) return;
    const bool use_v = mode != 3, use_a = mode != 2;
    float4 qv = make_float4(0.f, 0.f, 0.f, 0.f), qa = qv;
    if (use_v) qv = __ldg(reinterpret_cast<const float4*>(q_video + qrow * kCmaD) + lane);
    if (use_a) qa = __ldg(reinterpret_cast<const float4*>(q_audio + qrow * kCmaD) + lane);
    const int ix0 = top_idx[qrow * kCmaSlots + lane], ix1 = top_idx[qrow * kCmaSlots + 32 + lane];
    float ex0 = top_exact[qrow * kCmaSlots + lane], ex1 = top_exact[qrow * kCmaSlots + 32 + lane];
    for (int e0 = 0; e0 < kCmaSlots; e0 += 4) {
        // four entries per step: their row loads are independent
        float4 cv[4], ca[4];
        bool in[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = e0 + u;
            const int ix = __shfl_sync(0xffffffffu, e < 32 ? ix0 : ix1, e & 31);
            const int64_t row = (int64_t)ix - cand_begin;
            in[u] = ix >= 0 && row >= 0 && row < num_cand;
            cv[u] = ca[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (in[u]) {
                if (use_v) cv[u] = __ldg(reinterpret_cast<const float4*>(c_video + row * kCmaD) + lane);
                if (use_a) ca[u] = __ldg(reinterpret_cast<const float4*>(c_audio + row * kCmaD) + lane);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (!in[u]) continue;            // warp-uniform
            const int e = e0 + u;
            const float dv = warp_sum_f(fmaf(cv[u].w, qv.w, fmaf(cv[u].z, qv.z, fmaf(cv[u].y, qv.y, cv[u].x * qv.x))));
            const float da = warp_sum_f(fmaf(ca[u].w, qa.w, fmaf(ca[u].z, qa.z, fmaf(ca[u].y, qa.y, ca[u].x * qa.x))));
            const float s = mode == 0 ? fminf(dv, da) : (mode == 1 ? fmaxf(dv, da) : (mode == 2 ? dv : da));
            if ((e & 31) == lane) {
                if (e < 32) ex0 = s; else ex1 = s;
            }
        }
    }
    top_exact[qrow * kCmaSlots + lane] = ex0;
    top_exact[qrow * kCmaSlots + 32 + lane] = ex1;
}

// one warp per query: certificate + selection of the exact top-(pos_k+1) into slots 0..pos_k (exact descending, index ascending)
__global__ void __launch_bounds__(256) cma_certify_kernel(int64_t num_queries, int pos_k, float eps, float* __restrict__ top_val,
                                                          int* __restrict__ top_idx, const float* __restrict__ top_exact, int* fail_count,
                                                          int* fail_list) {
    const int64_t qrow = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (qrow >= num_queries) return;
    float a[2] = {top_val[qrow * kCmaSlots + lane], top_val[qrow * kCmaSlots + 32 + lane]};
    int ix[2] = {top_idx[qrow * kCmaSlots + lane], top_idx[qrow * kCmaSlots + 32 + lane]};
    float ex[2] = {top_exact[qrow * kCmaSlots + lane], top_exact[qrow * kCmaSlots + 32 + lane]};
    // list minimum of the approximate scores (a_min) and whether the list is full
    float amin = fminf(ix[0] >= 0 ? a[0] : INFINITY, ix[1] >= 0 ? a[1] : INFINITY);
    int filled = (ix[0] >= 0) + (ix[1] >= 0);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        amin = fminf(amin, __shfl_xor_sync(0xffffffffu, amin, o));
        filled += __shfl_xor_sync(0xffffffffu, filled, o);
    }
    if (ix[0] < 0) ex[0] = -INFINITY;
    if (ix[1] < 0) ex[1] = -INFINITY;
    const int n = pos_k + 1;
    float out_v = -INFINITY;      // lane s (< 32) / s - 32 holds output slot s
    int out_i = -1;
    float out_v2 = -INFINITY;
    int out_i2 = -1;
    float last = -INFINITY;
    for (int s = 0; s < n; ++s) {
        // arg max over the remaining entries of (exact, -index)
        int which = ex[0] > ex[1] || (ex[0] == ex[1] && (unsigned)ix[0] <= (unsigned)ix[1]) ? 0 : 1;
        float bv = ex[which];
        int bi = ix[which], bl = lane * 2 + which;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            const int ol = __shfl_xor_sync(0xffffffffu, bl, o);
            if (ov > bv || (ov == bv && (unsigned)oi < (unsigned)bi)) { bv = ov; bi = oi; bl = ol; }
        }
        if ((bl >> 1) == lane) { ex[bl & 1] = -INFINITY; ix[bl & 1] = -1; }
        if (s < 32) { if (lane == s) { out_v = bv; out_i = bv == -INFINITY ? -1 : bi; } }
        else if (lane == s - 32) { out_v2 = bv; out_i2 = bv == -INFINITY ? -1 : bi; }
        last = bv;
    }
    const bool certified = filled < kCmaSlots || last > amin + eps;
    top_val[qrow * kCmaSlots + lane] = out_v;
    top_idx[qrow * kCmaSlots + lane] = out_i;
    top_val[qrow * kCmaSlots + 32 + lane] = out_v2;
    top_idx[qrow * kCmaSlots + 32 + lane] = out_i2;
    if (!certified && lane == 0 && fail_count) {
        const int k = atomicAdd(fail_count, 1);
        if (fail_list) fail_list[k] = (int)qrow;
    }
}

static int encode_f16_rows(CUtensorMap* map, const void* base, int64_t rows) {
    const TensorMapApi& api = tensor_map_api();
    if (!api.ok) { set_error("cma_tc: cuTensorMapEncodeTiled driver entry point unavailable"); return AVID_ECUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)kCmaD, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)kCmaD * 2};
    cuuint32_t box[2] = {64, 128};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = api.tiled(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cma_tc: cuTensorMapEncodeTiled failed (%d)", (int)r); return AVID_ECUDA; }
    return AVID_OK;
}

static int cma_tc_split(int64_t num_queries, void* workspace, size_t bytes, float** val, int** idx, float** exact) {
    AVID_REQUIRE(num_queries > 0 && workspace, "cma_topk: bad arguments");
    const size_t need = (size_t)num_queries * kCmaSlots * 12;
    if (bytes < need) {
        set_error("cma_topk: workspace of %zu bytes given, %zu needed", bytes, need);
        return AVID_EWORKSPACE;
    }
    *val = static_cast<float*>(workspace);
    *idx = reinterpret_cast<int*>(*val + (size_t)num_queries * kCmaSlots);
    *exact = reinterpret_cast<float*>(*idx + (size_t)num_queries * kCmaSlots);
    return AVID_OK;
}

}  // namespace avid

using namespace avid;

extern "C" {

int avid_cma_to_half(const float* x, void* out, int64_t n, void* stream) {
    AVID_REQUIRE(x && out && n > 0 && n % 4 == 0, "cma_to_half: n=%lld must be a positive multiple of 4", (long long)n);
    int64_t blocks = (n / 4 + 255) / 256;
    if (blocks > 16 * kNumSMs) blocks = 16 * kNumSMs;
    cma_to_half_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, static_cast<__half*>(out), n / 4);
    return check_launch("cma_to_half_kernel");
}

int avid_cma_topk_scan_tc(const void* q_video_h, const void* q_audio_h, int64_t num_queries, const void* cand_video_h, const void* cand_audio_h,
                          int64_t cand_begin, int64_t num_cand, int32_t mode, void* workspace, size_t workspace_bytes, void* stream) {
    CmaTcParams p;
    int rc = cma_tc_split(num_queries, workspace, workspace_bytes, &p.top_val, &p.top_idx, &p.top_exact);
    if (rc) return rc;
    AVID_REQUIRE(q_video_h && q_audio_h && cand_video_h && cand_audio_h, "cma_topk_scan_tc: NULL pointer");
    AVID_REQUIRE(mode >= 0 && mode <= 3, "cma_topk_scan_tc: unknown mode %d", mode);
    AVID_REQUIRE(num_cand > 0 && cand_begin >= 0 && cand_begin + num_cand < ((int64_t)1 << 31) - 256 && num_queries < ((int64_t)1 << 31) - 256,
                 "cma_topk_scan_tc: bad candidate / query range");
    p.num_queries = num_queries;  p.num_cand = num_cand;  p.cand_begin = cand_begin;  p.mode = mode;
    CUtensorMap maps[4];
    if ((rc = encode_f16_rows(&maps[0], q_video_h, num_queries))) return rc;
    if ((rc = encode_f16_rows(&maps[1], q_audio_h, num_queries))) return rc;
    if ((rc = encode_f16_rows(&maps[2], cand_video_h, num_cand))) return rc;
    if ((rc = encode_f16_rows(&maps[3], cand_audio_h, num_cand))) return rc;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(cma_scan_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kCmaSmem);
        if (e != cudaSuccess) { set_error("cma_topk_scan_tc: smem attribute: %s", cudaGetErrorString(e)); return AVID_ECUDA; }
        configured = true;
    }
    const int64_t qtiles = (num_queries + 127) / 128;
    const unsigned grid = (unsigned)(qtiles < kNumSMs ? qtiles : kNumSMs);
    cma_scan_tc_kernel<<<grid, kCmaThreads, kCmaSmem, static_cast<cudaStream_t>(stream)>>>(maps[0], maps[1], maps[2], maps[3], p);
    return check_launch("cma_scan_tc_kernel");
}

int avid_cma_topk_rescore(const float* q_video, const float* q_audio, int64_t num_queries, const float* cand_video, const float* cand_audio,
                          int64_t cand_begin, int64_t num_cand, int32_t mode, void* workspace, size_t workspace_bytes, void* stream) {
    float *val, *exact;
    int* idx;
    int rc = cma_tc_split(num_queries, workspace, workspace_bytes, &val, &idx, &exact);
    if (rc) return rc;
    AVID_REQUIRE(q_video && q_audio && cand_video && cand_audio && num_cand > 0 && mode >= 0 && mode <= 3, "cma_topk_rescore: bad arguments");
    const int64_t threads = num_queries * 32;
    cma_rescore_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        q_video, q_audio, num_queries, cand_video, cand_audio, cand_begin, num_cand, mode, idx, exact);
    return check_launch("cma_rescore_kernel");
}

int avid_cma_topk_certify(int64_t num_queries, int32_t pos_k, float eps, void* workspace, size_t workspace_bytes, int32_t* fail_count,
                          int32_t* fail_list, void* stream) {
    float *val, *exact;
    int* idx;
    int rc = cma_tc_split(num_queries, workspace, workspace_bytes, &val, &idx, &exact);
    if (rc) return rc;
    AVID_REQUIRE(pos_k > 0 && pos_k < kCmaSlots && eps >= 0.f, "cma_topk_certify: bad arguments");
    const int64_t threads = num_queries * 32;
    cma_certify_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(num_queries, pos_k, eps, val, idx, exact,
                                                                                                       fail_count, fail_list);
    return check_launch("cma_certify_kernel");
}

}  // extern "C"
