// Fused memory-bank gather + score + NCE loss + closed-form gradient (include/avid_b200.h,
// "Criterion").  Replaces the ATen chain of criterions/avid.py:47-80, avid_cma.py:150-209 and
// criterions/nce.py:38-58 of the reference: the (B,K,128) gathers are never materialised, every
// bank row is read from HBM exactly once (one coalesced 512-byte request per warp) and serves the
// forward score, the loss term and the gradient w.r.t. the embedding in the same pass.
//
// Kernel 1  nce_gather_kernel   grid (B, splits) x 128 threads.  A warp streams its contiguous range of rows in quads:
//           groups of 8 lanes own one row each (16 columns per lane, 128-byte coalesced group loads), a ring of 2 quads x
//           2 banks in flight per warp (the next quad is requested before the current one is scored); dot product = 16 FMAs
//           + 3 shuffles, the NCE math runs per group (lane l8 = key l8), the gradient axpy needs no broadcast.  Partial
//           (grad_hat, loss) per split go to the workspace.
//           The LAST split CTA of a query (ticket counter) sums the partials in a fixed order (deterministic) and applies
//           the backward of F.normalize; the last query forms the batch means and the coefficient mix: one launch per step.
//           The B + 1 ticket counters live at the start of the workspace, must be zero on entry and are left zero.
// Kernel 2  nce_reduce_finalize_kernel  grid (B) x 256 threads: the same tail as a separate launch, for the sharded
//           protocol (avid_nce_finalize runs after the all-reduce of the partials).
#include <math.h>
#include <stdlib.h>
#include "common.cuh"

namespace avid {

constexpr int kGatherThreads = 128;
constexpr int kGatherWarps = kGatherThreads / 32;

struct KeyDev {
    int ctx, bank, pos_mode, num_neg;
    float weight;
};

struct NceParams {
    const float* emb[2];
    const int64_t* y;
    const float* bank[2];
    int64_t N, row_begin, row_end;
    int B;
    float inv_mean_batch;
    int K;
    const int64_t* neg_idx;
    uint64_t seed, offset;
    const int32_t* positive_set;
    int pos_k;       // rows of positive_set used (0 when no key has pos_mode 1)
    int num_keys;
    KeyDev keys[AVID_MAX_KEYS];
    const float* Z;  // device scalar; nullptr => scores-only pass (partition function)
    float inv_T;
    int splits, kc;
    float* part_grad;   // [splits][2][B][128]
    float* part_loss;   // [splits][num_keys][B]
    float* scores;      // optional [num_keys][B][1 + pos_k_stride + K]
    int score_pos_k;    // pos_k used for the score layout
    int64_t* neg_idx_out;
    unsigned int* counter;      // [B + 1] tickets: splits done per query, queries done; zero on entry, left zero
    bool need[2][2];    // need[bank][ctx]: some key scores this bank against this context
    unsigned long long* dbg;    // AVID_NCE_DEBUG: 8 %globaltimer stamps per CTA (first 1024 CTAs), else nullptr
    int coef_lane[4];   // pair 2 * bank + ctx: the one key that scores it, or -1 (none, or several: summed over the group)
    bool bank_used[2];
    // tail of the fused kernel: the last split of a query reduces it, the last query forms the batch means
    float* grad_hat[2];         // optional (B,128): reduced gradient w.r.t. the normalised embedding (sharded protocol)
    float* loss_part;           // (num_keys, B) per-query loss terms
    float* grad[2];             // (B,128) gradient w.r.t. the raw embedding (do_finalize)
    float* loss_keys;
    float* loss_total;
    float weights[AVID_MAX_KEYS];
    int do_finalize;
    int* bad_index;             // optional: set to 1 when a y[b] lies outside [0, N)
    // packed per-rank records of a sharded step: query b = (group b / group_batch, row b % group_batch); emb / y / neg_idx of
    // group g start in_group_stride BYTES after those of group g - 1, grad_hat / loss_part out_group_stride bytes (0: dense)
    int group_batch;
    size_t in_group_stride, out_group_stride;
};

__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
    return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
}
__device__ __forceinline__ void axpy4(float4& acc, float c, const float4& r) {
    acc.x = fmaf(c, r.x, acc.x);
    acc.y = fmaf(c, r.y, acc.y);
    acc.z = fmaf(c, r.z, acc.z);
    acc.w = fmaf(c, r.w, acc.w);
}

// Row layout inside a warp: a bank row (128 floats, 512 B) is owned by a GROUP of 8 lanes; lane l8 of the group holds the
// four float4 at columns 4*l8 + 32*j (j = 0..3), so each load instruction of the group covers 128 contiguous bytes and a
// warp instruction fetches 4 rows.  A dot product is 16 FMAs per lane + 3 shuffles for 4 rows at once.
__device__ __forceinline__ float group_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    return v;
}
__device__ __forceinline__ float dot16(const float4 (&a)[4], const float4 (&b)[4]) {
    return (dot4(a[0], b[0]) + dot4(a[1], b[1])) + (dot4(a[2], b[2]) + dot4(a[3], b[3]));      // four independent FMA chains
}

// A warp keeps TWO row quads (4 rows x 2 banks x 512 B = 4 KB each) in flight in registers and walks them as a ring: the next
// request of a slot is issued as soon as its quad is scored.  What the %globaltimer stamps of scripts/debug_nce_timeline.py show
// at B = 64, K = 1024 (384 CTAs, 11 quads per warp): 2.0 us until the first rows are requested, 14 us in the loop -- 67 MB at
// 4.8 TB/s with 12.6 MB in flight, i.e. the loop is bound by the bytes in flight (Little: ~2.6 us per round trip under that
// load), NOT by the order of requests and scoring (requesting both quads together after scoring both gives the same times at
// every K) -- then 6 us of tail (partials, fence, ticket, last split reduces, fence, ticket, last query) and ~6 us between the
// events and the first / last instruction.  Second session of round 2: the per-quad instruction count went from ~290 to ~170
// (key constants folded per lane, (bank, ctx) pairs as template parameter, rcp / lg2 approximations behind a series for
// log1p, packed item descriptors, coefficient by one shuffle), the Philox rounds of the first chunk run while y is in flight,
// the last split CTA finalises a query with one warp per context and no block barrier: 36.9 -> 30.7 us at K = 1024 (same box,
// A/B), 75.8 -> 73.7 us at K = 4096, 213.9 -> 228.2 us at K = 16384 (0.77 -> 0.72 of the measured HBM rate: not understood;
// four resident CTAs per SM under a 128-register cap -- 152 B of spills -- are slower still: 0.67, and 32.8 us at K = 1024).
// Measured and dropped (round 2): landing the rows in a per-warp shared-memory ring instead of registers -- with per-row
// cp.async.bulk copies (request-rate bound: ~35 cycles per 512-byte request and SM; K = 1024: 49 us vs 39 us) and with cp.async
// 16 B per lane (58 us; K = 16384: 0.60 instead of 0.77 of the HBM peak).  Also dropped: a per-lane prefetch.global.L2 of the
// chunk's rows before scoring it (32 scattered lines per instruction: K = 1024 47 us, K = 16384 0.52), and TMA tile::gather4
// requests (4 rows per request, scripts/probes/gather4_tma.cu: correct, but the TMA unit serves ~one 512-byte row per ~60 cycles
// and SM: K = 1024 49 us, K = 16384 0.35 of the HBM peak with a 6-stage ring).
constexpr int kGatherSmem = 0;

__device__ __forceinline__ void stamp(const NceParams& p, int slot) {
    if (p.dbg && threadIdx.x == 0) {
        const unsigned cta = blockIdx.y * gridDim.x + blockIdx.x;
        if (cta < 1024u) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            p.dbg[cta * 8 + slot] = t;
        }
    }
}

// lane-parallel description of 32 consecutive items of a warp's range, packed for the shuffles that hand it to the groups:
// off = row index inside this rank's shard; tag = (kk << 2) | (kind + 1) with kind 0 = negative kk, 1 = self, 2 = positive-set
// entry kk; tag 0 = nothing to score (past the range, or a row another shard holds)
struct ItemDesc {
    int off, tag;
};

struct GatherCtx {
    int b, npos, k_begin, w_end;
    int64_t y;
    const int32_t* pos_row;
    const int64_t* negs;
};

// `bits`: the Philox output of the item's counter when the caller drew it ahead of time (the first chunk, before y arrived)
__device__ __forceinline__ ItemDesc describe_item(const NceParams& p, const GatherCtx& g, int item, const uint32_t* bits = nullptr) {
    ItemDesc d{0, 0};
    if (item < g.w_end) {
        int kind, kk;
        int64_t idx;
        if (item < g.npos) {
            kind = item == 0 ? 1 : 2;
            kk = item - 1;
            idx = item == 0 ? g.y : (g.pos_row ? (int64_t)g.pos_row[item - 1] : -1);
        } else {
            kind = 0;
            kk = g.k_begin + (item - g.npos);
            if (g.negs) idx = g.negs[kk];
            else if (bits) idx = finish_negative(bits[0], bits[1], p.N, g.y, g.pos_row, p.pos_k);
            else idx = draw_negative(p.seed, p.offset, g.b, kk, p.K, p.N, g.y, g.pos_row, p.pos_k);
            if (p.neg_idx_out) p.neg_idx_out[(size_t)g.b * p.K + kk] = idx;
        }
        if (idx >= p.row_begin && idx < p.row_end) {      // rows another shard holds are scored there
            d.off = (int)(idx - p.row_begin);
            d.tag = (kk << 2) | (kind + 1);
        }
    }
    return d;
}

// request quad q (0..7) of the described chunk: group grp of the warp takes item 4 q + grp
__device__ __forceinline__ void request_quad(const NceParams& p, const ItemDesc& d, int q, int grp, int l8, float4 (&rv)[4], float4 (&ra)[4],
                                             int& tag) {
    const int src = 4 * q + grp;
    const int off = __shfl_sync(0xffffffffu, d.off, src);
    tag = __shfl_sync(0xffffffffu, d.tag, src);
    if (tag != 0) {      // otherwise the registers keep the (finite) rows of an earlier quad: they are scored with coefficient 0
        const size_t o = (size_t)off * kD;
        if (p.bank_used[0]) {
            const float4* r = reinterpret_cast<const float4*>(p.bank[0] + o) + l8;
#pragma unroll
            for (int j = 0; j < 4; ++j) rv[j] = ld_stream(r + 8 * j);
        }
        if (p.bank_used[1]) {
            const float4* r = reinterpret_cast<const float4*>(p.bank[1] + o) + l8;
#pragma unroll
            for (int j = 0; j < 4; ++j) ra[j] = ld_stream(r + 8 * j);
        }
    }
}

__device__ __forceinline__ float rcp_approx(float x) {      // x normal and positive here: 1 ulp, no denormal scaling code
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// log(1 + x) for x >= 0 without the libm slow paths: the alternating series below 2^-4 (truncation < 1e-8 relative), lg2.approx
// above (absolute error 1.7e-7 on a value >= 0.06).  The loss terms of the negatives are ~1/K each, where log(1 + x) formed from
// a rounded 1 + x would lose 4 digits.
__device__ __forceinline__ float log1p_pos(float x) {
    const float small = x * fmaf(-x, fmaf(-x, fmaf(-x, fmaf(-x, fmaf(-x, 1.0f / 6.0f, 0.2f), 0.25f), 1.0f / 3.0f), 0.5f), 1.0f);
    return x < 0.0625f ? small : 0.69314718056f * lg2_approx(1.0f + x);
}

// what lane l8 of every group needs of "its" key (key l8), folded once per kernel
struct LaneKey {
    int sel;            // 2 * bank + ctx: which of the four dot products the key scores
    int neg_limit;      // negatives kk < neg_limit count for the key (0: key slot unused)
    int pos_ok;         // bit 0: scores the self positive, bit 1: scores the positive-set entries
    float c, inv_c;     // K_key * Z (nce.py:42-57) and its reciprocal
    float w_neg, w_pos; // weight / mean_batch / T, and the same * -(1 / number of positives the key averages over)
    float inv_p;
};

// NEED >= 0: the (bank, ctx) pairs the keys score, bit 2 * bank + ctx, known at compile time (6 = cross, 9 = self, 15 = joint:
// no uniform branches in the hot loop); NEED < 0: read from the parameters
template <int NEED>
__device__ __forceinline__ bool needs(const NceParams& p, int bank, int ctx) {
    return NEED >= 0 ? ((NEED >> (2 * bank + ctx)) & 1) != 0 : p.need[bank][ctx];
}

// coefficient of the pair `sel` for this group's row: every key's coefficient sits on lane (key) of the group
template <int NEED>
__device__ __forceinline__ float pair_coef(const NceParams& p, int sel, float coef, int my_sel, int lane) {
    if (!needs<NEED>(p, sel >> 1, sel & 1)) return 0.f;
    if (p.coef_lane[sel] >= 0) return __shfl_sync(0xffffffffu, coef, (lane & 24) | p.coef_lane[sel]);      // one key scores the pair
    return group_sum(my_sel == sel ? coef : 0.f);
}

// score one quad: lane l8 of a group evaluates key l8 for the group's row (the transcendental NCE math is lane-parallel over
// the keys), the coefficients are handed to the 8 lanes of the group, the gradient axpy needs no broadcast of the row.
template <int NEED, bool TRAIN>
__device__ __forceinline__ void score_quad(const NceParams& p, const float4 (&rv)[4], const float4 (&ra)[4], int tag,
                                           const float4 (&e_ctx)[2][4], float4 (&acc)[2][4], float& loss_acc, const LaneKey& key,
                                           int my_key, int b, int lane) {
    if (__all_sync(0xffffffffu, tag == 0)) return;      // a quad of rows other shards hold (or the padding of the last quad)
    const float d00 = needs<NEED>(p, 0, 0) ? group_sum(dot16(rv, e_ctx[0])) : 0.f;      // d[bank][ctx]
    const float d01 = needs<NEED>(p, 0, 1) ? group_sum(dot16(rv, e_ctx[1])) : 0.f;
    const float d10 = needs<NEED>(p, 1, 0) ? group_sum(dot16(ra, e_ctx[0])) : 0.f;
    const float d11 = needs<NEED>(p, 1, 1) ? group_sum(dot16(ra, e_ctx[1])) : 0.f;
    const int kind = (tag & 3) - 1, kk = tag >> 2;
    const bool applies = kind == 0 ? kk < key.neg_limit : (kind > 0 && ((key.pos_ok >> (kind - 1)) & 1));
    float coef = 0.f;
    if (applies) {
        float d;
        if (NEED == 6) d = key.sel == 1 ? d01 : d10;
        else if (NEED == 9) d = key.sel == 0 ? d00 : d11;
        else d = key.sel < 2 ? (key.sel == 0 ? d00 : d01) : (key.sel == 2 ? d10 : d11);
        const float s = d * p.inv_T;
        if (p.scores) {      // optional raw scores (always on in the partition-function pass)
            const int slot = kind == 1 ? 0 : (kind == 2 ? 1 + kk : 1 + p.score_pos_k + kk);
            p.scores[((size_t)my_key * p.B + b) * (size_t)(1 + p.score_pos_k + p.K) + slot] = s;
        }
        if (TRAIN) {
            // nce.py:42-57 with c = K_key * Z: a negative contributes log(1 + e / c) and pulls with e / (e + c), a positive
            // contributes log(1 + c / e) (averaged over the positive set) and pushes with c / (e + c)
            const float e = expf(s);
            const float inv_ec = rcp_approx(e + key.c);
            const bool neg = kind == 0;
            const float x = neg ? e * key.inv_c : key.c * rcp_approx(e);
            const float term = log1p_pos(x);
            loss_acc += neg ? term : key.inv_p * term;
            coef = neg ? key.w_neg * (e * inv_ec) : key.w_pos * (key.c * inv_ec);
        }
    }
    if (!TRAIN) return;
    const float cf00 = pair_coef<NEED>(p, 0, coef, key.sel, lane), cf01 = pair_coef<NEED>(p, 1, coef, key.sel, lane);
    const float cf10 = pair_coef<NEED>(p, 2, coef, key.sel, lane), cf11 = pair_coef<NEED>(p, 3, coef, key.sel, lane);
    if (needs<NEED>(p, 0, 0)) {
#pragma unroll
        for (int j = 0; j < 4; ++j) axpy4(acc[0][j], cf00, rv[j]);
    }
    if (needs<NEED>(p, 1, 0)) {
#pragma unroll
        for (int j = 0; j < 4; ++j) axpy4(acc[0][j], cf10, ra[j]);
    }
    if (needs<NEED>(p, 0, 1)) {
#pragma unroll
        for (int j = 0; j < 4; ++j) axpy4(acc[1][j], cf01, rv[j]);
    }
    if (needs<NEED>(p, 1, 1)) {
#pragma unroll
        for (int j = 0; j < 4; ++j) axpy4(acc[1][j], cf11, ra[j]);
    }
}

// x / max(||x||, 1e-12) in place for a row held in the group layout (avid.py:52-53)
__device__ __forceinline__ void normalize_group(float4 (&e)[4]) {
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) ss += dot4(e[j], e[j]);
    const float inv = 1.0f / fmaxf(sqrtf(group_sum(ss)), 1e-12f);
#pragma unroll
    for (int j = 0; j < 4; ++j) e[j] = make_float4(e[j].x * inv, e[j].y * inv, e[j].z * inv, e[j].w * inv);
}

// tail of the training kernel: partial (grad_hat, loss) of this CTA to the workspace; the last split CTA of a query reduces it,
// the last query forms the batch means
__device__ __forceinline__ void finish_query(const NceParams& p, float4 (&acc)[2][4], float loss_acc, int b, int bl, int g_i, int split,
                                             const float* emb0, const float* emb1) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane >> 3, l8 = lane & 7;
    // sum the 4 groups of the warp (lanes with equal l8), then the 4 warps of the CTA in a fixed order
    __shared__ float4 s_acc[kGatherWarps][2][32];
    __shared__ float s_loss[kGatherWarps][AVID_MAX_KEYS];
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float4 v = acc[c][j];
#pragma unroll
            for (int o = 8; o <= 16; o <<= 1) {
                v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
                v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
                v.z += __shfl_xor_sync(0xffffffffu, v.z, o);
                v.w += __shfl_xor_sync(0xffffffffu, v.w, o);
            }
            if (grp == 0) s_acc[warp][c][l8 + 8 * j] = v;
        }
    {
        float v = loss_acc;             // lane l8 holds the terms of key l8: sum over the 4 groups (xor 8, 16)
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        if (lane < AVID_MAX_KEYS) s_loss[warp][lane] = v;
    }
    __syncthreads();
    if (threadIdx.x < 64) {
        const int ctx = threadIdx.x >> 5, l = threadIdx.x & 31;
        float4 t = s_acc[0][ctx][l];
#pragma unroll
        for (int w = 1; w < kGatherWarps; ++w) {
            const float4 o = s_acc[w][ctx][l];
            t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w;
        }
        reinterpret_cast<float4*>(p.part_grad + ((size_t)(split * 2 + ctx) * p.B + b) * kD)[l] = t;
    } else if (threadIdx.x - 64 < p.num_keys) {
        const int k = threadIdx.x - 64;
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < kGatherWarps; ++w) t += s_loss[w][k];
        p.part_loss[((size_t)split * p.num_keys + k) * p.B + b] = t;
    }

    // ---- the last split CTA of this query reduces it (fixed order over splits: deterministic) ----
    __shared__ unsigned int s_ticket;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_ticket = atomicAdd(p.counter + b, 1u);
    __syncthreads();
    stamp(p, 3);
    if (s_ticket != (unsigned)(p.splits - 1)) return;
    __threadfence();
    // warp 0 / 1: context video / audio, lane = 4 embedding columns (no block-wide synchronisation on this path); warp 2: loss terms
    const size_t out_off = (size_t)g_i * p.out_group_stride;
    if (warp < 2) {
        const int ctx = warp;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.do_finalize) x = reinterpret_cast<const float4*>((ctx ? emb1 : emb0) + (size_t)bl * kD)[lane];
        const float4* src = reinterpret_cast<const float4*>(p.part_grad + ((size_t)ctx * p.B + b) * kD) + lane;
        const size_t split_stride = (size_t)2 * p.B * (kD / 4);
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = 0; s < p.splits; ++s) {
            const float4 o = __ldcg(src + s * split_stride);
            g.x += o.x; g.y += o.y; g.z += o.z; g.w += o.w;
        }
        if (p.grad_hat[ctx]) {
            float* dst = reinterpret_cast<float*>(reinterpret_cast<char*>(p.grad_hat[ctx]) + out_off) + (size_t)bl * kD + 4 * lane;
            dst[0] = g.x; dst[1] = g.y; dst[2] = g.z; dst[3] = g.w;
        }
        if (p.do_finalize) {
            // backward of x -> x / max(||x||, eps):  (g - ehat <ehat, g>) / ||x||   (g / eps when clamped)
            const float n = sqrtf(warp_sum(dot4(x, x)));
            const bool clamped = !(n > 1e-12f);
            const float4 eh = clamped ? make_float4(0.f, 0.f, 0.f, 0.f) : make_float4(x.x / n, x.y / n, x.z / n, x.w / n);
            const float dotp = warp_sum(dot4(eh, g));
            float* dst = p.grad[ctx] + (size_t)b * kD + 4 * lane;      // packed behind the loss scalars: not 16-byte aligned
            dst[0] = clamped ? g.x / 1e-12f : (g.x - eh.x * dotp) / n;
            dst[1] = clamped ? g.y / 1e-12f : (g.y - eh.y * dotp) / n;
            dst[2] = clamped ? g.z / 1e-12f : (g.z - eh.z * dotp) / n;
            dst[3] = clamped ? g.w / 1e-12f : (g.w - eh.w * dotp) / n;
        }
    } else if (warp == 2) {
        if (lane < p.num_keys) {
            float l = 0.f;
            for (int s = 0; s < p.splits; ++s) l += __ldcg(p.part_loss + ((size_t)s * p.num_keys + lane) * p.B + b);
            reinterpret_cast<float*>(reinterpret_cast<char*>(p.loss_part) + out_off)[(size_t)lane * (p.group_batch > 0 ? p.group_batch : p.B) + bl] = l;
        }
        if (lane == 0) p.counter[b] = 0u;      // ready for the next launch
    }
    if (!p.do_finalize) return;

    // ---- the last query: batch means + coefficient mix (avid.py:216-233 / avid_cma.py:338-359) ----
    __threadfence();
    __syncthreads();
    stamp(p, 4);
    if (threadIdx.x == 0) s_ticket = atomicAdd(p.counter + p.B, 1u);
    __syncthreads();
    stamp(p, 5);
    if (s_ticket != (unsigned)(p.B - 1)) return;
    __threadfence();
    __shared__ float s_keyloss[AVID_MAX_KEYS];
    for (int k = warp; k < p.num_keys; k += kGatherWarps) {
        float sum = 0.f;
        for (int i = lane; i < p.B; i += 32) sum += __ldcg(p.loss_part + (size_t)k * p.B + i);
        sum = warp_sum(sum) * p.inv_mean_batch;
        if (lane == 0) {
            s_keyloss[k] = sum;
            p.loss_keys[k] = sum;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
        for (int k = 0; k < p.num_keys; ++k) tot += p.weights[k] * s_keyloss[k];
        *p.loss_total = tot;
        p.counter[p.B] = 0u;
    }
    stamp(p, 6);
}

template <int NEED, bool TRAIN>
__global__ void __launch_bounds__(kGatherThreads, 3) nce_gather_kernel(const NceParams p) {
    const int b = blockIdx.x, split = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane >> 3, l8 = lane & 7;
    // gathered queries of a sharded step arrive as one packed record per rank (group): query b = (group g, row bl)
    const int g_i = p.group_batch > 0 ? b / p.group_batch : 0, bl = p.group_batch > 0 ? b - g_i * p.group_batch : b;
    const size_t in_off = (size_t)g_i * p.in_group_stride;          // bytes
    const float* emb0 = reinterpret_cast<const float*>(reinterpret_cast<const char*>(p.emb[0]) + in_off);
    const float* emb1 = reinterpret_cast<const float*>(reinterpret_cast<const char*>(p.emb[1]) + in_off);

    stamp(p, 0);
    // everything the first row requests depend on is asked for first (y), the embeddings are requested before and
    // normalised after the first two quads are on their way, and the Philox rounds of the first chunk run while y is in flight
    GatherCtx gc;
    gc.b = b;
    gc.y = reinterpret_cast<const int64_t*>(reinterpret_cast<const char*>(p.y) + in_off)[bl];
    gc.negs = p.neg_idx ? reinterpret_cast<const int64_t*>(reinterpret_cast<const char*>(p.neg_idx) + in_off) + (size_t)bl * p.K : nullptr;
    float4 e_ctx[2][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        e_ctx[0][j] = reinterpret_cast<const float4*>(emb0 + (size_t)bl * kD)[l8 + 8 * j];
        e_ctx[1][j] = reinterpret_cast<const float4*>(emb1 + (size_t)bl * kD)[l8 + 8 * j];
    }
    const float Z = TRAIN ? *p.Z : 1.0f;

    // item stream of this CTA: [self, positives...] (split 0 only) then negatives [k_begin, k_end); warp w owns a contiguous
    // range of it (a multiple of 4 items, so quads never straddle two warps).  The layout does not depend on y: with an
    // out-of-range y the positive slots stay in the stream and score nothing.
    gc.npos = (split == 0) ? 1 + ((p.positive_set && p.pos_k > 0) ? p.pos_k : 0) : 0;
    gc.k_begin = split * p.kc;
    const int n_items = gc.npos + max(0, min(p.K, gc.k_begin + p.kc) - gc.k_begin);
    const int per_warp = (((n_items + kGatherWarps - 1) / kGatherWarps) + 3) & ~3;
    const int w_begin = warp * per_warp;
    gc.w_end = min(n_items, w_begin + per_warp);
    const int nq = gc.w_end > w_begin ? (gc.w_end - w_begin + 3) >> 2 : 0;
    uint32_t bits[4] = {0u, 0u, 0u, 0u};
    {
        const int item = w_begin + lane;
        if (!gc.negs && item >= gc.npos && item < gc.w_end)
            Philox::generate(p.seed, p.offset + (uint64_t)b * (uint64_t)p.K + (uint64_t)(gc.k_begin + (item - gc.npos)), bits);
    }

    if (gc.y < 0 || gc.y >= p.N) {       // the reference raises IndexError here (avid.py:57-58): report it, score no positive, read nothing
        if (p.bad_index && threadIdx.x == 0 && split == 0) *p.bad_index = 1;
        gc.y = -1;
    }
    gc.pos_row = (p.positive_set && p.pos_k > 0 && gc.y >= 0) ? p.positive_set + (size_t)gc.y * p.pos_k : nullptr;

    float4 acc[2][4];                 // grad_hat for ctx video / audio, this lane's 16 columns
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[c][j] = make_float4(0.f, 0.f, 0.f, 0.f);
    float loss_acc = 0.f;             // lane l8: the terms of key l8 over this group's rows
    const int my_key = l8;
    LaneKey key;
    {
        const bool on = my_key < p.num_keys;
        const KeyDev k = p.keys[on ? my_key : 0];
        key.sel = on ? 2 * k.bank + k.ctx : -1;
        key.neg_limit = on ? k.num_neg : 0;
        key.pos_ok = on ? (k.pos_mode == 0 ? 1 : 2) : 0;
        key.c = (float)k.num_neg * Z;
        key.inv_c = 1.0f / key.c;
        key.inv_p = k.pos_mode == 0 ? 1.0f : 1.0f / (float)p.pos_k;
        key.w_neg = k.weight * p.inv_mean_batch * p.inv_T;
        key.w_pos = -key.w_neg * key.inv_p;
    }

    float4 rv0[4], ra0[4], rv1[4], ra1[4];      // the two quads in flight
#pragma unroll
    for (int j = 0; j < 4; ++j) rv0[j] = ra0[j] = rv1[j] = ra1[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    int tag0 = 0, tag1 = 0;
    ItemDesc desc = describe_item(p, gc, w_begin + lane, bits);
    if (nq > 0) request_quad(p, desc, 0, grp, l8, rv0, ra0, tag0);
    if (nq > 1) request_quad(p, desc, 1, grp, l8, rv1, ra1, tag1);
    normalize_group(e_ctx[0]);
    normalize_group(e_ctx[1]);
    stamp(p, 1);

    for (int t = 0; t < nq; t += 2) {
        score_quad<NEED, TRAIN>(p, rv0, ra0, tag0, e_ctx, acc, loss_acc, key, my_key, b, lane);
        if (t + 2 < nq) {
            if (((t + 2) & 7) == 0) desc = describe_item(p, gc, w_begin + 4 * (t + 2) + lane);      // the next 32 items
            request_quad(p, desc, (t + 2) & 7, grp, l8, rv0, ra0, tag0);
        }
        if (t + 1 < nq) {
            score_quad<NEED, TRAIN>(p, rv1, ra1, tag1, e_ctx, acc, loss_acc, key, my_key, b, lane);
            if (t + 3 < nq) request_quad(p, desc, (t + 3) & 7, grp, l8, rv1, ra1, tag1);
        }
    }
    stamp(p, 2);
    if constexpr (TRAIN) finish_query(p, acc, loss_acc, b, bl, g_i, split, emb0, emb1);
}

struct FinalizeParams {
    const float* emb[2];
    int B, num_keys, splits;
    float inv_mean_batch;
    float weights[AVID_MAX_KEYS];
    const float* part_grad;   // [splits][2][B][128]   (do_reduce)
    const float* part_loss;   // [splits][num_keys][B]
    float* grad_hat[2];       // (B,128) each: written when do_reduce, read when !do_reduce
    float* loss_part;         // (num_keys, B)
    float* grad[2];           // finalize outputs
    float* loss_keys;
    float* loss_total;
    unsigned int* counter;
    int do_reduce, do_finalize;
};

__global__ void __launch_bounds__(256) nce_reduce_finalize_kernel(const FinalizeParams p) {
    const int b = blockIdx.x, t = threadIdx.x;
    const int ctx = t >> 7, e = t & 127;
    __shared__ float s_red[8];
    __shared__ bool s_last;

    float g;
    if (p.do_reduce) {
        g = 0.f;
        for (int s = 0; s < p.splits; ++s) g += p.part_grad[((size_t)(s * 2 + ctx) * p.B + b) * kD + e];
        if (p.grad_hat[ctx]) p.grad_hat[ctx][(size_t)b * kD + e] = g;
        if (t < p.num_keys) {
            float l = 0.f;
            for (int s = 0; s < p.splits; ++s) l += p.part_loss[((size_t)s * p.num_keys + t) * p.B + b];
            p.loss_part[(size_t)t * p.B + b] = l;
        }
    } else {
        g = p.grad_hat[ctx][(size_t)b * kD + e];
    }
    if (!p.do_finalize) return;

    // backward of x -> x / max(||x||, eps):  (g - ehat <ehat, g>) / ||x||   (g / eps when clamped)
    const float x = p.emb[ctx][(size_t)b * kD + e];
    float xx = warp_sum(x * x);
    if ((t & 31) == 0) s_red[t >> 5] = xx;
    __syncthreads();
    const float n2 = s_red[ctx * 4] + s_red[ctx * 4 + 1] + s_red[ctx * 4 + 2] + s_red[ctx * 4 + 3];
    __syncthreads();
    const float n = sqrtf(n2);
    const bool clamped = !(n > 1e-12f);
    const float eh = clamped ? 0.f : x / n;
    float pg = warp_sum(eh * g);
    if ((t & 31) == 0) s_red[t >> 5] = pg;
    __syncthreads();
    const float dotp = s_red[ctx * 4] + s_red[ctx * 4 + 1] + s_red[ctx * 4 + 2] + s_red[ctx * 4 + 3];
    const float out = clamped ? g / 1e-12f : (g - eh * dotp) / n;
    p.grad[ctx][(size_t)b * kD + e] = out;

    // last CTA: batch means + coefficient mix (avid.py:216-233 / avid_cma.py:338-359)
    __threadfence();
    __syncthreads();
    if (t == 0) s_last = atomicAdd(p.counter, 1u) == (unsigned)(gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    __shared__ float s_keyloss[AVID_MAX_KEYS];
    const int w = t >> 5, l = t & 31;
    if (w < p.num_keys) {
        float sum = 0.f;
        for (int i = l; i < p.B; i += 32) sum += __ldcg(p.loss_part + (size_t)w * p.B + i);
        sum = warp_sum(sum) * p.inv_mean_batch;
        if (l == 0) {
            s_keyloss[w] = sum;
            p.loss_keys[w] = sum;
        }
    }
    __syncthreads();
    if (t == 0) {
        float tot = 0.f;
        for (int k = 0; k < p.num_keys; ++k) tot += p.weights[k] * s_keyloss[k];
        *p.loss_total = tot;
        *p.counter = 0u;        // the ticket counters of the workspace are left zero (nce_gather_kernel relies on it)
    }
}

// mean (or, sharded, sum) of exp(score) over the negatives of one key: nce.py:21-36
__global__ void __launch_bounds__(1024) nce_partition_kernel(const float* scores, int B, int K, int stride, int neg_off,
                                                             float scale, float* out) {
    __shared__ double s_red[32];
    double acc = 0.0;
    for (size_t i = threadIdx.x; i < (size_t)B * K; i += blockDim.x) {
        const float s = scores[(i / K) * (size_t)stride + neg_off + (i % K)];
        acc += (double)expf(s);   // slots of rows not held were pre-filled with -inf -> 0
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < (blockDim.x >> 5) ? s_red[threadIdx.x] : 0.0;
        v = warp_sum(v);
        if (threadIdx.x == 0) *out = (float)(v * (double)scale);
    }
}

__global__ void fill_kernel(float* p, size_t n, float v) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

__global__ void sample_negatives_kernel(const int64_t* y, int B, int K, int64_t N, const int32_t* positive_set, int pos_k,
                                        uint64_t seed, uint64_t offset, int64_t* out) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= (size_t)B * K) return;
    const int b = (int)(i / K), k = (int)(i % K);
    const int64_t yb = y[b];
    const int32_t* pos_row = positive_set ? positive_set + (size_t)yb * pos_k : nullptr;
    out[i] = draw_negative(seed, offset, b, k, K, N, yb, pos_row, pos_k);
}

// ------------------------------------------------------------------------------------------------
// Negatives per CTA (a multiple of 16 = 4 warps x one quad) chosen to minimise waves x (items + fixed per-CTA cost) with
// 3 resident CTAs per SM, so the grid neither ends in a thin second wave nor starves the SMs.
static void choose_split(int B, int K, int* splits, int* kc) {
    static const int forced = [] { const char* e = getenv("AVID_NCE_KC"); return e ? atoi(e) : 0; }();     // tuning experiments only
    if (forced > 0) {
        *kc = forced;
        *splits = (K + forced - 1) / forced;
        return;
    }
    const int slots = 3 * kNumSMs;
    long best_cost = -1;
    int best_c = 16;
    const int max_m = (K + 15) / 16;
    for (int m = 1; m <= max_m; ++m) {
        const int c = 16 * m;
        const int sp = (K + c - 1) / c;
        if (sp > 256) continue;
        const long ctas = (long)B * sp;
        const long waves = (ctas + slots - 1) / slots;
        const long cost = waves * (c + 96);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_c = c; }
    }
    *kc = best_c;
    *splits = (K + best_c - 1) / best_c;
    if (*splits < 1) *splits = 1;
}

struct Workspace {
    float *part_grad, *part_loss, *grad_hat[2], *loss_part, *scores;
    unsigned int* counter;
    unsigned long long* dbg;
    size_t bytes;
};

static Workspace carve(void* base, int B, int K, int pos_k, int num_keys, bool with_scores) {
    int splits, kc;
    choose_split(B, K, &splits, &kc);
    Workspace w;
    char* p = static_cast<char*>(base);
    size_t off = 0;
    auto take = [&](size_t nbytes) {
        char* r = p ? p + off : nullptr;
        off += (nbytes + 255) & ~size_t(255);
        return r;
    };
    w.counter = reinterpret_cast<unsigned int*>(take(sizeof(unsigned int) * (size_t)(B + 1)));      // zero on entry, left zero
    w.dbg = reinterpret_cast<unsigned long long*>(take(sizeof(unsigned long long) * 1024 * 8));      // AVID_NCE_DEBUG stamps
    w.part_grad = reinterpret_cast<float*>(take(sizeof(float) * (size_t)splits * 2 * B * kD));
    w.part_loss = reinterpret_cast<float*>(take(sizeof(float) * (size_t)splits * AVID_MAX_KEYS * B));
    w.grad_hat[0] = reinterpret_cast<float*>(take(sizeof(float) * (size_t)B * kD));
    w.grad_hat[1] = reinterpret_cast<float*>(take(sizeof(float) * (size_t)B * kD));
    w.loss_part = reinterpret_cast<float*>(take(sizeof(float) * (size_t)AVID_MAX_KEYS * B));
    w.scores = with_scores ? reinterpret_cast<float*>(take(sizeof(float) * (size_t)B * (1 + pos_k + K))) : nullptr;
    w.bytes = off;
    (void)num_keys;
    return w;
}

// the instantiation for the (bank, ctx) pairs the keys score: cross / self / joint are compile-time, anything else generic
static int launch_gather(const NceParams& p, bool train, cudaStream_t st) {
    const dim3 grid(p.B, p.splits);
    const int mask = (p.need[0][0] ? 1 : 0) | (p.need[0][1] ? 2 : 0) | (p.need[1][0] ? 4 : 0) | (p.need[1][1] ? 8 : 0);
    if (!train) nce_gather_kernel<-1, false><<<grid, kGatherThreads, kGatherSmem, st>>>(p);
    else if (mask == 6) nce_gather_kernel<6, true><<<grid, kGatherThreads, kGatherSmem, st>>>(p);
    else if (mask == 9) nce_gather_kernel<9, true><<<grid, kGatherThreads, kGatherSmem, st>>>(p);
    else if (mask == 15) nce_gather_kernel<15, true><<<grid, kGatherThreads, kGatherSmem, st>>>(p);
    else nce_gather_kernel<-1, true><<<grid, kGatherThreads, kGatherSmem, st>>>(p);
    return check_launch(train ? "nce_gather_kernel" : "nce_gather_kernel(scores)");
}

static int fill_params(const avid_nce_args_t* a, NceParams* p) {
    AVID_REQUIRE(a != nullptr, "nce: args is NULL");
    AVID_REQUIRE(a->emb_video && a->emb_audio && a->y && a->bank_video && a->bank_audio, "nce: NULL tensor pointer");
    AVID_REQUIRE(a->batch > 0 && a->num_neg > 0, "nce: batch (%d) and num_neg (%d) must be positive", a->batch, a->num_neg);
    AVID_REQUIRE(a->num_keys > 0 && a->num_keys <= AVID_MAX_KEYS, "nce: num_keys %d out of range", a->num_keys);
    AVID_REQUIRE(a->row_begin >= 0 && a->row_end <= a->num_rows && a->row_begin < a->row_end, "nce: bad row range [%lld,%lld) of %lld",
                 (long long)a->row_begin, (long long)a->row_end, (long long)a->num_rows);
    AVID_REQUIRE(a->temperature > 0.f, "nce: temperature must be positive");
    bool any_set = false;
    p->need[0][0] = p->need[0][1] = p->need[1][0] = p->need[1][1] = false;
    int pair_keys[4] = {0, 0, 0, 0};
    for (int s = 0; s < 4; ++s) p->coef_lane[s] = -1;
    AVID_REQUIRE(a->row_end - a->row_begin < (int64_t)1 << 31 && a->num_neg < (1 << 28), "nce: more than 2^31 rows per shard or 2^28 negatives");
    for (int k = 0; k < a->num_keys; ++k) {
        const avid_nce_key_t& key = a->keys[k];
        AVID_REQUIRE((key.ctx | 1) == 1 && (key.bank | 1) == 1 && (key.pos_mode | 1) == 1, "nce: key %d has bad ctx/bank/pos_mode", k);
        AVID_REQUIRE(key.num_neg > 0 && key.num_neg <= a->num_neg, "nce: key %d num_neg %d not in (0,%d]", k, key.num_neg, a->num_neg);
        p->keys[k] = KeyDev{key.ctx, key.bank, key.pos_mode, key.num_neg, key.weight};
        p->need[key.bank][key.ctx] = true;
        p->coef_lane[2 * key.bank + key.ctx] = pair_keys[2 * key.bank + key.ctx]++ == 0 ? k : -1;
        any_set |= key.pos_mode == 1;
    }
    AVID_REQUIRE(!any_set || (a->positive_set && a->pos_k > 0 && a->pos_k <= 64), "nce: a positive-set key needs positive_set and 0 < pos_k <= 64");
    AVID_REQUIRE(a->neg_idx || a->num_rows > 1 + (a->positive_set ? a->pos_k : 0), "nce: bank too small to draw negatives from");
    p->bank_used[0] = p->need[0][0] || p->need[0][1];
    p->bank_used[1] = p->need[1][0] || p->need[1][1];
    p->emb[0] = a->emb_video;  p->emb[1] = a->emb_audio;
    p->y = a->y;
    p->bank[0] = a->bank_video;  p->bank[1] = a->bank_audio;
    p->N = a->num_rows;  p->row_begin = a->row_begin;  p->row_end = a->row_end;
    p->B = a->batch;
    p->inv_mean_batch = 1.0f / (float)(a->mean_batch > 0 ? a->mean_batch : a->batch);
    p->K = a->num_neg;
    p->neg_idx = a->neg_idx;  p->seed = a->seed;  p->offset = a->offset;
    // positive_set also steers the negative draw (avid_cma.py:200-207) even if no key scores the set
    p->positive_set = a->positive_set;
    p->pos_k = a->positive_set ? a->pos_k : 0;
    p->num_keys = a->num_keys;
    p->Z = a->avg_exp_score;
    p->inv_T = 1.0f / a->temperature;
    choose_split(a->batch, a->num_neg, &p->splits, &p->kc);
    p->scores = a->scores;
    p->score_pos_k = p->pos_k;
    p->neg_idx_out = a->neg_idx_out;
    p->bad_index = a->bad_index;
    AVID_REQUIRE(a->group_batch >= 0 && (a->group_batch == 0 || a->batch % a->group_batch == 0), "nce: batch %d is not a multiple of group_batch %d",
                 a->batch, a->group_batch);
    AVID_REQUIRE(a->in_group_stride % 16 == 0 && a->out_group_stride % 16 == 0, "nce: group strides must be multiples of 16 bytes");
    p->group_batch = a->group_batch;
    p->in_group_stride = a->group_batch > 0 ? (size_t)a->in_group_stride : 0;
    p->out_group_stride = a->group_batch > 0 ? (size_t)a->out_group_stride : 0;
    return AVID_OK;
}

}  // namespace avid

using namespace avid;

extern "C" {

size_t avid_nce_workspace_bytes(int32_t batch, int32_t num_neg, int32_t pos_k, int32_t num_keys) {
    if (batch <= 0 || num_neg <= 0) return 0;
    return carve(nullptr, batch, num_neg, pos_k < 0 ? 0 : pos_k, num_keys, true).bytes;
}

int avid_nce_forward_backward(const avid_nce_args_t* a, void* workspace, size_t workspace_bytes, void* stream) {
    NceParams p;
    int rc = fill_params(a, &p);
    if (rc) return rc;
    AVID_REQUIRE(a->avg_exp_score, "nce: avg_exp_score is NULL (run avid_nce_partition_mean on the first batch)");
    const bool sharded = a->row_begin != 0 || a->row_end != a->num_rows;
    AVID_REQUIRE(a->loss_keys && a->loss_total && a->grad_video && a->grad_audio || sharded, "nce: NULL output pointer");
    AVID_REQUIRE(!sharded || (a->grad_hat_video && a->grad_hat_audio && a->loss_part), "nce: sharded mode needs grad_hat_* and loss_part");
    Workspace w = carve(workspace, a->batch, a->num_neg, p.pos_k, a->num_keys, false);
    if (!workspace || workspace_bytes < w.bytes) {
        set_error("nce: workspace of %zu bytes given, %zu needed", workspace_bytes, w.bytes);
        return AVID_EWORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    p.part_grad = w.part_grad;  p.part_loss = w.part_loss;  p.counter = w.counter;
    static const bool debug = getenv("AVID_NCE_DEBUG") != nullptr;
    p.dbg = debug ? w.dbg : nullptr;
    p.grad_hat[0] = a->grad_hat_video;  p.grad_hat[1] = a->grad_hat_audio;
    p.loss_part = a->loss_part ? a->loss_part : w.loss_part;
    p.grad[0] = a->grad_video;  p.grad[1] = a->grad_audio;
    p.loss_keys = a->loss_keys;  p.loss_total = a->loss_total;
    for (int k = 0; k < AVID_MAX_KEYS; ++k) p.weights[k] = k < a->num_keys ? a->keys[k].weight : 0.f;
    p.do_finalize = sharded ? 0 : 1;
    AVID_REQUIRE(sharded || a->group_batch == 0, "nce: packed per-rank records (group_batch) are for the sharded protocol only");
    return launch_gather(p, true, st);
}

int avid_nce_finalize(const avid_nce_args_t* a, void* workspace, size_t workspace_bytes, void* stream) {
    AVID_REQUIRE(a && a->emb_video && a->emb_audio && a->grad_hat_video && a->grad_hat_audio && a->loss_part, "nce_finalize: NULL input");
    AVID_REQUIRE(a->loss_keys && a->loss_total && a->grad_video && a->grad_audio, "nce_finalize: NULL output");
    AVID_REQUIRE(a->batch > 0 && a->num_keys > 0 && a->num_keys <= AVID_MAX_KEYS, "nce_finalize: bad batch / num_keys");
    AVID_REQUIRE(workspace && workspace_bytes >= 256, "nce_finalize: workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FinalizeParams f;
    f.emb[0] = a->emb_video;  f.emb[1] = a->emb_audio;
    f.B = a->batch;  f.num_keys = a->num_keys;  f.splits = 0;
    f.inv_mean_batch = 1.0f / (float)(a->mean_batch > 0 ? a->mean_batch : a->batch);
    for (int k = 0; k < AVID_MAX_KEYS; ++k) f.weights[k] = k < a->num_keys ? a->keys[k].weight : 0.f;
    f.part_grad = nullptr;  f.part_loss = nullptr;
    f.grad_hat[0] = a->grad_hat_video;  f.grad_hat[1] = a->grad_hat_audio;
    f.loss_part = a->loss_part;
    f.grad[0] = a->grad_video;  f.grad[1] = a->grad_audio;
    f.loss_keys = a->loss_keys;  f.loss_total = a->loss_total;
    f.counter = static_cast<unsigned int*>(workspace);
    f.do_reduce = 0;  f.do_finalize = 1;
    cudaError_t e = cudaMemsetAsync(workspace, 0, 4, st);
    if (e != cudaSuccess) { set_error("nce_finalize: memset: %s", cudaGetErrorString(e)); return AVID_ECUDA; }
    nce_reduce_finalize_kernel<<<a->batch, 256, 0, st>>>(f);
    return check_launch("nce_reduce_finalize_kernel");
}

int avid_nce_partition_mean(const avid_nce_args_t* a, int32_t key, float* out_mean, void* workspace, size_t workspace_bytes, void* stream) {
    NceParams p;
    int rc = fill_params(a, &p);
    if (rc) return rc;
    AVID_REQUIRE(key >= 0 && key < a->num_keys && out_mean, "nce_partition: bad key / NULL output");
    Workspace w = carve(workspace, a->batch, a->num_neg, p.pos_k, a->num_keys, true);
    if (!workspace || workspace_bytes < w.bytes) {
        set_error("nce_partition: workspace of %zu bytes given, %zu needed", workspace_bytes, w.bytes);
        return AVID_EWORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // scores-only pass for this single key
    NceParams q = p;
    q.num_keys = 1;
    q.keys[0] = p.keys[key];
    q.need[0][0] = q.need[0][1] = q.need[1][0] = q.need[1][1] = false;
    q.need[q.keys[0].bank][q.keys[0].ctx] = true;
    q.bank_used[0] = q.keys[0].bank == 0;
    q.bank_used[1] = q.keys[0].bank == 1;
    q.Z = nullptr;
    q.scores = w.scores;
    q.counter = nullptr;
    q.dbg = nullptr;
    q.part_grad = nullptr;  q.part_loss = nullptr;
    q.neg_idx_out = nullptr;
    const int stride = 1 + q.score_pos_k + a->num_neg;
    const size_t n = (size_t)a->batch * stride;
    fill_kernel<<<(unsigned)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184), 256, 0, st>>>(w.scores, n, -INFINITY);
    if ((rc = check_launch("fill_kernel"))) return rc;
    for (int s = 0; s < 4; ++s) q.coef_lane[s] = -1;
    q.coef_lane[2 * q.keys[0].bank + q.keys[0].ctx] = 0;
    if ((rc = launch_gather(q, false, st))) return rc;
    const bool sharded = a->row_begin != 0 || a->row_end != a->num_rows;
    const int kn = q.keys[0].num_neg;
    const float scale = sharded ? 1.0f : 1.0f / ((float)a->batch * (float)kn);
    nce_partition_kernel<<<1, 1024, 0, st>>>(w.scores, a->batch, kn, stride, 1 + q.score_pos_k, scale, out_mean);
    return check_launch("nce_partition_kernel");
}

int avid_sample_negatives(const int64_t* y, int32_t batch, int32_t num_neg, int64_t num_rows,
                          const int32_t* positive_set, int32_t pos_k, uint64_t seed, uint64_t offset,
                          int64_t* neg_idx_out, void* stream) {
    AVID_REQUIRE(y && neg_idx_out && batch > 0 && num_neg > 0, "sample_negatives: bad arguments");
    AVID_REQUIRE(num_rows > 1 + (positive_set ? pos_k : 0), "sample_negatives: bank too small");
    const size_t n = (size_t)batch * num_neg;
    sample_negatives_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        y, batch, num_neg, num_rows, positive_set, positive_set ? pos_k : 0, seed, offset, neg_idx_out);
    return check_launch("sample_negatives_kernel");
}

}  // extern "C"
