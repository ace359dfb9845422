// Fused memory-bank gather + score + NCE loss + closed-form gradient (include/avid_b200.h,
// "Criterion").  Replaces the ATen chain of criterions/avid.py:47-80, avid_cma.py:150-209 and
// criterions/nce.py:38-58 of the reference: the (B,K,128) gathers are never materialised, every
// bank row is read from HBM exactly once (one coalesced 512-byte request per warp) and serves the
// forward score, the loss term and the gradient w.r.t. the embedding in the same pass.
//
// Kernel 1  nce_gather_kernel   grid (B, splits) x 128 threads.  A warp streams rows 8 at a time:
//           groups of 8 lanes own one row each (16 columns per lane, 128-byte coalesced group loads), 2 x 4 rows x
//           2 banks in flight per warp; dot product = 16 FMAs + 3 shuffles, the NCE math runs per group, the
//           gradient axpy needs no broadcast.  Partial (grad_hat, loss) per split go to the workspace.
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
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) s += dot4(a[j], b[j]);
    return s;
}

// x / max(||x||, 1e-12) for the row `b` of a (B,128) matrix in the group layout (avid.py:52-53)
__device__ __forceinline__ void load_normalized(const float* emb, int b, int l8, float4 (&out)[4]) {
    const float4* src = reinterpret_cast<const float4*>(emb + (size_t)b * kD);
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        out[j] = src[l8 + 8 * j];
        ss += dot4(out[j], out[j]);
    }
    const float inv = 1.0f / fmaxf(sqrtf(group_sum(ss)), 1e-12f);
#pragma unroll
    for (int j = 0; j < 4; ++j) out[j] = make_float4(out[j].x * inv, out[j].y * inv, out[j].z * inv, out[j].w * inv);
}

// a warp keeps 2 row quads (4 rows each) per bank in flight: 2 x 4 rows x 2 banks x 512 B = 8 KB.
// Measured and dropped (round 2): landing the rows in a per-warp shared-memory ring instead of registers -- with per-row
// cp.async.bulk copies (request-rate bound: ~35 cycles per 512-byte request and SM; K = 1024: 49 us vs 39 us) and with cp.async
// 16 B per lane (58 us; K = 16384: 0.60 instead of 0.77 of the HBM peak): the extra shared-memory hop costs more than the deeper
// queue gives, because the kernel is bound by its few dependent round trips, not by bytes in flight.  Also dropped: a per-lane
// prefetch.global.L2 of the chunk's rows before scoring it (32 scattered lines per instruction: K = 1024 47 us, K = 16384 0.52),
// and TMA tile::gather4 requests (4 rows per request, scripts/probes/gather4_tma.cu: correct, but the TMA unit serves ~one 512-byte
// row per ~60 cycles and SM: K = 1024 49 us, K = 16384 0.35 of the HBM peak with a 6-stage ring).  Same box, register path:
// 37 us / 0.54 (K = 4096) / 0.77 (K = 16384).  Little's law on those numbers: ~96 KB in flight per SM at 5.0 TB/s is an
// effective round trip of ~2.8 us for random 512-byte rows, so K = 1024 (8 dependent round trips per warp) cannot go below ~25 us
// in this structure.
constexpr int kGatherSmem = 0;

__global__ void __launch_bounds__(kGatherThreads, 3) nce_gather_kernel(const NceParams p) {
    const int b = blockIdx.x, split = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane >> 3, l8 = lane & 7;
    // gathered queries of a sharded step arrive as one packed record per rank (group): query b = (group g, row bl)
    const int g_i = p.group_batch > 0 ? b / p.group_batch : 0, bl = p.group_batch > 0 ? b - g_i * p.group_batch : b;
    const size_t in_off = (size_t)g_i * p.in_group_stride;          // bytes
    const float* emb0 = reinterpret_cast<const float*>(reinterpret_cast<const char*>(p.emb[0]) + in_off);
    const float* emb1 = reinterpret_cast<const float*>(reinterpret_cast<const char*>(p.emb[1]) + in_off);
    const int64_t* negs = p.neg_idx ? reinterpret_cast<const int64_t*>(reinterpret_cast<const char*>(p.neg_idx) + in_off) + (size_t)bl * p.K : nullptr;

    float4 e_ctx[2][4];
    load_normalized(emb0, bl, l8, e_ctx[0]);
    load_normalized(emb1, bl, l8, e_ctx[1]);
    int64_t y = reinterpret_cast<const int64_t*>(reinterpret_cast<const char*>(p.y) + in_off)[bl];
    if (y < 0 || y >= p.N) {       // the reference raises IndexError here (avid.py:57-58): report it, score no positive, read nothing
        if (p.bad_index && threadIdx.x == 0 && split == 0) *p.bad_index = 1;
        y = -1;
    }
    const float Z = p.Z ? *p.Z : 1.0f;
    const int32_t* pos_row = (p.positive_set && p.pos_k > 0 && y >= 0) ? p.positive_set + (size_t)y * p.pos_k : nullptr;

    // item stream of this CTA: [self, positives...] (split 0 only) then negatives [k_begin, k_end)
    const int npos = (split == 0) ? 1 + (pos_row ? p.pos_k : 0) : 0;
    const int k_begin = split * p.kc, k_end = min(p.K, k_begin + p.kc);
    const int n_items = npos + max(0, k_end - k_begin);

    float4 acc[2][4];                 // grad_hat for ctx video / audio, this lane's 16 columns
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[c][j] = make_float4(0.f, 0.f, 0.f, 0.f);
    // the transcendental NCE math is lane-parallel inside a group: lane l8 evaluates (row quad l8 >> 2, key (l8 & 3) + 4 * pass)
    float loss_acc[2] = {0.f, 0.f};
    const int key_passes = (p.num_keys + 3) >> 2;
    const int u_sel = l8 >> 2;

    for (int c0 = warp * 32; c0 < n_items; c0 += kGatherWarps * 32) {
        // each lane describes one item of the chunk: kind 0 = negative k, 1 = self, 2 = positive-set entry
        const int item = c0 + lane;
        int kind = -1, kk = 0;
        int64_t idx = -1;
        if (item < n_items) {
            if (item < npos) {
                kind = item == 0 ? 1 : 2;
                kk = item - 1;
                idx = item == 0 ? y : (int64_t)pos_row[item - 1];
            } else {
                kind = 0;
                kk = k_begin + (item - npos);
                idx = negs ? negs[kk] : draw_negative(p.seed, p.offset, b, kk, p.K, p.N, y, pos_row, p.pos_k);
                if (p.neg_idx_out) p.neg_idx_out[(size_t)b * p.K + kk] = idx;
            }
        }
        if (!(idx >= p.row_begin && idx < p.row_end)) kind = -1;      // rows another shard holds are scored there
        const int n_chunk = min(32, n_items - c0);

        for (int j0 = 0; j0 < n_chunk; j0 += 8) {
            float4 rv[2][4], ra[2][4];
            int kind_u[2], kk_u[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int src = (j0 + 4 * u + grp) & 31;          // the item this group scores in quad u
                const int64_t idx_u = __shfl_sync(0xffffffffu, idx, src);
                kind_u[u] = __shfl_sync(0xffffffffu, kind, src);
                kk_u[u] = __shfl_sync(0xffffffffu, kk, src);
                if (j0 + 4 * u + grp >= n_chunk) kind_u[u] = -1;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    rv[u][j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    ra[u][j] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                if (kind_u[u] >= 0) {
                    const size_t off = (size_t)(idx_u - p.row_begin) * kD;
                    if (p.bank_used[0]) {
                        const float4* r = reinterpret_cast<const float4*>(p.bank[0] + off) + l8;
#pragma unroll
                        for (int j = 0; j < 4; ++j) rv[u][j] = ld_stream(r + 8 * j);
                    }
                    if (p.bank_used[1]) {
                        const float4* r = reinterpret_cast<const float4*>(p.bank[1] + off) + l8;
#pragma unroll
                        for (int j = 0; j < 4; ++j) ra[u][j] = ld_stream(r + 8 * j);
                    }
                }
            }
            // d[u][bank][ctx] of this group's two rows, on all 8 lanes of the group
            float d[2][2][2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                d[u][0][0] = p.need[0][0] ? group_sum(dot16(rv[u], e_ctx[0])) : 0.f;
                d[u][0][1] = p.need[0][1] ? group_sum(dot16(rv[u], e_ctx[1])) : 0.f;
                d[u][1][0] = p.need[1][0] ? group_sum(dot16(ra[u], e_ctx[0])) : 0.f;
                d[u][1][1] = p.need[1][1] ? group_sum(dot16(ra[u], e_ctx[1])) : 0.f;
            }
            const int kind_m = u_sel ? kind_u[1] : kind_u[0];
            const int kk_m = u_sel ? kk_u[1] : kk_u[0];
            float cf[2][2][2] = {{{0.f, 0.f}, {0.f, 0.f}}, {{0.f, 0.f}, {0.f, 0.f}}};      // [u][bank][ctx]: dL/ds / T summed over keys
#pragma unroll
            for (int kp = 0; kp < 2; ++kp) {
                if (kp >= key_passes) break;
                const int my_key = (l8 & 3) + 4 * kp;
                float coef = 0.f;
                if (my_key < p.num_keys && kind_m >= 0) {
                    const KeyDev key = p.keys[my_key];
                    const bool applies = (kind_m == 0 && kk_m < key.num_neg) || (kind_m == 1 && key.pos_mode == 0) ||
                                         (kind_m == 2 && key.pos_mode == 1);
                    if (applies) {
                        const float d0 = key.bank == 0 ? (key.ctx == 0 ? d[0][0][0] : d[0][0][1]) : (key.ctx == 0 ? d[0][1][0] : d[0][1][1]);
                        const float d1 = key.bank == 0 ? (key.ctx == 0 ? d[1][0][0] : d[1][0][1]) : (key.ctx == 0 ? d[1][1][0] : d[1][1][1]);
                        const float s = (u_sel ? d1 : d0) * p.inv_T;
                        if (p.scores) {
                            const int slot = kind_m == 1 ? 0 : (kind_m == 2 ? 1 + kk_m : 1 + p.score_pos_k + kk_m);
                            p.scores[((size_t)my_key * p.B + b) * (size_t)(1 + p.score_pos_k + p.K) + slot] = s;
                        }
                        if (p.Z) {
                            // nce.py:42-57 with c = K_key * Z
                            const float c = (float)key.num_neg * Z;
                            const float e = expf(s);
                            if (kind_m == 0) {
                                loss_acc[kp] += log1pf(e / c);
                                coef = key.weight * p.inv_mean_batch * (e / (e + c));
                            } else {
                                const float inv_p = key.pos_mode == 0 ? 1.0f : 1.0f / (float)p.pos_k;
                                loss_acc[kp] += inv_p * log1pf(c / e);
                                coef = -key.weight * p.inv_mean_batch * inv_p * (c / (e + c));
                            }
                            coef *= p.inv_T;
                        }
                    }
                }
                if (!p.Z) continue;
                // hand every (quad, key) coefficient to the 8 lanes of the group
                const int nk = min(4, p.num_keys - 4 * kp);
                for (int q = 0; q < nk; ++q) {
                    const int bank = p.keys[q + 4 * kp].bank, ctx = p.keys[q + 4 * kp].ctx;      // uniform
                    const float c0 = __shfl_sync(0xffffffffu, coef, (lane & 24) | q);
                    const float c1 = __shfl_sync(0xffffffffu, coef, (lane & 24) | 4 | q);
                    if (bank == 0) {
                        if (ctx == 0) { cf[0][0][0] += c0; cf[1][0][0] += c1; } else { cf[0][0][1] += c0; cf[1][0][1] += c1; }
                    } else {
                        if (ctx == 0) { cf[0][1][0] += c0; cf[1][1][0] += c1; } else { cf[0][1][1] += c0; cf[1][1][1] += c1; }
                    }
                }
            }
            if (!p.Z) continue;
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    if (p.need[0][c]) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) axpy4(acc[c][j], cf[u][0][c], rv[u][j]);
                    }
                    if (p.need[1][c]) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) axpy4(acc[c][j], cf[u][1][c], ra[u][j]);
                    }
                }
        }
    }
    if (!p.Z) return;

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
#pragma unroll
    for (int kp = 0; kp < 2; ++kp) {
        float v = loss_acc[kp];         // lane l8 holds the terms of key (l8 & 3) + 4 * kp: sum over quads (xor 4) and groups (xor 8, 16)
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        if (lane < 4) s_loss[warp][lane + 4 * kp] = v;
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
    __shared__ float s_red[8];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_ticket = atomicAdd(p.counter + b, 1u);
    __syncthreads();
    if (s_ticket != (unsigned)(p.splits - 1)) return;
    __threadfence();
    const int e = threadIdx.x;          // 0..127: one embedding column, both contexts
    const size_t out_off = (size_t)g_i * p.out_group_stride;
    float g[2];
#pragma unroll
    for (int ctx = 0; ctx < 2; ++ctx) {
        float a = 0.f;
        for (int s = 0; s < p.splits; ++s) a += __ldcg(p.part_grad + ((size_t)(s * 2 + ctx) * p.B + b) * kD + e);
        g[ctx] = a;
        if (p.grad_hat[ctx]) reinterpret_cast<float*>(reinterpret_cast<char*>(p.grad_hat[ctx]) + out_off)[(size_t)bl * kD + e] = a;
    }
    if (e < p.num_keys) {
        float l = 0.f;
        for (int s = 0; s < p.splits; ++s) l += __ldcg(p.part_loss + ((size_t)s * p.num_keys + e) * p.B + b);
        reinterpret_cast<float*>(reinterpret_cast<char*>(p.loss_part) + out_off)[(size_t)e * (p.group_batch > 0 ? p.group_batch : p.B) + bl] = l;
    }
    if (e == 0) p.counter[b] = 0u;      // ready for the next launch
    if (!p.do_finalize) return;

    // backward of x -> x / max(||x||, eps):  (g - ehat <ehat, g>) / ||x||   (g / eps when clamped)
#pragma unroll
    for (int ctx = 0; ctx < 2; ++ctx) {
        const float x = (ctx ? emb1 : emb0)[(size_t)bl * kD + e];
        const float xx = warp_sum(x * x);
        __syncthreads();
        if (lane == 0) s_red[warp] = xx;
        __syncthreads();
        const float n = sqrtf(s_red[0] + s_red[1] + s_red[2] + s_red[3]);
        const bool clamped = !(n > 1e-12f);
        const float eh = clamped ? 0.f : x / n;
        const float pg = warp_sum(eh * g[ctx]);
        if (lane == 0) s_red[4 + warp] = pg;
        __syncthreads();
        const float dotp = s_red[4] + s_red[5] + s_red[6] + s_red[7];
        p.grad[ctx][(size_t)b * kD + e] = clamped ? g[ctx] / 1e-12f : (g[ctx] - eh * dotp) / n;
    }

    // ---- the last query: batch means + coefficient mix (avid.py:216-233 / avid_cma.py:338-359) ----
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_ticket = atomicAdd(p.counter + p.B, 1u);
    __syncthreads();
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
// Items per CTA (a multiple of 128 = 4 warps x 32-item chunks) chosen to minimise waves x (items + fixed per-CTA cost) with
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
    int best_c = 128;
    const int max_m = (K + 127) / 128;
    for (int m = 1; m <= max_m; ++m) {
        const int c = 128 * m;
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

static int configure_gather() {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(nce_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGatherSmem);
        if (e != cudaSuccess) { set_error("nce: shared-memory attribute: %s", cudaGetErrorString(e)); return AVID_ECUDA; }
        configured = true;
    }
    return AVID_OK;
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
    for (int k = 0; k < a->num_keys; ++k) {
        const avid_nce_key_t& key = a->keys[k];
        AVID_REQUIRE((key.ctx | 1) == 1 && (key.bank | 1) == 1 && (key.pos_mode | 1) == 1, "nce: key %d has bad ctx/bank/pos_mode", k);
        AVID_REQUIRE(key.num_neg > 0 && key.num_neg <= a->num_neg, "nce: key %d num_neg %d not in (0,%d]", k, key.num_neg, a->num_neg);
        p->keys[k] = KeyDev{key.ctx, key.bank, key.pos_mode, key.num_neg, key.weight};
        p->need[key.bank][key.ctx] = true;
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
    p.grad_hat[0] = a->grad_hat_video;  p.grad_hat[1] = a->grad_hat_audio;
    p.loss_part = a->loss_part ? a->loss_part : w.loss_part;
    p.grad[0] = a->grad_video;  p.grad[1] = a->grad_audio;
    p.loss_keys = a->loss_keys;  p.loss_total = a->loss_total;
    for (int k = 0; k < AVID_MAX_KEYS; ++k) p.weights[k] = k < a->num_keys ? a->keys[k].weight : 0.f;
    p.do_finalize = sharded ? 0 : 1;
    AVID_REQUIRE(sharded || a->group_batch == 0, "nce: packed per-rank records (group_batch) are for the sharded protocol only");
    if ((rc = configure_gather())) return rc;
    nce_gather_kernel<<<dim3(a->batch, p.splits), kGatherThreads, kGatherSmem, st>>>(p);
    return check_launch("nce_gather_kernel");
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
    q.part_grad = nullptr;  q.part_loss = nullptr;
    q.neg_idx_out = nullptr;
    const int stride = 1 + q.score_pos_k + a->num_neg;
    const size_t n = (size_t)a->batch * stride;
    fill_kernel<<<(unsigned)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184), 256, 0, st>>>(w.scores, n, -INFINITY);
    if ((rc = check_launch("fill_kernel"))) return rc;
    if ((rc = configure_gather())) return rc;
    nce_gather_kernel<<<dim3(a->batch, q.splits), kGatherThreads, kGatherSmem, st>>>(q);
    if ((rc = check_launch("nce_gather_kernel(scores)"))) return rc;
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
