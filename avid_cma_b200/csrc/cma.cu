// CMA positive mining (include/avid_b200.h, avid_cma_topk_*): brute-force cosine top-(k+1) of every
// query row against every candidate row in both modalities, agreement = min / max / one modality,
// streaming top-k -- the N x N similarity matrix is never written.  Replaces
// CMASampler.sample_instance (criterions/avid_cma.py:42-73), which re-streams both whole banks from
// HBM for every 16 queries; here a CTA keeps a 64-query tile resident in shared memory and streams
// the candidates once per 64 queries, fp32 FMA throughout so the ranking matches the fp32 reference.
#include <math.h>
#include "common.cuh"

namespace avid {

constexpr int kTQ = 64, kTC = 64, kLD = kD + 4, kSlots = 64;

struct CmaSmem {
    float q[2][kTQ][kLD];
    float c[2][kTC][kLD];
    float sim[kTQ][kTC + 1];
    float val[kTQ][kSlots + 1];
    int idx[kTQ][kSlots + 1];
};

__device__ __forceinline__ void load_tile(float (*dst)[kLD], const float* src, int64_t rows_left, int tid) {
    // 64 rows x 128 floats, one float4 per (row, lane) -> coalesced 512-byte rows
    for (int i = tid; i < kTQ * 32; i += 256) {
        const int r = i >> 5, l = i & 31;
        float4 v = make_float4(0, 0, 0, 0);
        if (r < rows_left) v = reinterpret_cast<const float4*>(src + (size_t)r * kD)[l];
        *reinterpret_cast<float4*>(&dst[r][l * 4]) = v;
    }
}

__global__ void __launch_bounds__(256, 1) cma_scan_kernel(const float* q_video, const float* q_audio, int64_t num_queries,
                                                          const float* c_video, const float* c_audio, int64_t cand_begin,
                                                          int64_t num_cand, int mode, int slots, float* top_val, int* top_idx) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CmaSmem& sm = *reinterpret_cast<CmaSmem*>(smem_raw);
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t q0 = (int64_t)blockIdx.x * kTQ;
    const int64_t q_left = num_queries - q0;
    const bool use_v = mode != 3, use_a = mode != 2;

    if (use_v) load_tile(sm.q[0], q_video + (size_t)q0 * kD, q_left, tid);
    if (use_a) load_tile(sm.q[1], q_audio + (size_t)q0 * kD, q_left, tid);
    // running top lists of this query tile (thread q < 64 owns list q)
    float thr = 0.f;
    int min_pos = 0;
    if (tid < kTQ) {
        thr = INFINITY;
        for (int s = 0; s < slots; ++s) {
            float v = -INFINITY;
            int ix = -1;
            if (tid < q_left) {
                v = top_val[(size_t)(q0 + tid) * kSlots + s];
                ix = top_idx[(size_t)(q0 + tid) * kSlots + s];
            }
            sm.val[tid][s] = v;
            sm.idx[tid][s] = ix;
            if (v < thr) { thr = v; min_pos = s; }
        }
    }

    for (int64_t c0 = 0; c0 < num_cand; c0 += kTC) {
        __syncthreads();   // previous tile fully consumed (sim scan + c tiles)
        if (use_v) load_tile(sm.c[0], c_video + (size_t)c0 * kD, num_cand - c0, tid);
        if (use_a) load_tile(sm.c[1], c_audio + (size_t)c0 * kD, num_cand - c0, tid);
        __syncthreads();

        float sv[4][4], sa[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) sv[i][j] = sa[i][j] = 0.f;
        // thread (ty, tx) owns queries ty + 16 i and candidates tx + 16 j: conflict-free LDS.128
#pragma unroll 4
        for (int k = 0; k < kD; k += 4) {
            float4 a[4], b[4];
            if (use_v) {
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(&sm.q[0][ty + 16 * i][k]);
#pragma unroll
                for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4*>(&sm.c[0][tx + 16 * j][k]);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        sv[i][j] = fmaf(a[i].w, b[j].w, fmaf(a[i].z, b[j].z, fmaf(a[i].y, b[j].y, fmaf(a[i].x, b[j].x, sv[i][j]))));
            }
            if (use_a) {
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(&sm.q[1][ty + 16 * i][k]);
#pragma unroll
                for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4*>(&sm.c[1][tx + 16 * j][k]);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        sa[i][j] = fmaf(a[i].w, b[j].w, fmaf(a[i].z, b[j].z, fmaf(a[i].y, b[j].y, fmaf(a[i].x, b[j].x, sa[i][j]))));
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float s;
                if (mode == 0) s = fminf(sv[i][j], sa[i][j]);        // consensus (avid_cma.py:56-57)
                else if (mode == 1) s = fmaxf(sv[i][j], sa[i][j]);   // union     (avid_cma.py:58-59)
                else if (mode == 2) s = sv[i][j];
                else s = sa[i][j];
                sm.sim[ty + 16 * i][tx + 16 * j] = s;
            }
        __syncthreads();
        if (tid < kTQ && tid < q_left) {
            const int n = (int)min((int64_t)kTC, num_cand - c0);
            for (int c = 0; c < n; ++c) {
                const float s = sm.sim[tid][c];
                if (s > thr) {
                    sm.val[tid][min_pos] = s;
                    sm.idx[tid][min_pos] = (int)(cand_begin + c0 + c);
                    thr = INFINITY;
                    for (int t = 0; t < slots; ++t) {
                        const float v = sm.val[tid][t];
                        if (v < thr) { thr = v; min_pos = t; }
                    }
                }
            }
        }
    }
    __syncthreads();
    if (tid < kTQ && tid < q_left)
        for (int s = 0; s < slots; ++s) {
            top_val[(size_t)(q0 + tid) * kSlots + s] = sm.val[tid][s];
            top_idx[(size_t)(q0 + tid) * kSlots + s] = sm.idx[tid][s];
        }
}

__global__ void cma_begin_kernel(float* top_val, int* top_idx, float* top_exact, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        top_val[i] = -INFINITY;
        top_idx[i] = -1;
        top_exact[i] = -INFINITY;
    }
}

// sorted=True top-(k+1) (descending similarity, ties by ascending index), first entry dropped
// (avid_cma.py:68-69), remaining indices sorted ascending (avid_cma.py:70)
__global__ void cma_finish_kernel(const float* top_val, const int* top_idx, int64_t num_queries, int pos_k, int32_t* out) {
    const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (q >= num_queries) return;
    float v[kSlots];
    int ix[kSlots];
    const int n = pos_k + 1;
    for (int s = 0; s < n; ++s) {
        v[s] = top_val[(size_t)q * kSlots + s];
        ix[s] = top_idx[(size_t)q * kSlots + s];
    }
    // find and drop the best entry
    int best = 0;
    for (int s = 1; s < n; ++s)
        if (v[s] > v[best] || (v[s] == v[best] && ix[s] < ix[best])) best = s;
    ix[best] = ix[n - 1];
    // insertion sort of the remaining pos_k indices
    for (int s = 1; s < pos_k; ++s) {
        const int key = ix[s];
        int t = s - 1;
        while (t >= 0 && ix[t] > key) { ix[t + 1] = ix[t]; --t; }
        ix[t + 1] = key;
    }
    for (int s = 0; s < pos_k; ++s) out[(size_t)q * pos_k + s] = ix[s];
}

}  // namespace avid

using namespace avid;

extern "C" {

size_t avid_cma_topk_workspace_bytes(int64_t num_queries) {
    // running top lists: similarity, candidate row and -- for the tensor-core path (cma_tc.cu) -- the exact fp32 similarity
    return num_queries > 0 ? (size_t)num_queries * kSlots * (2 * sizeof(float) + sizeof(int)) : 0;
}

static int cma_split(int64_t num_queries, void* workspace, size_t bytes, float** val, int** idx) {
    AVID_REQUIRE(num_queries > 0 && workspace, "cma_topk: bad arguments");
    if (bytes < avid_cma_topk_workspace_bytes(num_queries)) {
        set_error("cma_topk: workspace of %zu bytes given, %zu needed", bytes, avid_cma_topk_workspace_bytes(num_queries));
        return AVID_EWORKSPACE;
    }
    *val = static_cast<float*>(workspace);
    *idx = reinterpret_cast<int*>(*val + (size_t)num_queries * kSlots);
    return AVID_OK;
}

int avid_cma_topk_begin(int64_t num_queries, void* workspace, size_t workspace_bytes, void* stream) {
    float* val; int* idx;
    int rc = cma_split(num_queries, workspace, workspace_bytes, &val, &idx);
    if (rc) return rc;
    cma_begin_kernel<<<4 * kNumSMs, 256, 0, static_cast<cudaStream_t>(stream)>>>(val, idx, reinterpret_cast<float*>(idx + (size_t)num_queries * kSlots),
                                                                                 (size_t)num_queries * kSlots);
    return check_launch("cma_begin_kernel");
}

int avid_cma_topk_scan(const float* q_video, const float* q_audio, int64_t num_queries,
                       const float* cand_video, const float* cand_audio, int64_t cand_begin, int64_t num_cand,
                       int32_t mode, int32_t pos_k, void* workspace, size_t workspace_bytes, void* stream) {
    float* val; int* idx;
    int rc = cma_split(num_queries, workspace, workspace_bytes, &val, &idx);
    if (rc) return rc;
    AVID_REQUIRE(q_video && q_audio && cand_video && cand_audio, "cma_topk_scan: NULL pointer");
    AVID_REQUIRE(mode >= 0 && mode <= 3, "cma_topk_scan: unknown mode %d", mode);
    AVID_REQUIRE(pos_k > 0 && pos_k < kSlots, "cma_topk_scan: pos_k %d not in (0,%d)", pos_k, kSlots);
    AVID_REQUIRE(num_cand > 0 && cand_begin >= 0 && cand_begin + num_cand < (int64_t)1 << 31, "cma_topk_scan: bad candidate range");
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(cma_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CmaSmem));
        if (e != cudaSuccess) { set_error("cma_topk_scan: smem attribute: %s", cudaGetErrorString(e)); return AVID_ECUDA; }
        configured = true;
    }
    const unsigned grid = (unsigned)((num_queries + kTQ - 1) / kTQ);
    cma_scan_kernel<<<grid, 256, sizeof(CmaSmem), static_cast<cudaStream_t>(stream)>>>(
        q_video, q_audio, num_queries, cand_video, cand_audio, cand_begin, num_cand, mode, pos_k + 1, val, idx);
    return check_launch("cma_scan_kernel");
}

int avid_cma_topk_finish(int64_t num_queries, int32_t pos_k, int32_t* positive_set_out, void* workspace, size_t workspace_bytes, void* stream) {
    float* val; int* idx;
    int rc = cma_split(num_queries, workspace, workspace_bytes, &val, &idx);
    if (rc) return rc;
    AVID_REQUIRE(positive_set_out && pos_k > 0 && pos_k < kSlots, "cma_topk_finish: bad arguments");
    cma_finish_kernel<<<(unsigned)((num_queries + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(val, idx, num_queries, pos_k, positive_set_out);
    return check_launch("cma_finish_kernel");
}

}  // extern "C"
