// Data-parallel gradient exchange fused with the optimizer, over NVLink peer memory (include/avid_b200.h, "Sharded optimizer").
//
// The reference averages the 21.3 M fp32 gradients of all ranks with DistributedDataParallel (NCCL all-reduce, utils/main_utils.py:112)
// and then every rank runs the same torch.optim.Adam update on all parameters (main_utils.py:250-256).  Here each rank owns the
// shard [rank * S, (rank + 1) * S) of the flat parameter vector:
//
//   adam_shard_kernel    reduce-scatter + Adam in ONE pass: the owner reads its shard of EVERY rank's flat gradient buffer straight
//                        from peer memory (symmetric allocations, P2P loads over NVLink / NVSwitch), sums them in rank order, applies
//                        the Adam update to its shard of the parameters and of the moments (which only the owner keeps);
//   pull_shards_kernel   all-gather by P2P loads: every rank copies the updated shards of the other ranks into its own flat
//                        parameter buffer, which the parameters of the model are views of.
//
// Per step and GPU (W ranks): 2 x (W - 1) / W x 85 MB cross NVLink (an all-reduce moves the same), no gradient ever makes a second
// trip through HBM, the optimizer reads / writes 1 / W of the moments, and nothing of it needs an SM carve-out next to the persistent
// convolution kernels because it runs after the backward pass in ~0.2 ms.  Ordering between the ranks (gradients complete before
// peers read them, shards updated before peers pull them) is the caller's: two symmetric-memory barriers per step (optim.py).
#include "common.cuh"

namespace avid {

struct PeerPtrs {
    const float* p[AVID_MAX_PEERS];
};

__global__ void __launch_bounds__(256) adam_shard_kernel(float* __restrict__ param, const PeerPtrs grads, int world, float* __restrict__ m,
                                                         float* __restrict__ v, int64_t begin, int64_t count4, float lr_over_bc1, float beta1, float beta2,
                                                         float omb1, float omb2, float eps, float weight_decay, float bc2_sqrt, float grad_scale) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < world; ++r) {       // fixed order: the sum does not depend on who computes it
            const float4 t = ld_stream(reinterpret_cast<const float4*>(grads.p[r] + begin) + i);
            g.x += t.x;  g.y += t.y;  g.z += t.z;  g.w += t.w;
        }
        float4 p4 = reinterpret_cast<float4*>(param + begin)[i];
        float4 m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i];
        float pp[4] = {p4.x, p4.y, p4.z, p4.w}, gg[4] = {g.x, g.y, g.z, g.w}, mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {           // torch.optim.Adam with L2 weight decay, the arithmetic of adam_multi_kernel (misc.cu)
            const float gi = fmaf(weight_decay, pp[j], gg[j] * grad_scale);
            mm[j] = beta1 * mm[j] + omb1 * gi;
            vv[j] = beta2 * vv[j] + omb2 * gi * gi;
            const float denom = sqrtf(vv[j]) / bc2_sqrt + eps;
            pp[j] = pp[j] - lr_over_bc1 * (mm[j] / denom);
        }
        reinterpret_cast<float4*>(param + begin)[i] = make_float4(pp[0], pp[1], pp[2], pp[3]);
        reinterpret_cast<float4*>(m)[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
        reinterpret_cast<float4*>(v)[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
    }
}

// blockIdx.y = which of the other ranks' shards; the shard of rank r is read from rank r's buffer
__global__ void __launch_bounds__(256) pull_shards_kernel(float* __restrict__ param, const PeerPtrs params, int world, int rank, int64_t shard4) {
    const int r = (int)blockIdx.y + ((int)blockIdx.y >= rank ? 1 : 0);
    if (r >= world) return;
    const float4* src = reinterpret_cast<const float4*>(params.p[r]) + (int64_t)r * shard4;
    float4* dst = reinterpret_cast<float4*>(param) + (int64_t)r * shard4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < shard4; i += (int64_t)gridDim.x * blockDim.x) dst[i] = ld_stream(src + i);
}

}  // namespace avid

using namespace avid;

extern "C" {

int avid_adam_shard_step(float* param_flat, const avid_peer_ptrs_t* grads, int32_t world, float* exp_avg, float* exp_avg_sq, int64_t begin,
                         int64_t count, int64_t step, double lr, double beta1, double beta2, double eps, double weight_decay, double grad_scale,
                         void* stream) {
    AVID_REQUIRE(param_flat && grads && exp_avg && exp_avg_sq, "adam_shard_step: NULL pointer");
    AVID_REQUIRE(world >= 1 && world <= AVID_MAX_PEERS, "adam_shard_step: world %d not in [1, %d]", world, AVID_MAX_PEERS);
    AVID_REQUIRE(begin >= 0 && count > 0 && begin % 4 == 0 && count % 4 == 0 && step > 0, "adam_shard_step: shard [%lld, +%lld) must be 4-element aligned, step > 0",
                 (long long)begin, (long long)count);
    PeerPtrs g;
    for (int r = 0; r < AVID_MAX_PEERS; ++r) g.p[r] = r < world ? static_cast<const float*>(grads->ptr[r]) : nullptr;
    for (int r = 0; r < world; ++r) AVID_REQUIRE(g.p[r], "adam_shard_step: gradient buffer of rank %d is NULL", r);
    // hyper-parameters arrive as doubles (python floats) and are rounded to fp32 the way torch.optim.Adam's kernels see them:
    // beta and (1 - beta) are rounded separately -- fl(1 - fl(0.999)) differs from fl(1 - 0.999) by 4.7e-5 relative
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    const int64_t count4 = count / 4;
    int64_t blocks = (count4 + 255) / 256;
    if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
    adam_shard_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(param_flat, g, world, exp_avg, exp_avg_sq, begin, count4,
                                                                                      (float)(lr / bc1), (float)beta1, (float)beta2, (float)(1.0 - beta1),
                                                                                      (float)(1.0 - beta2), (float)eps, (float)weight_decay,
                                                                                      (float)sqrt(bc2), (float)grad_scale);
    return check_launch("adam_shard_kernel");
}

int avid_pull_shards(float* param_flat, const avid_peer_ptrs_t* params, int32_t world, int32_t rank, int64_t shard, void* stream) {
    AVID_REQUIRE(param_flat && params, "pull_shards: NULL pointer");
    AVID_REQUIRE(world >= 1 && world <= AVID_MAX_PEERS && rank >= 0 && rank < world, "pull_shards: bad world / rank");
    AVID_REQUIRE(shard > 0 && shard % 4 == 0, "pull_shards: the shard length must be a positive multiple of 4");
    if (world == 1) return AVID_OK;
    PeerPtrs p;
    for (int r = 0; r < AVID_MAX_PEERS; ++r) p.p[r] = r < world ? static_cast<const float*>(params->ptr[r]) : nullptr;
    for (int r = 0; r < world; ++r) AVID_REQUIRE(p.p[r], "pull_shards: parameter buffer of rank %d is NULL", r);
    const int64_t shard4 = shard / 4;
    int64_t bx = (shard4 + 255) / 256;
    const int64_t cap = (8 * kNumSMs + world - 2) / (world - 1);
    if (bx > cap) bx = cap;
    pull_shards_kernel<<<dim3((unsigned)bx, (unsigned)(world - 1)), 256, 0, static_cast<cudaStream_t>(stream)>>>(param_flat, p, world, rank, shard4);
    return check_launch("pull_shards_kernel");
}

}  // extern "C"
