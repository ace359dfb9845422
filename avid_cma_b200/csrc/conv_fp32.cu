// fp32 implicit-GEMM convolutions for the encoders (include/avid_b200.h, AVID_MATH_FP32): forward,
// input gradient and filter gradient of every nn.Conv3d / nn.Conv2d of the reference towers
// (models/video.py:20, models/audio.py:22, models/network_blocks.py:18-20,35-49) on channels-last
// activations [n, t, h, w, c] and tap-major filters [tap, c_src, c_dst].  No im2col buffer exists:
// the A tile of the GEMM is gathered straight from the activation tensor (one float4 of 4 channels
// per thread per tap), zero padding and stride handled in the address computation.
//
//   forward : out[m, co]  = sum_{tap, ci} in[src(m, tap), ci]  * w [tap, ci, co]   (+ addend)
//   dgrad   : din[m, ci]  = sum_{tap, co} dout[src'(m, tap), co] * wT[tap, co, ci] (+ addend)
//   wgrad   : dw[tap, ci, co] = sum_m in[src(m, tap), ci] * dout[m, co]            (split over m)
//
// CUDA-core FMA path: exact fp32 products, used as the parity mode and as the fallback of the
// tcgen05 path (conv_tc.cu).
#include "common.cuh"

namespace avid {

struct Geom {
    // destination pixels (GEMM rows) and the tensor the A operand is gathered from
    int n, td, hd, wd, cd;      // destination [n, td, hd, wd, cd]
    int ts, hs, ws, cs;         // source      [n, ts, hs, ws, cs]
    int kt, kh, kw, st, sh, sw, pt, ph, pw;
    int taps;
    int M;                      // n * td * hd * wd
};

constexpr int BM = 128, BK = 16, LDA = BM + 4;

// source coordinate of destination coordinate d for tap k; false when it falls in the padding
template <bool DGRAD>
__device__ __forceinline__ bool src_coord(int d, int k, int stride, int pad, int size, int& s) {
    if (!DGRAD) {
        s = d * stride - pad + k;
        return s >= 0 && s < size;
    } else {
        const int num = d + pad - k;
        if (num < 0) return false;
        s = num / stride;
        return s * stride == num && s < size;
    }
}

template <int BN, bool DGRAD>
__global__ void __launch_bounds__(256) conv_igemm_kernel(const Geom g, const float* __restrict__ src,
                                                         const float* __restrict__ filt, const float* __restrict__ addend,
                                                         float* __restrict__ dst) {
    constexpr int TN = BN / 16;              // columns per thread (4 or 8)
    constexpr int BLD = BN;                  // B smem row
    constexpr int B4 = BK * BN / 4 / 256;    // float4 of B per thread per chunk (1 or 2)
    __shared__ __align__(16) float As[2][BK][LDA];
    __shared__ __align__(16) float Bs[2][BK][BLD];

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

    // K decomposition: a chunk of 16 = tpc taps x cpc source channels
    const int cpc = g.cs < BK ? g.cs : BK;
    const int tpc = BK / cpc;
    const int cchunks = g.cs / cpc;
    const int nchunks = ((g.taps + tpc - 1) / tpc) * cchunks;

    // A loader: rows (tid>>2) and (tid>>2)+64, float4 index q = tid&3 inside the 16-wide chunk
    const int q = tid & 3;
    const int a_tap_sub = (q * 4) / cpc, a_c_sub = (q * 4) % cpc;
    int a_n[2], a_t[2], a_h[2], a_w[2];
    bool a_ok[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        int m = m0 + (tid >> 2) + 64 * i;
        a_ok[i] = m < g.M;
        if (!a_ok[i]) m = 0;
        a_w[i] = m % g.wd;  m /= g.wd;
        a_h[i] = m % g.hd;  m /= g.hd;
        a_t[i] = m % g.td;  a_n[i] = m / g.td;
    }
    // B loader
    const int b_row = (tid * 4) / BN;          // + (256*4/BN) * j
    const int b_col = (tid * 4) % BN;

    float4 ra[2], rb[B4];
    auto load_chunk = [&](int kc) {
        const int tg = kc / cchunks, c_off = (kc - tg * cchunks) * cpc;
        {   // A
            const int tap = tg * tpc + a_tap_sub;
            const int kti = tap / (g.kh * g.kw), r = tap - kti * g.kh * g.kw;
            const int khi = r / g.kw, kwi = r - khi * g.kw;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                int t, h, w;
                ra[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (a_ok[i] && tap < g.taps && src_coord<DGRAD>(a_t[i], kti, g.st, g.pt, g.ts, t) &&
                    src_coord<DGRAD>(a_h[i], khi, g.sh, g.ph, g.hs, h) && src_coord<DGRAD>(a_w[i], kwi, g.sw, g.pw, g.ws, w)) {
                    const size_t off = ((((size_t)a_n[i] * g.ts + t) * g.hs + h) * g.ws + w) * g.cs + c_off + a_c_sub;
                    ra[i] = __ldg(reinterpret_cast<const float4*>(src + off));
                }
            }
        }
#pragma unroll
        for (int j = 0; j < B4; ++j) {   // B: row k -> (tap, channel)
            const int k = b_row + (1024 / BN) * j;
            const int tap = tg * tpc + k / cpc, c = c_off + k % cpc;
            rb[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (tap < g.taps) rb[j] = __ldg(reinterpret_cast<const float4*>(filt + ((size_t)tap * g.cs + c) * g.cd + n0 + b_col));
        }
    };
    auto store_chunk = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int row = (tid >> 2) + 64 * i;
            As[buf][q * 4 + 0][row] = ra[i].x;
            As[buf][q * 4 + 1][row] = ra[i].y;
            As[buf][q * 4 + 2][row] = ra[i].z;
            As[buf][q * 4 + 3][row] = ra[i].w;
        }
#pragma unroll
        for (int j = 0; j < B4; ++j)
            *reinterpret_cast<float4*>(&Bs[buf][b_row + (1024 / BN) * j][b_col]) = rb[j];
    };

    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    load_chunk(0);
    store_chunk(0);
    __syncthreads();
    for (int kc = 0; kc < nchunks; ++kc) {
        const int cur = kc & 1;
        if (kc + 1 < nchunks) load_chunk(kc + 1);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[8], b[TN];
            *reinterpret_cast<float4*>(&a[0]) = *reinterpret_cast<const float4*>(&As[cur][k][ty * 8]);
            *reinterpret_cast<float4*>(&a[4]) = *reinterpret_cast<const float4*>(&As[cur][k][ty * 8 + 4]);
#pragma unroll
            for (int j = 0; j < TN; j += 4)
                *reinterpret_cast<float4*>(&b[j]) = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4 + 16 * j]);
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kc + 1 < nchunks) store_chunk(cur ^ 1);
        __syncthreads();
    }

    // epilogue: thread owns rows ty*8 + i and columns tx*4 + 64*(j/4) + (j%4)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + ty * 8 + i;
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < TN; j += 4) {
            const size_t off = (size_t)m * g.cd + n0 + tx * 4 + 16 * j;
            float4 v = make_float4(acc[i][j], acc[i][j + 1], acc[i][j + 2], acc[i][j + 3]);
            if (addend) {
                const float4 r = __ldg(reinterpret_cast<const float4*>(addend + off));
                v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
            }
            *reinterpret_cast<float4*>(dst + off) = v;
        }
    }
}

// ---- filter gradient -----------------------------------------------------------------------------
// GEMM rows i = (tap, ci) tile of 64, columns co tile of 64, reduction over destination pixels m.
constexpr int WBM = 64, WBN = 64, WBK = 16;

__global__ void __launch_bounds__(256) conv_wgrad_kernel(const Geom g, const float* __restrict__ src,
                                                         const float* __restrict__ dout, float* __restrict__ dfilt,
                                                         int m_per_split, int use_atomic) {
    __shared__ __align__(16) float As[2][WBK][WBM];
    __shared__ __align__(16) float Bs[2][WBK][WBN];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;

    const int cpm = g.cs < WBM ? g.cs : WBM;     // source channels per row tile
    const int tpm = WBM / cpm;                   // taps per row tile
    const int ctiles = g.cs / cpm;
    const int tg = blockIdx.x / ctiles, c_off = (blockIdx.x - tg * ctiles) * cpm;
    const int n0 = blockIdx.y * WBN;
    const int m_begin = blockIdx.z * m_per_split;
    const int m_end = min(g.M, m_begin + m_per_split);
    const int nchunks = (m_end - m_begin + WBK - 1) / WBK;

    // loader: pixel p = tid>>4 of the chunk, float4 index tid&15 along the row / column tile
    const int p = tid >> 4, f4 = (tid & 15) * 4;
    const int tap = tg * tpm + f4 / cpm, c = c_off + f4 % cpm;
    const int kti = tap / (g.kh * g.kw), rr = tap - kti * g.kh * g.kw;
    const int khi = rr / g.kw, kwi = rr - khi * g.kw;
    int m = m_begin + p;
    int pw_ = m % g.wd;  int tmp = m / g.wd;
    int ph_ = tmp % g.hd;  tmp /= g.hd;
    int pt_ = tmp % g.td;  int pn_ = tmp / g.td;

    float4 ra, rb;
    auto load_chunk = [&]() {
        ra = make_float4(0.f, 0.f, 0.f, 0.f);
        rb = ra;
        if (m < m_end) {
            int t, h, w;
            if (tap < g.taps && src_coord<false>(pt_, kti, g.st, g.pt, g.ts, t) && src_coord<false>(ph_, khi, g.sh, g.ph, g.hs, h) &&
                src_coord<false>(pw_, kwi, g.sw, g.pw, g.ws, w))
                ra = __ldg(reinterpret_cast<const float4*>(src + ((((size_t)pn_ * g.ts + t) * g.hs + h) * g.ws + w) * g.cs + c));
            rb = __ldg(reinterpret_cast<const float4*>(dout + (size_t)m * g.cd + n0 + f4));
        }
        // advance this thread's pixel by one chunk
        m += WBK;
        pw_ += WBK;
        while (pw_ >= g.wd) {
            pw_ -= g.wd;
            if (++ph_ == g.hd) {
                ph_ = 0;
                if (++pt_ == g.td) { pt_ = 0; ++pn_; }
            }
        }
    };
    auto store_chunk = [&](int buf) {
        *reinterpret_cast<float4*>(&As[buf][p][f4]) = ra;
        *reinterpret_cast<float4*>(&Bs[buf][p][f4]) = rb;
    };

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    if (nchunks > 0) {
        load_chunk();
        store_chunk(0);
    }
    __syncthreads();
    for (int kc = 0; kc < nchunks; ++kc) {
        const int cur = kc & 1;
        if (kc + 1 < nchunks) load_chunk();
#pragma unroll
        for (int k = 0; k < WBK; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kc + 1 < nchunks) store_chunk(cur ^ 1);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int il = ty * 4 + i;
        const int otap = tg * tpm + il / cpm, oc = c_off + il % cpm;
        if (otap >= g.taps) continue;
        float* o = dfilt + ((size_t)otap * g.cs + oc) * g.cd + n0 + tx * 4;
        if (use_atomic) {
#pragma unroll
            for (int j = 0; j < 4; ++j) atomicAdd(o + j, acc[i][j]);
        } else {
            *reinterpret_cast<float4*>(o) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        }
    }
}

static int make_geom(const avid_conv_shape_t* s, bool dgrad, Geom* g) {
    AVID_REQUIRE(s != nullptr, "conv: shape is NULL");
    AVID_REQUIRE(s->n > 0 && s->ti > 0 && s->hi > 0 && s->wi > 0 && s->to > 0 && s->ho > 0 && s->wo > 0, "conv: non-positive extent");
    AVID_REQUIRE(s->ci > 0 && s->ci % 4 == 0 && (s->ci < 16 ? s->ci == 4 || s->ci == 8 : s->ci % 16 == 0),
                 "conv: ci=%d must be 4, 8 or a multiple of 16 (pad the channels)", s->ci);
    AVID_REQUIRE(s->co > 0 && s->co % 64 == 0, "conv: co=%d must be a multiple of 64", s->co);
    AVID_REQUIRE(s->kt > 0 && s->kh > 0 && s->kw > 0 && s->st > 0 && s->sh > 0 && s->sw > 0 && s->pt >= 0 && s->ph >= 0 && s->pw >= 0, "conv: bad filter geometry");
    AVID_REQUIRE((s->ti + 2 * s->pt - s->kt) / s->st + 1 == s->to && (s->hi + 2 * s->ph - s->kh) / s->sh + 1 == s->ho &&
                 (s->wi + 2 * s->pw - s->kw) / s->sw + 1 == s->wo, "conv: output extent does not match input/filter/stride/padding");
    g->n = s->n;
    g->kt = s->kt; g->kh = s->kh; g->kw = s->kw;
    g->st = s->st; g->sh = s->sh; g->sw = s->sw;
    g->pt = s->pt; g->ph = s->ph; g->pw = s->pw;
    g->taps = s->kt * s->kh * s->kw;
    if (!dgrad) {
        g->td = s->to; g->hd = s->ho; g->wd = s->wo; g->cd = s->co;
        g->ts = s->ti; g->hs = s->hi; g->ws = s->wi; g->cs = s->ci;
    } else {
        g->td = s->ti; g->hd = s->hi; g->wd = s->wi; g->cd = s->ci;
        g->ts = s->to; g->hs = s->ho; g->ws = s->wo; g->cs = s->co;
    }
    const int64_t M = (int64_t)g->n * g->td * g->hd * g->wd;
    AVID_REQUIRE(M < ((int64_t)1 << 31) - 256, "conv: too many pixels");
    g->M = (int)M;
    return AVID_OK;
}

int conv_fp32_forward(const avid_conv_shape_t* s, const float* in, const float* filt, const float* addend, float* out, cudaStream_t st) {
    Geom g;
    int rc = make_geom(s, false, &g);
    if (rc) return rc;
    const unsigned gx = (g.M + BM - 1) / BM;
    if (g.cd % 128 == 0)
        conv_igemm_kernel<128, false><<<dim3(gx, g.cd / 128), 256, 0, st>>>(g, in, filt, addend, out);
    else
        conv_igemm_kernel<64, false><<<dim3(gx, g.cd / 64), 256, 0, st>>>(g, in, filt, addend, out);
    return check_launch("conv_igemm_kernel<fwd>");
}

int conv_fp32_dgrad(const avid_conv_shape_t* s, const float* dout, const float* filt_t, const float* addend, float* din, cudaStream_t st) {
    Geom g;
    int rc = make_geom(s, true, &g);
    if (rc) return rc;
    AVID_REQUIRE(g.cd % 64 == 0, "conv_dgrad: ci=%d must be a multiple of 64 (the stem has no input gradient)", g.cd);
    const unsigned gx = (g.M + BM - 1) / BM;
    if (g.cd % 128 == 0)
        conv_igemm_kernel<128, true><<<dim3(gx, g.cd / 128), 256, 0, st>>>(g, dout, filt_t, addend, din);
    else
        conv_igemm_kernel<64, true><<<dim3(gx, g.cd / 64), 256, 0, st>>>(g, dout, filt_t, addend, din);
    return check_launch("conv_igemm_kernel<dgrad>");
}

int conv_fp32_wgrad(const avid_conv_shape_t* s, const float* in, const float* dout, float* dfilt, cudaStream_t st) {
    Geom g;
    int rc = make_geom(s, false, &g);
    if (rc) return rc;
    AVID_REQUIRE(g.cs < WBM ? WBM % g.cs == 0 : g.cs % WBM == 0, "conv_wgrad: ci=%d must divide or be a multiple of 64", g.cs);
    const int cpm = g.cs < WBM ? g.cs : WBM, tpm = WBM / cpm;
    const unsigned gx = ((g.taps + tpm - 1) / tpm) * (g.cs / cpm), gy = g.cd / WBN;
    int splits = (4 * kNumSMs + (int)(gx * gy) - 1) / (int)(gx * gy);
    const int max_splits = (g.M + 8 * WBK - 1) / (8 * WBK);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    int m_per_split = (g.M + splits - 1) / splits;
    m_per_split = ((m_per_split + WBK - 1) / WBK) * WBK;
    splits = (g.M + m_per_split - 1) / m_per_split;
    conv_wgrad_kernel<<<dim3(gx, gy, splits), 256, 0, st>>>(g, in, dout, dfilt, m_per_split, splits > 1 ? 1 : 0);
    return check_launch("conv_wgrad_kernel");
}

}  // namespace avid
