// Video input step on the GPU (include/avid_b200.h, avid_video_prep): what the reference's CPU loader workers do to every clip with
// Pillow through torchvision (datasets/preprocessing.py:15-57, VideoPrep_MSC_CJ with augment=True; utils/videotransforms/
// video_transforms.py:73-98,303-391,393-476, volume_transforms.py:14-70, tensor_transforms.py:13-38):
//     crop (top, left, h, w) -> Image.resize((out_w, out_h), BILINEAR) -> optional horizontal flip
//     -> brightness / saturation / hue / contrast in the drawn order (uint8 images between the ops) -> float / 255 -> (x - mean) / std
// for one clip of T uint8 RGB frames, BIT-IDENTICAL to Pillow 12 (oracle/video.py pins the arithmetic against Pillow itself and
// against goldens of the unmodified reference classes).  The arithmetic follows Pillow's C sources:
//   * Resample.c: separable convolution with a triangle filter whose support grows with the down-scaling factor, weights normalised in
//     double, rounded to 22-bit fixed point, horizontal pass first with a uint8 intermediate image (precompute_coeffs,
//     normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc / Vertical_8bpc).  The weights are computed ON THE DEVICE in IEEE double
//     with explicitly rounded operations (no FMA contraction), so no host table, copy or synchronisation is needed;
//   * Blend.c (ImageEnhance): out = (UINT8)(d + alpha * (x - d)) in float32, clipped when alpha is outside [0, 1]; the degenerate
//     image d is black (brightness), the luma (saturation) or the frame's mean luma int(mean + 0.5) (contrast -> a per-frame
//     reduction, which is why the ops before / from the contrast op run in two kernels);
//   * Convert.c: L = (19595 R + 38470 G + 7471 B + 0x8000) >> 16; rgb2hsv_row / hsv2rgb with float variables and double constants.
// Layout: frames (T, H, W, 3) uint8 -> out (3, T, out_h, out_w) float32 (ClipToTensor's C x T x H x W).  HBM-bound byte work: one thread
// per output pixel, coalesced along the width.
#include "common.cuh"

namespace avid {

constexpr int kPrecisionBits = 32 - 8 - 2;      // Resample.c PRECISION_BITS
constexpr int kMaxTaps = 64;                    // ksize limit (down-scaling by up to ~31x)

struct VideoPrepArgs {
    avid_video_prep_t p;
    int ksize_x, ksize_y;
    int first_ops;              // ops [0, first_ops) run in the vertical-pass kernel, [first_ops, num_ops) in the finishing kernel (contrast first)
    // per-clip pointers (workspace regions of this clip)
    const uint8_t* frames;
    float* out;
    int *bx, *kx, *by, *ky;
    unsigned int* lsum;
    uint8_t *tmp, *img;
};

constexpr int kVideoBatch = 16;                 // clips per launch: their descriptors travel by value in the kernel parameters (blockIdx.z = clip)
struct VideoBatch {
    VideoPrepArgs c[kVideoBatch];
};

// ---- Resample.c precompute_coeffs + normalize_coeffs_8bpc, one thread per output coordinate ----
__global__ void __launch_bounds__(128) video_coeffs_kernel(const VideoBatch b) {
    const VideoPrepArgs& a = b.c[blockIdx.z];
    const bool y_axis = blockIdx.y != 0;
    const int in_size = y_axis ? a.p.crop_h : a.p.crop_w, out_size = y_axis ? a.p.out_h : a.p.out_w, ksize = y_axis ? a.ksize_y : a.ksize_x;
    int* const bounds = y_axis ? a.by : a.bx;
    int* const kk = y_axis ? a.ky : a.kx;
    if (blockIdx.x == 0 && blockIdx.y == 0)                          // the luma sums of the contrast op start at zero (before any thread leaves)
        for (int t = threadIdx.x; t < a.p.frames; t += blockDim.x) a.lsum[t] = 0u;
    const int xx = blockIdx.x * blockDim.x + threadIdx.x;
    if (xx >= out_size) return;
    const double scale = __ddiv_rn((double)in_size, (double)out_size);
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = filterscale;                               // bilinear: support 1.0 * filterscale
    const double ss = __ddiv_rn(1.0, filterscale);
    const double center = __dmul_rn((double)xx + 0.5, scale);         // in0 = 0
    int xmin = (int)__dadd_rn(__dsub_rn(center, support), 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)__dadd_rn(__dadd_rn(center, support), 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
        double t = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss);
        if (t < 0.0) t = -t;
        ww = __dadd_rn(ww, t < 1.0 ? __dsub_rn(1.0, t) : 0.0);
    }
    int* k = kk + (size_t)xx * ksize;
    for (int x = 0; x < ksize; ++x) {
        double w = 0.0;
        if (x < xmax) {
            double t = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss);
            if (t < 0.0) t = -t;
            w = t < 1.0 ? __dsub_rn(1.0, t) : 0.0;
            if (ww != 0.0) w = __ddiv_rn(w, ww);
        }
        const double v = __dmul_rn(w, (double)(1 << kPrecisionBits));
        k[x] = w < 0 ? (int)__dadd_rn(-0.5, v) : (int)__dadd_rn(0.5, v);
    }
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
}

__device__ __forceinline__ uint8_t clip8(int v) {      // Resample.c clip8: lookup of (ss >> PRECISION_BITS) clamped to [0, 255]
    v >>= kPrecisionBits;
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// horizontal pass over the crop: tmp[t][y][ox][c], y in [0, crop_h)
__global__ void __launch_bounds__(256) video_resample_h_kernel(const VideoBatch bt) {
    const VideoPrepArgs& a = bt.c[blockIdx.z];
    const avid_video_prep_t& p = a.p;
    const uint8_t* __restrict__ frames = a.frames;
    const int* __restrict__ bounds = a.bx;
    const int* __restrict__ kk = a.kx;
    uint8_t* __restrict__ tmp = a.tmp;
    const int64_t total = (int64_t)p.frames * p.crop_h * p.out_w;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int ox = (int)(i % p.out_w);
        const int64_t r = i / p.out_w;
        const int y = (int)(r % p.crop_h), t = (int)(r / p.crop_h);
        const int xmin = bounds[2 * ox], cnt = bounds[2 * ox + 1];
        const int* k = kk + (size_t)ox * a.ksize_x;
        const uint8_t* src = frames + (((size_t)t * p.height + (p.crop_top + y)) * p.width + (p.crop_left + xmin)) * 3;
        int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
        for (int x = 0; x < cnt; ++x) {
            const int w = k[x];
            s0 += (int)src[3 * x] * w;
            s1 += (int)src[3 * x + 1] * w;
            s2 += (int)src[3 * x + 2] * w;
        }
        uint8_t* dst = tmp + (size_t)i * 3;
        dst[0] = clip8(s0);  dst[1] = clip8(s1);  dst[2] = clip8(s2);
    }
}

// ---- colour arithmetic ----
__device__ __forceinline__ int luma(int r, int g, int b) { return (r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16; }

// Blend.c: im1 = degenerate value d, im2 = image value x
__device__ __forceinline__ int blend(int d, int x, float alpha) {
    if (alpha == 0.0f) return d;
    if (alpha == 1.0f) return x;
    const float t = __fadd_rn((float)d, __fmul_rn(alpha, (float)(x - d)));
    if (alpha >= 0.0f && alpha <= 1.0f) return (int)t & 0xFF;           // (UINT8) of a value in [0, 255]
    if (t <= 0.0f) return 0;
    if (t >= 255.0f) return 255;
    return (int)t;
}

__device__ __forceinline__ int clip255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

// Convert.c rgb2hsv_row
__device__ __forceinline__ void rgb2hsv(int r, int g, int b, int& uh, int& us, int& uv) {
    const int maxc = max(r, max(g, b)), minc = min(r, min(g, b));
    uv = maxc;
    if (minc == maxc) { uh = 0; us = 0; return; }
    const float cr = (float)(maxc - minc);
    const float s = __fdiv_rn(cr, (float)maxc);
    const float rc = __fdiv_rn((float)(maxc - r), cr), gc = __fdiv_rn((float)(maxc - g), cr), bc = __fdiv_rn((float)(maxc - b), cr);
    float h;
    if (r == maxc) h = __fsub_rn(bc, gc);
    else if (g == maxc) h = __double2float_rn(__dsub_rn(__dadd_rn(2.0, (double)rc), (double)bc));
    else h = __double2float_rn(__dsub_rn(__dadd_rn(4.0, (double)gc), (double)rc));
    const double hh = __dadd_rn(__ddiv_rn((double)h, 6.0), 1.0);       // in [0.83, 1.84): fmod(hh, 1.0) = hh - floor(hh), exact
    h = __double2float_rn(hh - floor(hh));
    uh = clip255((int)__dmul_rn((double)h, 255.0));
    us = clip255((int)__dmul_rn((double)s, 255.0));
}

__device__ __forceinline__ int c_round(double a) { return (int)(a >= 0.0 ? floor(__dadd_rn(a, 0.5)) : ceil(__dsub_rn(a, 0.5))); }

// Convert.c hsv2rgb
__device__ __forceinline__ void hsv2rgb(int h, int s, int v, int& r, int& g, int& b) {
    if (s == 0) { r = g = b = v; return; }
    const double hf = __ddiv_rn(__dmul_rn((double)h, 6.0), 255.0);
    const int i = (int)floor(hf);
    const double f = (double)__double2float_rn(__dsub_rn(hf, (double)i));
    const double fs = (double)__double2float_rn(__ddiv_rn((double)s, 255.0));
    const double vf = (double)v;
    const int p = clip255(c_round(__dmul_rn(vf, __dsub_rn(1.0, fs))));
    const int q = clip255(c_round(__dmul_rn(vf, __dsub_rn(1.0, __dmul_rn(fs, f)))));
    const int t = clip255(c_round(__dmul_rn(vf, __dsub_rn(1.0, __dmul_rn(fs, __dsub_rn(1.0, f))))));
    switch (i % 6) {
        case 0: r = v; g = t; b = p; break;
        case 1: r = q; g = v; b = p; break;
        case 2: r = p; g = v; b = t; break;
        case 3: r = p; g = q; b = v; break;
        case 4: r = t; g = p; b = v; break;
        default: r = v; g = p; b = q; break;
    }
}

// one jitter op on one pixel; `mean_l` = the frame's int(mean luma + 0.5) for the contrast op
__device__ __forceinline__ void apply_op(int kind, float f, int hue_shift, int mean_l, int& r, int& g, int& b) {
    if (kind == AVID_VIDEO_OP_BRIGHTNESS) {
        r = blend(0, r, f);  g = blend(0, g, f);  b = blend(0, b, f);
    } else if (kind == AVID_VIDEO_OP_SATURATION) {
        const int l = luma(r, g, b);
        r = blend(l, r, f);  g = blend(l, g, f);  b = blend(l, b, f);
    } else if (kind == AVID_VIDEO_OP_CONTRAST) {
        r = blend(mean_l, r, f);  g = blend(mean_l, g, f);  b = blend(mean_l, b, f);
    } else {                                        // hue: uint8 shift of H by np.uint8(hue_factor * 255) (computed by the host in double), wraps
        int h, s, v;
        rgb2hsv(r, g, b, h, s, v);
        h = (h + hue_shift) & 0xFF;
        hsv2rgb(h, s, v, r, g, b);
    }
}

// vertical pass + flip + the ops before the contrast op; accumulates the per-frame luma sums the contrast op needs
__global__ void __launch_bounds__(256) video_resample_v_kernel(const VideoBatch bt) {
    const VideoPrepArgs& a = bt.c[blockIdx.z];
    const avid_video_prep_t& p = a.p;
    const uint8_t* __restrict__ tmp = a.tmp;
    const int* __restrict__ bounds = a.by;
    const int* __restrict__ kk = a.ky;
    uint8_t* __restrict__ img = a.img;
    unsigned int* __restrict__ luma_sum = a.lsum;
    const int per_frame = p.out_h * p.out_w;
    const int t = blockIdx.y;
    if (t >= p.frames) return;
    unsigned int lsum = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < per_frame; i += gridDim.x * blockDim.x) {
        const int ox = i % p.out_w, oy = i / p.out_w;
        const int ymin = bounds[2 * oy], cnt = bounds[2 * oy + 1];
        const int* k = kk + (size_t)oy * a.ksize_y;
        const int sx = p.flip ? p.out_w - 1 - ox : ox;                 // FLIP_LEFT_RIGHT after the resize
        const uint8_t* src = tmp + (((size_t)t * p.crop_h + ymin) * p.out_w + sx) * 3;
        int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
        for (int y = 0; y < cnt; ++y) {
            const int w = k[y];
            const uint8_t* q = src + (size_t)y * p.out_w * 3;
            s0 += (int)q[0] * w;  s1 += (int)q[1] * w;  s2 += (int)q[2] * w;
        }
        int r = clip8(s0), g = clip8(s1), b = clip8(s2);
        for (int o = 0; o < a.first_ops; ++o) apply_op(p.op_kind[o], p.op_factor[o], p.hue_shift & 0xFF, 0, r, g, b);
        uint8_t* dst = img + ((size_t)t * per_frame + i) * 3;
        dst[0] = (uint8_t)r;  dst[1] = (uint8_t)g;  dst[2] = (uint8_t)b;
        lsum += (unsigned int)luma(r, g, b);
    }
    if (a.first_ops < p.num_ops) {                  // a contrast op follows: frame sum of L (at most 255 * out_h * out_w < 2^32, checked on the host)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
        if ((threadIdx.x & 31) == 0 && lsum) atomicAdd(luma_sum + t, lsum);
    }
}

// the contrast op and the ops after it, then ClipToTensor (/ 255) and Normalize ((x - mean) / std): out (3, T, out_h, out_w) float32
__global__ void __launch_bounds__(256) video_finish_kernel(const VideoBatch bt) {
    const VideoPrepArgs& a = bt.c[blockIdx.z];
    const avid_video_prep_t& p = a.p;
    const uint8_t* __restrict__ img = a.img;
    const unsigned int* __restrict__ luma_sum = a.lsum;
    float* __restrict__ out = a.out;
    const int per_frame = p.out_h * p.out_w;
    const int t = blockIdx.y;
    if (t >= p.frames) return;
    int mean_l = 0;
    if (a.first_ops < p.num_ops)                    // ImageEnhance.Contrast: int(ImageStat.Stat(L).mean[0] + 0.5)
        mean_l = (int)__dadd_rn(__ddiv_rn((double)luma_sum[t], (double)per_frame), 0.5);
    const size_t plane = (size_t)p.frames * per_frame;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < per_frame; i += gridDim.x * blockDim.x) {
        const uint8_t* src = img + ((size_t)t * per_frame + i) * 3;
        int c[3] = {src[0], src[1], src[2]};
        for (int o = a.first_ops; o < p.num_ops; ++o) apply_op(p.op_kind[o], p.op_factor[o], p.hue_shift & 0xFF, mean_l, c[0], c[1], c[2]);
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            float v = __fdiv_rn((float)c[ch], 255.0f);
            if (p.normalize) v = __fdiv_rn(__fsub_rn(v, p.mean[ch]), p.std[ch]);
            out[ch * plane + (size_t)t * per_frame + i] = v;
        }
    }
}

static int ksize_of(int in_size, int out_size) {
    double scale = (double)in_size / out_size;
    if (scale < 1.0) scale = 1.0;
    return (int)ceil(scale) * 2 + 1;
}

static int plan(const avid_video_prep_t* p, VideoPrepArgs& a, size_t off[6]) {
    AVID_REQUIRE(p, "video_prep: NULL parameters");
    AVID_REQUIRE(p->frames > 0 && p->height > 0 && p->width > 0 && p->out_h > 0 && p->out_w > 0, "video_prep: empty clip or output");
    AVID_REQUIRE(p->crop_h > 0 && p->crop_w > 0 && p->crop_top >= 0 && p->crop_left >= 0 && p->crop_top + p->crop_h <= p->height &&
                     p->crop_left + p->crop_w <= p->width,
                 "video_prep: crop box (%d, %d, %d, %d) outside the %d x %d frame", p->crop_top, p->crop_left, p->crop_h, p->crop_w, p->height, p->width);
    AVID_REQUIRE(p->num_ops >= 0 && p->num_ops <= 4, "video_prep: at most 4 colour ops");
    AVID_REQUIRE((int64_t)p->out_h * p->out_w < (1 << 24) && (int64_t)p->frames * p->height * p->width < ((int64_t)1 << 31), "video_prep: clip too large");
    a.p = *p;
    a.ksize_x = ksize_of(p->crop_w, p->out_w);
    a.ksize_y = ksize_of(p->crop_h, p->out_h);
    AVID_REQUIRE(a.ksize_x <= kMaxTaps && a.ksize_y <= kMaxTaps, "video_prep: down-scaling factor too large");
    int n_contrast = 0;
    a.first_ops = p->num_ops;
    for (int o = 0; o < p->num_ops; ++o) {
        AVID_REQUIRE(p->op_kind[o] >= AVID_VIDEO_OP_BRIGHTNESS && p->op_kind[o] <= AVID_VIDEO_OP_CONTRAST, "video_prep: unknown op %d", p->op_kind[o]);
        if (p->op_kind[o] == AVID_VIDEO_OP_HUE) AVID_REQUIRE(p->op_factor[o] >= -0.5f && p->op_factor[o] <= 0.5f, "hue_factor is not in [-0.5, 0.5].");
        if (p->op_kind[o] == AVID_VIDEO_OP_CONTRAST) {
            if (n_contrast++ == 0) a.first_ops = o;
        }
    }
    AVID_REQUIRE(n_contrast <= 1, "video_prep: at most one contrast op");
    // workspace: x bounds | x weights | y bounds | y weights | luma sums | horizontal-pass image | resized image
    size_t o = 0;
    auto take = [&](size_t bytes) { const size_t at = o; o += (bytes + 255) & ~(size_t)255; return at; };
    off[0] = take((size_t)p->out_w * 2 * 4 + (size_t)p->out_w * a.ksize_x * 4);
    off[1] = take((size_t)p->out_h * 2 * 4 + (size_t)p->out_h * a.ksize_y * 4);
    off[2] = take((size_t)p->frames * 4);
    off[3] = take((size_t)p->frames * p->crop_h * p->out_w * 3);
    off[4] = take((size_t)p->frames * p->out_h * p->out_w * 3);
    off[5] = o;
    return AVID_OK;
}

// bind one clip's pointers to its workspace region
static void bind(VideoPrepArgs& a, const size_t off[6], const uint8_t* frames, float* out, uint8_t* ws) {
    a.frames = frames;  a.out = out;
    a.bx = reinterpret_cast<int*>(ws + off[0]);  a.kx = a.bx + 2 * a.p.out_w;
    a.by = reinterpret_cast<int*>(ws + off[1]);  a.ky = a.by + 2 * a.p.out_h;
    a.lsum = reinterpret_cast<unsigned int*>(ws + off[2]);
    a.tmp = ws + off[3];
    a.img = ws + off[4];
}

// up to kVideoBatch clips per launch sequence: 4 launches whatever the clip count
static int launch_group(const VideoBatch& b, int count, cudaStream_t st) {
    int max_out = 0, max_frames = 0, rc;
    int64_t max_h = 0;
    int max_pf = 0;
    for (int i = 0; i < count; ++i) {
        const avid_video_prep_t& p = b.c[i].p;
        max_out = max(max_out, max(p.out_w, p.out_h));
        max_frames = max(max_frames, p.frames);
        max_h = max(max_h, (int64_t)p.frames * p.crop_h * p.out_w);
        max_pf = max(max_pf, p.out_h * p.out_w);
    }
    video_coeffs_kernel<<<dim3((max_out + 127) / 128, 2, count), 128, 0, st>>>(b);
    if ((rc = check_launch("video_coeffs_kernel"))) return rc;
    int blocks = (int)((max_h + 255) / 256);
    if (blocks > 32 * kNumSMs) blocks = 32 * kNumSMs;
    video_resample_h_kernel<<<dim3(blocks, 1, count), 256, 0, st>>>(b);
    if ((rc = check_launch("video_resample_h_kernel"))) return rc;
    int bpf = (max_pf + 255) / 256;
    if (bpf > 4 * kNumSMs) bpf = 4 * kNumSMs;
    video_resample_v_kernel<<<dim3(bpf, max_frames, count), 256, 0, st>>>(b);
    if ((rc = check_launch("video_resample_v_kernel"))) return rc;
    video_finish_kernel<<<dim3(bpf, max_frames, count), 256, 0, st>>>(b);
    return check_launch("video_finish_kernel");
}

}  // namespace avid

using namespace avid;

extern "C" {

size_t avid_video_prep_workspace_bytes(const avid_video_prep_t* p) {
    VideoPrepArgs a;
    size_t off[6];
    return plan(p, a, off) == AVID_OK ? off[5] : 0;
}

int avid_video_prep(const uint8_t* frames, const avid_video_prep_t* p, float* out, void* workspace, size_t workspace_bytes, void* stream) {
    const uint8_t* f[1] = {frames};
    float* o[1] = {out};
    return avid_video_prep_batch(f, p, 1, o, workspace, workspace_bytes, stream);
}

size_t avid_video_prep_batch_workspace_bytes(const avid_video_prep_t* params, int32_t count) {
    size_t total = 0;
    for (int i = 0; params && i < count; ++i) {
        const size_t one = avid_video_prep_workspace_bytes(params + i);
        if (one == 0) return 0;
        total += one;
    }
    return total;
}

int avid_video_prep_batch(const uint8_t* const* frames, const avid_video_prep_t* params, int32_t count, float* const* out, void* workspace,
                          size_t workspace_bytes, void* stream) {
    AVID_REQUIRE(count >= 0, "video_prep: negative clip count");
    AVID_REQUIRE(count == 0 || (frames && params && out && workspace), "video_prep: NULL pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    size_t used = 0;
    for (int base = 0; base < count; base += kVideoBatch) {
        VideoBatch b;
        const int n = min(kVideoBatch, count - base);
        for (int i = 0; i < n; ++i) {
            size_t off[6];
            const int rc = plan(params + base + i, b.c[i], off);
            if (rc) return rc;
            AVID_REQUIRE(frames[base + i] && out[base + i], "video_prep: NULL clip pointer");
            AVID_REQUIRE(used + off[5] <= workspace_bytes, "video_prep: workspace of %zu bytes is too small", workspace_bytes);
            bind(b.c[i], off, frames[base + i], out[base + i], ws + used);
            used += off[5];
        }
        for (int i = n; i < kVideoBatch; ++i) b.c[i] = b.c[0];
        const int rc = launch_group(b, n, st);
        if (rc) return rc;
    }
    return AVID_OK;
}

}  // extern "C"
