// C-ABI entry points of the convolution family (include/avid_b200.h): dispatch on the math mode.
#include "common.cuh"

namespace avid {
int conv_fp32_forward(const avid_conv_shape_t* s, const float* in, const float* filt, const float* addend, float* out, cudaStream_t st);
int conv_fp32_dgrad(const avid_conv_shape_t* s, const float* dout, const float* filt_t, const float* addend, float* din, cudaStream_t st);
int conv_fp32_wgrad(const avid_conv_shape_t* s, const float* in, const float* dout, float* dfilt, cudaStream_t st);
}  // namespace avid

using namespace avid;

extern "C" {

int avid_conv_forward(const avid_conv_shape_t* s, const float* in, const float* filt, const float* addend, float* out, int32_t math, void* stream) {
    AVID_REQUIRE(in && filt && out, "conv_forward: NULL pointer");
    if (math == AVID_MATH_FP32) return conv_fp32_forward(s, in, filt, addend, out, static_cast<cudaStream_t>(stream));
    set_error("conv_forward: math mode %d is not built into this library", math);
    return AVID_EUNSUPPORTED;
}

int avid_conv_dgrad(const avid_conv_shape_t* s, const float* dout, const float* filt_t, const float* addend, float* din, int32_t math, void* stream) {
    AVID_REQUIRE(dout && filt_t && din, "conv_dgrad: NULL pointer");
    if (math == AVID_MATH_FP32) return conv_fp32_dgrad(s, dout, filt_t, addend, din, static_cast<cudaStream_t>(stream));
    set_error("conv_dgrad: math mode %d is not built into this library", math);
    return AVID_EUNSUPPORTED;
}

int avid_conv_wgrad(const avid_conv_shape_t* s, const float* in, const float* dout, float* dfilt, int32_t math, void* stream) {
    AVID_REQUIRE(in && dout && dfilt, "conv_wgrad: NULL pointer");
    if (math == AVID_MATH_FP32) return conv_fp32_wgrad(s, in, dout, dfilt, static_cast<cudaStream_t>(stream));
    set_error("conv_wgrad: math mode %d is not built into this library", math);
    return AVID_EUNSUPPORTED;
}

}  // extern "C"
