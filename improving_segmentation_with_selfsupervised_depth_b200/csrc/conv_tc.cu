// placeholder — replaced by the tcgen05 implicit-GEMM family
#include "common.cuh"
extern "C" int segsde_tc_available(void) { return 0; }
extern "C" int segsde_conv2d_fwd_tc(const segsde_nhwc_t*, const segsde_nhwc_t*, const float*, const float*,
                                    const segsde_nhwc_t*, const segsde_conv_desc_t*, void*) { return SEGSDE_E_UNSUPPORTED; }
extern "C" int segsde_conv2d_dgrad_tc(const segsde_nhwc_t*, const float*, const segsde_nhwc_t*,
                                      const segsde_nhwc_t*, const segsde_conv_desc_t*, void*) { return SEGSDE_E_UNSUPPORTED; }
extern "C" int segsde_conv2d_wgrad_tc(const segsde_nhwc_t*, const segsde_nhwc_t*, const segsde_nhwc_t*, float*,
                                      float*, const segsde_conv_desc_t*, void*) { return SEGSDE_E_UNSUPPORTED; }
