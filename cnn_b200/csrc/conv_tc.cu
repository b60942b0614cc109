// conv_tc.cu -- tcgen05 implicit-GEMM convolution (placeholder until the kernels land).
#include "common.cuh"

bool conv_tc_supported(int, int, int, int) { return false; }
int conv_fwd_tc(cnn_ctx*, const float*, const float*, const float*, float*, int, int, int, int, int, int, int) {
    cnn_set_error("tcgen05 conv forward not built");
    return CNN_ERR_UNSUPPORTED;
}
int conv_wgrad_tc(cnn_ctx*, const float*, const float*, float*, float*, int, int, int, int, int, int, int, float) {
    cnn_set_error("tcgen05 conv wgrad not built");
    return CNN_ERR_UNSUPPORTED;
}
int conv_dgrad_tc(cnn_ctx*, const float*, const float*, float*, int, int, int, int, int, int, int) {
    cnn_set_error("tcgen05 conv dgrad not built");
    return CNN_ERR_UNSUPPORTED;
}
