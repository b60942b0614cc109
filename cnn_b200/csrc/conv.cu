// conv.cu -- C-ABI entry points of Conv2D and the algorithm dispatch.
#include <cstdlib>

#include "common.cuh"

namespace {
int check(const char* who, int B, int Cin, int H, int W, int Cout, int k, int s) {
    // the reference asserts k odd and >= 3 (conv2d.cpp:14); k == 1 is accepted as well
    CNN_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && s > 0 && k > 0 && (k & 1), "%s: bad conv parameters", who);
    CNN_REQUIRE(H >= k && W >= k, "%s: input %dx%d smaller than kernel %d", who, H, W, k);
    return CNN_OK;
}
// AUTO picks per operator from B200 measurements (profiles/): the tcgen05 implicit GEMM wins
// once the layer has enough input channels to fill K blocks; for the first layer (Cin <= 4,
// K = 27, 3 gradient channels) building split operand tiles costs more than the arithmetic, so
// forward and input gradient run on fp32 CUDA cores there -- the TMA-staged constant-bank kernels of
// conv_thin.cu for the reference's 3 -> 16 layer, the generic SIMT kernels otherwise (the TMA
// row-staged tensor-core weight gradient already wins).  CNN_CONV_TCGEN05 forces tensor cores.
bool use_tc(const cnn_ctx* ctx, int Cin, int Cout, int k, int s, bool wgrad = false) {
    if (ctx->conv_algo == CNN_CONV_SIMT) return false;
    if (!conv_tc_supported(Cin, Cout, k, s)) return false;
    if (ctx->conv_algo == CNN_CONV_AUTO && Cin <= 4 && !(wgrad && k == 3)) return false;
    return true;
}
}  // namespace

extern "C" {

int cnn_conv2d_forward(cnn_ctx* ctx, const float* x, const float* w, const float* bias, float* y, int B,
                       int Cin, int H, int W, int Cout, int k, int stride) {
    CNN_REQUIRE(ctx && x && w && bias && y, "cnn_conv2d_forward: NULL argument");
    if (int rc = check("cnn_conv2d_forward", B, Cin, H, W, Cout, k, stride)) return rc;
    if (ctx->conv_algo == CNN_CONV_AUTO && conv_thin_supported(ctx, Cin, H, W, Cout, k, stride))
        return conv_fwd_thin(ctx, x, w, bias, y, B, H, W);
    if (ctx->conv_algo == CNN_CONV_AUTO && conv_s2_supported(ctx, Cin, H, W, Cout, k, stride))
        return conv_fwd_s2(ctx, x, w, bias, y, nullptr, B, Cin, H, W, Cout);
    if (ctx->conv_algo == CNN_CONV_AUTO && conv_s1_supported(ctx, Cin, H, W, Cout, k, stride))
        return conv_fwd_s1(ctx, x, w, bias, y, nullptr, B, Cin, H, W, Cout);
    if (use_tc(ctx, Cin, Cout, k, stride)) return conv_fwd_tc(ctx, x, w, bias, y, B, Cin, H, W, Cout, k, stride);
    if (ctx->conv_algo == CNN_CONV_TCGEN05) {
        cnn_set_error("cnn_conv2d_forward: shape not supported by the tcgen05 path");
        return CNN_ERR_UNSUPPORTED;
    }
    return conv_fwd_simt(ctx, x, w, bias, y, B, Cin, H, W, Cout, k, stride);
}

int cnn_conv2d_relu_maxpool_forward(cnn_ctx* ctx, const float* x, const float* w, const float* bias, float* y_conv,
                                    float* y_relu, float* y_pool, int32_t* mask, int B, int Cin, int H, int W, int Cout,
                                    int k, int stride, int pool_k, int pool_step) {
    CNN_REQUIRE(ctx && x && w && bias && y_conv && y_relu && y_pool, "cnn_conv2d_relu_maxpool_forward: NULL argument");
    if (int rc = check("cnn_conv2d_relu_maxpool_forward", B, Cin, H, W, Cout, k, stride)) return rc;
    if (ctx->conv_algo != CNN_CONV_AUTO || !conv_thin_pool_supported(ctx, Cin, H, W, Cout, k, stride, pool_k, pool_step)) {
        cnn_set_error("cnn_conv2d_relu_maxpool_forward: shape not served by the fused kernel");
        return CNN_ERR_UNSUPPORTED;
    }
    return conv_fwd_thin_relu_pool(ctx, x, w, bias, y_conv, y_relu, y_pool, mask, B, H, W);
}

int cnn_conv2d_backward_weights(cnn_ctx* ctx, const float* x, const float* delta, float* dw, float* db,
                                int B, int Cin, int H, int W, int Cout, int k, int stride, float scale) {
    CNN_REQUIRE(ctx && x && delta && dw && db, "cnn_conv2d_backward_weights: NULL argument");
    if (int rc = check("cnn_conv2d_backward_weights", B, Cin, H, W, Cout, k, stride)) return rc;
    if (ctx->conv_algo == CNN_CONV_AUTO && conv_thin_supported(ctx, Cin, H, W, Cout, k, stride) && !getenv("CNN_DBG_NOTHINWG"))
        return conv_wgrad_thin(ctx, x, delta, dw, db, B, H, W, scale);
    if (ctx->conv_algo == CNN_CONV_AUTO && conv_s2_supported(ctx, Cin, H, W, Cout, k, stride))
        return conv_wgrad_s2(ctx, x, delta, dw, db, B, Cin, H, W, Cout, scale);
    if (ctx->conv_algo == CNN_CONV_AUTO && conv_s1_supported(ctx, Cin, H, W, Cout, k, stride) && !getenv("CNN_DBG_NOS1WG"))
        return conv_wgrad_s1(ctx, x, delta, dw, db, B, Cin, H, W, Cout, scale);
    if (use_tc(ctx, Cin, Cout, k, stride, true))
        return conv_wgrad_tc(ctx, x, delta, dw, db, B, Cin, H, W, Cout, k, stride, scale);
    if (ctx->conv_algo == CNN_CONV_TCGEN05) {
        cnn_set_error("cnn_conv2d_backward_weights: shape not supported by the tcgen05 path");
        return CNN_ERR_UNSUPPORTED;
    }
    return conv_wgrad_simt(ctx, x, delta, dw, db, B, Cin, H, W, Cout, k, stride, scale);
}

int cnn_conv2d_backward_data(cnn_ctx* ctx, const float* w, const float* delta, float* dx, int B, int Cin,
                             int H, int W, int Cout, int k, int stride) {
    CNN_REQUIRE(ctx && w && delta && dx, "cnn_conv2d_backward_data: NULL argument");
    if (int rc = check("cnn_conv2d_backward_data", B, Cin, H, W, Cout, k, stride)) return rc;
    if (ctx->conv_algo == CNN_CONV_AUTO && conv_thin_supported(ctx, Cin, H, W, Cout, k, stride))
        return conv_dgrad_thin(ctx, w, delta, dx, B, H, W);
    if (ctx->conv_algo == CNN_CONV_AUTO && conv_s2_supported(ctx, Cin, H, W, Cout, k, stride))
        return conv_dgrad_s2(ctx, w, delta, dx, nullptr, B, Cin, H, W, Cout);
    if (ctx->conv_algo == CNN_CONV_AUTO && conv_s1_supported(ctx, Cin, H, W, Cout, k, stride))
        return conv_dgrad_s1(ctx, w, delta, dx, nullptr, B, Cin, H, W, Cout);
    if (use_tc(ctx, Cin, Cout, k, stride)) return conv_dgrad_tc(ctx, w, delta, dx, B, Cin, H, W, Cout, k, stride);
    if (ctx->conv_algo == CNN_CONV_TCGEN05) {
        cnn_set_error("cnn_conv2d_backward_data: shape not supported by the tcgen05 path");
        return CNN_ERR_UNSUPPORTED;
    }
    return conv_dgrad_simt(ctx, w, delta, dx, B, Cin, H, W, Cout, k, stride);
}

}  // extern "C"
