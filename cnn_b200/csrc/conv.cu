// conv.cu -- C-ABI entry points of Conv2D and the algorithm dispatch.
#include <algorithm>
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
// ---- 1x1 convolutions with stride > 1 (BASELINE config 5's transition layers) -------------------------------------
// y[b][o][oy][ox] = bias[o] + sum_i x[b][i][oy*s][ox*s] * w[o][i]: only every s-th row / column of the input takes part.
// The sampled pixels are gathered once into a dense [B][Cin][OH][OW] tensor and the three passes become the stride-1
// 1x1 GEMMs of the tensor-core path; the input gradient is scattered back (cells no window covers stay 0).
__global__ void __launch_bounds__(256) subsample_kernel(const float* __restrict__ x, float* __restrict__ xs, int planes, int H, int W,
                                                         int OH, int OW, int s) {
    const size_t n = (size_t)planes * OH * OW;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int ox = (int)(i % OW);
        const size_t t = i / OW;
        const int oy = (int)(t % OH);
        const size_t pl = t / OH;
        xs[i] = x[(pl * H + (size_t)oy * s) * W + (size_t)ox * s];
    }
}
__global__ void __launch_bounds__(256) upsample_scatter_kernel(const float* __restrict__ dxs, float* __restrict__ dx, int planes, int H,
                                                                int W, int OH, int OW, int s) {
    const size_t n = (size_t)planes * H * W;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int xx = (int)(i % W);
        const size_t t = i / W;
        const int y = (int)(t % H);
        const size_t pl = t / H;
        const bool hit = y % s == 0 && xx % s == 0 && y / s < OH && xx / s < OW;
        dx[i] = hit ? dxs[(pl * OH + y / s) * OW + xx / s] : 0.f;
    }
}
bool use_sub1x1(const cnn_ctx* ctx, int Cin, int Cout, int k, int s) {
    return ctx->conv_algo != CNN_CONV_SIMT && k == 1 && s > 1 && conv_tc_supported(Cin, Cout, 1, 1) && Cin > 4;
}
float* subsampled(cnn_ctx* ctx, const float* x, int B, int Cin, int H, int W, int s, size_t extra_floats, float** extra) {
    const int OH = (H - 1) / s + 1, OW = (W - 1) / s + 1;
    const size_t n = (size_t)B * Cin * OH * OW;
    float* a = reinterpret_cast<float*>(cnn_arena(ctx, (n + extra_floats) * sizeof(float) + 256));
    if (!a) return nullptr;
    if (extra) *extra = a + ((n + 63) / 64) * 64;
    if (x) {
        subsample_kernel<<<(unsigned)std::min<size_t>((n + 255) / 256, 148 * 16), 256, 0, ctx->stream>>>(x, a, B * Cin, H, W, OH, OW, s);
        ++ctx->launches;
    }
    return a;
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
    if (ctx->conv_algo == CNN_CONV_AUTO && conv_s1_first_supported(ctx, Cin, H, W, Cout, k, stride))
        return conv_s1_first_fwd(ctx, x, w, bias, y, nullptr, nullptr, B, H, W, Cout);
    if (use_tc(ctx, Cin, Cout, k, stride)) return conv_fwd_tc(ctx, x, w, bias, y, B, Cin, H, W, Cout, k, stride);
    if (use_sub1x1(ctx, Cin, Cout, k, stride)) {
        const int OH = (H - 1) / stride + 1, OW = (W - 1) / stride + 1;
        float* xs = subsampled(ctx, x, B, Cin, H, W, stride, 0, nullptr);
        CNN_REQUIRE(xs, "conv 1x1: arena allocation failed");
        return conv_fwd_tc(ctx, xs, w, bias, y, B, Cin, OH, OW, Cout, 1, 1);
    }
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
    if (ctx->conv_algo == CNN_CONV_AUTO && conv_s1_first_supported(ctx, Cin, H, W, Cout, k, stride) && !getenv("CNN_DBG_NOS1FIRSTWG"))
        return conv_s1_first_wgrad(ctx, x, delta, dw, db, B, H, W, Cout, scale);
    if (use_tc(ctx, Cin, Cout, k, stride, true))
        return conv_wgrad_tc(ctx, x, delta, dw, db, B, Cin, H, W, Cout, k, stride, scale);
    if (use_sub1x1(ctx, Cin, Cout, k, stride)) {
        const int OH = (H - 1) / stride + 1, OW = (W - 1) / stride + 1;
        float* xs = subsampled(ctx, x, B, Cin, H, W, stride, 0, nullptr);
        CNN_REQUIRE(xs, "conv 1x1: arena allocation failed");
        return conv_wgrad_tc(ctx, xs, delta, dw, db, B, Cin, OH, OW, Cout, 1, 1, scale);
    }
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
    if (use_sub1x1(ctx, Cin, Cout, k, stride)) {
        const int OH = (H - 1) / stride + 1, OW = (W - 1) / stride + 1;
        float* dxs = subsampled(ctx, nullptr, B, Cin, H, W, stride, 0, nullptr);
        CNN_REQUIRE(dxs, "conv 1x1: arena allocation failed");
        if (int rc = conv_dgrad_tc(ctx, w, delta, dxs, B, Cin, OH, OW, Cout, 1, 1)) return rc;
        const size_t n = (size_t)B * Cin * H * W;
        CNN_LAUNCH(ctx, upsample_scatter_kernel, (unsigned)std::min<size_t>((n + 255) / 256, 148 * 16), 256, 0, dxs, dx, B * Cin, H, W, OH,
                   OW, stride);
        return CNN_OK;
    }
    if (ctx->conv_algo == CNN_CONV_TCGEN05) {
        cnn_set_error("cnn_conv2d_backward_data: shape not supported by the tcgen05 path");
        return CNN_ERR_UNSUPPORTED;
    }
    return conv_dgrad_simt(ctx, w, delta, dx, B, Cin, H, W, Cout, k, stride);
}

}  // extern "C"
