// layers.cpp -- Conv2D / MaxPool2D / ReLU / LinearLayer / BatchNorm2D with the reference's
// signatures (cpu/include/architectures.h:49-173); every method body is a thin call into the
// C ABI of libcnn_b200.  Semantics kept from the reference (file:line in cpu/src):
//   * layers own persistent output buffers sized by the FIRST batch (conv2d.cpp:46-52); a later
//     smaller batch fills a prefix and the returned vector keeps its original length (§3.2 quirk)
//   * forward remembers its input only when !no_grad (conv2d.cpp:62)
//   * ReLU / BatchNorm backward work in place on the caller's delta and return it (relu.cpp:30-44)
//   * LinearLayer allocates fresh output tensors every forward (linear.cpp:27-29)
//   * weights are drawn on the host with the reference's seeds (conv2d.cpp:22-30, linear.cpp:14-18)
#include <cassert>
#include <cstring>

#include "architectures.h"
#include "backend.h"

using namespace architectures;
using cnn_b200::batch_on_device;
using cnn_b200::check;
using cnn_b200::ctx;
using cnn_b200::DeviceBuffer;
using cnn_b200::make_slab;
using cnn_b200::make_views;
using cnn_b200::Slab;

data_type architectures::random_times = 10.f;
bool architectures::no_grad = false;

std::vector<tensor> Layer::get_output() const {
    if (!output.empty()) output[0]->sync_host();  // one D2H refreshes the whole slab
    for (const auto& t : output) t->sync_host();
    return output;
}

namespace {
int conv_out(int h, int k, int s) { return (h - k) / s + 1; }
}

// ================================================================================= Conv2D
Conv2D::Conv2D(std::string _name, const int _in_channels, const int _out_channels, const int _kernel_size,
               const int _stride)
    : Layer(_name), in_channels(_in_channels), out_channels(_out_channels), kernel_size(_kernel_size),
      stride(_stride), params_for_one_kernel(_in_channels * _kernel_size * _kernel_size) {
    assert(_kernel_size & 1 && _kernel_size >= 3 && "kernel size must be odd and >= 3");
    assert(_in_channels > 0 && _out_channels > 0 && _stride > 0);
    // Same draws as conv2d.cpp:22-30: every bias first, then the filters, N(0,1) / random_times.
    const size_t nw = (size_t)out_channels * params_for_one_kernel;
    host_params.resize(nw + out_channels);
    seed.seed(212);
    std::normal_distribution<float> engine(0.0, 1.0);
    for (int o = 0; o < out_channels; ++o) host_params[nw + o] = engine(seed) / random_times;
    for (size_t i = 0; i < nw; ++i) host_params[i] = engine(seed) / random_times;
}

void Conv2D::upload_params() {
    if (!params) params = std::make_shared<DeviceBuffer>(host_params.size() * sizeof(float));
    if (device_params_stale) {
        check(cnn_h2d(ctx(), params->dev, host_params.data(), host_params.size() * sizeof(float)), "h2d(conv params)");
        check(cnn_sync(ctx()), "sync");
        device_params_stale = false;
    }
}

void Conv2D::download_params() const {
    if (host_params_stale && params) {
        check(cnn_d2h(ctx(), host_params.data(), params->dev, host_params.size() * sizeof(float)), "d2h(conv params)");
        host_params_stale = false;
    }
}

std::vector<tensor> Conv2D::forward(const std::vector<tensor>& input) {
    const int B = (int)input.size(), H = input[0]->H, W = input[0]->W;
    const int OH = conv_out(H, kernel_size, stride), OW = conv_out(W, kernel_size, stride);
    if (output.empty()) {
        out_slab = make_slab(B, out_channels, OH, OW);
        output = make_views(out_slab, name + "_output_");
    }
    upload_params();
    std::shared_ptr<Slab> used;
    static thread_local std::shared_ptr<Slab> unused;
    std::shared_ptr<Slab>& stage = no_grad ? unused : in_slab;   // keep the upload alive for backward
    const float* x = batch_on_device(input, stage, used);
    if (!no_grad) { in_slab = used; in_H = H; in_W = W; in_B = B; }
    const size_t nw = (size_t)out_channels * params_for_one_kernel;
    check(cnn_conv2d_forward(ctx(), x, params->f(), params->f() + nw, out_slab->dev, B, in_channels, H, W,
                             out_channels, kernel_size, stride), "cnn_conv2d_forward");
    out_slab->device_written();
    return output;
}

std::vector<tensor> Conv2D::backward(std::vector<tensor>& delta) {
    const int B = (int)delta.size();
    assert(in_slab && "Conv2D::backward needs a forward with gradients enabled");
    if (!grads) grads = std::make_shared<DeviceBuffer>(host_params.size() * sizeof(float));
    if (delta_output.empty()) {
        dx_slab = make_slab(B, in_channels, in_H, in_W);
        delta_output = make_views(dx_slab, name + "_delta_");
    }
    std::shared_ptr<Slab> stage, used;
    const float* d = batch_on_device(delta, stage, used);
    const size_t nw = (size_t)out_channels * params_for_one_kernel;
    check(cnn_conv2d_backward_weights(ctx(), in_slab->dev, d, grads->f(), grads->f() + nw, B, in_channels, in_H,
                                      in_W, out_channels, kernel_size, stride, 1.f / (float)B),
          "cnn_conv2d_backward_weights");
    check(cnn_conv2d_backward_data(ctx(), params->f(), d, dx_slab->dev, B, in_channels, in_H, in_W, out_channels,
                                   kernel_size, stride), "cnn_conv2d_backward_data");
    dx_slab->device_written();
    if (stage) check(cnn_sync(ctx()), "sync");  // a temporary upload must outlive the kernels reading it
    return delta_output;
}

void Conv2D::update_gradients(const data_type learning_rate) {
    assert(grads && "update_gradients before backward");
    check(cnn_sgd_step(ctx(), params->f(), grads->f(), host_params.size(), learning_rate), "cnn_sgd_step");
    host_params_stale = true;
}

void Conv2D::save_weights(std::ofstream& writer) const {  // filters then biases, conv2d.cpp:220-226
    download_params();
    writer.write(reinterpret_cast<const char*>(host_params.data()),
                 static_cast<std::streamsize>(sizeof(data_type) * host_params.size()));
}

void Conv2D::load_weights(std::ifstream& reader) {
    reader.read(reinterpret_cast<char*>(host_params.data()),
                static_cast<std::streamsize>(sizeof(data_type) * host_params.size()));
    host_params_stale = false;
    device_params_stale = true;
}

int Conv2D::get_params_num() const { return (params_for_one_kernel + 1) * out_channels; }

// ================================================================================ MaxPool2D
std::vector<tensor> MaxPool2D::forward(const std::vector<tensor>& input) {
    const int B = (int)input.size(), C = input[0]->C, H = input[0]->H, W = input[0]->W;
    const int OH = conv_out(H, kernel_size, step), OW = conv_out(W, kernel_size, step);
    if (output.empty()) {
        out_slab = make_slab(B, C, OH, OW);
        output = make_views(out_slab, name + "_output_");
        if (!no_grad) {  // mask / delta buffers only if the FIRST forward ran with gradients (pool2d.cpp:23)
            mask = std::make_shared<DeviceBuffer>(sizeof(int32_t) * out_slab->count());
            dx_slab = make_slab(B, C, H, W);
            delta_output = make_views(dx_slab, name + "_delta_");
        }
    }
    in_C = C; in_H = H; in_W = W;
    std::shared_ptr<Slab> stage, used;
    const float* x = batch_on_device(input, stage, used);
    check(cnn_maxpool_forward(ctx(), x, out_slab->dev, (!no_grad && mask) ? mask->i() : nullptr, B, C, H, W,
                              kernel_size, step), "cnn_maxpool_forward");
    out_slab->device_written();
    if (stage) check(cnn_sync(ctx()), "sync");
    return output;
}

std::vector<tensor> MaxPool2D::backward(std::vector<tensor>& delta) {
    assert(mask && "MaxPool2D::backward needs a first forward with gradients enabled");
    const int B = (int)delta.size();
    std::shared_ptr<Slab> stage, used;
    const float* d = batch_on_device(delta, stage, used);
    check(cnn_maxpool_backward(ctx(), d, mask->i(), dx_slab->dev, B, in_C, in_H, in_W, kernel_size, step),
          "cnn_maxpool_backward");
    dx_slab->device_written();
    if (stage) check(cnn_sync(ctx()), "sync");
    return delta_output;
}

// ===================================================================================== ReLU
std::vector<tensor> ReLU::forward(const std::vector<tensor>& input) {
    const int B = (int)input.size();
    if (output.empty()) {
        out_slab = make_slab(B, input[0]->C, input[0]->H, input[0]->W);
        output = make_views(out_slab, name + "_output_");
    }
    std::shared_ptr<Slab> stage, used;
    const float* x = batch_on_device(input, stage, used);
    check(cnn_relu_forward(ctx(), x, out_slab->dev, out_slab->per * (size_t)B), "cnn_relu_forward");
    out_slab->device_written();
    if (stage) check(cnn_sync(ctx()), "sync");
    return output;
}

std::vector<tensor> ReLU::backward(std::vector<tensor>& delta) {
    const int B = (int)delta.size();
    std::shared_ptr<Slab> stage, used;
    float* d = batch_on_device(delta, stage, used);
    check(cnn_relu_backward(ctx(), d, out_slab->dev, out_slab->per * (size_t)B), "cnn_relu_backward");
    used->device_written();
    if (stage) {  // delta was plain host memory: the in-place contract means writing it back
        used->to_host();
        for (int b = 0; b < B; ++b) std::memcpy(delta[b]->data, used->host + (size_t)b * used->per, used->per * sizeof(float));
    }
    for (int b = 0; b < B; ++b) delta[b]->name = name + "_delta_" + std::to_string(b);
    return delta;
}

// ============================================================================== LinearLayer
LinearLayer::LinearLayer(std::string _name, const int _in_channels, const int _out_channels)
    : Layer(_name), in_channels(_in_channels), out_channels(_out_channels) {
    // linear.cpp:14-18: seed 1998, biases first, then W[in][out]
    const size_t nw = (size_t)in_channels * out_channels;
    host_params.resize(nw + out_channels);
    std::default_random_engine e(1998);
    std::normal_distribution<float> engine(0.0, 1.0);
    for (int i = 0; i < out_channels; ++i) host_params[nw + i] = engine(e) / random_times;
    for (size_t i = 0; i < nw; ++i) host_params[i] = engine(e) / random_times;
}

void LinearLayer::upload_params() {
    if (!params) params = std::make_shared<DeviceBuffer>(host_params.size() * sizeof(float));
    if (device_params_stale) {
        check(cnn_h2d(ctx(), params->dev, host_params.data(), host_params.size() * sizeof(float)), "h2d(linear params)");
        check(cnn_sync(ctx()), "sync");
        device_params_stale = false;
    }
}

void LinearLayer::download_params() const {
    if (host_params_stale && params) {
        check(cnn_d2h(ctx(), host_params.data(), params->dev, host_params.size() * sizeof(float)), "d2h(linear params)");
        host_params_stale = false;
    }
}

std::vector<tensor> LinearLayer::forward(const std::vector<tensor>& input) {
    const int B = (int)input.size();
    delta_shape = input[0]->get_shape();
    out_slab = make_slab(B, out_channels, 1, 1);           // fresh tensors every call (linear.cpp:27-29)
    output = make_views(out_slab, name + "_output_");
    upload_params();
    std::shared_ptr<Slab> used;
    static thread_local std::shared_ptr<Slab> unused;
    std::shared_ptr<Slab>& stage = no_grad ? unused : in_slab;
    const float* x = batch_on_device(input, stage, used);
    if (!no_grad) in_slab = used;
    const size_t nw = (size_t)in_channels * out_channels;
    check(cnn_linear_forward(ctx(), x, params->f(), params->f() + nw, out_slab->dev, B, in_channels, out_channels),
          "cnn_linear_forward");
    out_slab->device_written();
    return output;
}

std::vector<tensor> LinearLayer::backward(std::vector<tensor>& delta) {
    const int B = (int)delta.size();
    assert(in_slab && "LinearLayer::backward needs a forward with gradients enabled");
    if (!grads) grads = std::make_shared<DeviceBuffer>(host_params.size() * sizeof(float));
    if (delta_output.empty()) {
        dx_slab = make_slab(B, std::get<0>(delta_shape), std::get<1>(delta_shape), std::get<2>(delta_shape));
        delta_output = make_views(dx_slab, "linear_delta_");
    }
    std::shared_ptr<Slab> stage, used;
    const float* d = batch_on_device(delta, stage, used);
    const size_t nw = (size_t)in_channels * out_channels;
    check(cnn_linear_backward(ctx(), in_slab->dev, params->f(), d, grads->f(), grads->f() + nw, dx_slab->dev, B,
                              in_channels, out_channels, 1.f / (float)B), "cnn_linear_backward");
    dx_slab->device_written();
    if (stage) check(cnn_sync(ctx()), "sync");
    return delta_output;
}

void LinearLayer::update_gradients(const data_type learning_rate) {
    assert(grads && "update_gradients before backward");
    check(cnn_sgd_step(ctx(), params->f(), grads->f(), host_params.size(), learning_rate), "cnn_sgd_step");
    host_params_stale = true;
}

void LinearLayer::save_weights(std::ofstream& writer) const {  // W then bias, linear.cpp:105-108
    download_params();
    writer.write(reinterpret_cast<const char*>(host_params.data()),
                 static_cast<std::streamsize>(sizeof(data_type) * host_params.size()));
}

void LinearLayer::load_weights(std::ifstream& reader) {
    reader.read(reinterpret_cast<char*>(host_params.data()),
                static_cast<std::streamsize>(sizeof(data_type) * host_params.size()));
    host_params_stale = false;
    device_params_stale = true;
}

// ============================================================================== BatchNorm2D
BatchNorm2D::BatchNorm2D(std::string _name, const int _out_channels, const data_type _eps, const data_type _momentum)
    : Layer(_name), out_channels(_out_channels), eps(_eps), momentum(_momentum),
      host_params((size_t)4 * _out_channels, 0.f) {
    for (int o = 0; o < out_channels; ++o) host_params[o] = 1.f;  // gamma 1, beta 0, moving stats 0 (:17-21)
}

void BatchNorm2D::upload_params() {
    if (!params) params = std::make_shared<DeviceBuffer>(host_params.size() * sizeof(float));
    if (device_params_stale) {
        check(cnn_h2d(ctx(), params->dev, host_params.data(), host_params.size() * sizeof(float)), "h2d(bn params)");
        check(cnn_sync(ctx()), "sync");
        device_params_stale = false;
    }
}

void BatchNorm2D::download_params() const {
    if (host_params_stale && params) {
        check(cnn_d2h(ctx(), host_params.data(), params->dev, host_params.size() * sizeof(float)), "d2h(bn params)");
        host_params_stale = false;
    }
}

std::vector<tensor> BatchNorm2D::forward(const std::vector<tensor>& input) {
    const int B = (int)input.size(), H = input[0]->H, W = input[0]->W, C = out_channels;
    if (output.empty()) {
        out_slab = make_slab(B, C, H, W);
        xhat_slab = make_slab(B, C, H, W);
        output = make_views(out_slab, name + "_output_");
        batch_stats = std::make_shared<DeviceBuffer>(sizeof(float) * 2 * C);
    }
    upload_params();
    std::shared_ptr<Slab> used;
    static thread_local std::shared_ptr<Slab> unused;
    std::shared_ptr<Slab>& stage = no_grad ? unused : in_slab;
    const float* x = batch_on_device(input, stage, used);
    float* p = params->f();
    if (!no_grad) {
        in_slab = used;
        check(cnn_bn_forward_train(ctx(), x, p, p + C, p + 2 * C, p + 3 * C, batch_stats->f(), batch_stats->f() + C,
                                   xhat_slab->dev, out_slab->dev, B, C, H, W, eps, momentum), "cnn_bn_forward_train");
        host_params_stale = true;  // moving statistics changed on the device
    } else {
        check(cnn_bn_forward_eval(ctx(), x, p, p + C, p + 2 * C, p + 3 * C, xhat_slab->dev, out_slab->dev, B, C, H, W,
                                  eps), "cnn_bn_forward_eval");
    }
    out_slab->device_written();
    return output;
}

std::vector<tensor> BatchNorm2D::backward(std::vector<tensor>& delta) {
    const int B = (int)delta.size(), C = out_channels;
    assert(in_slab && "BatchNorm2D::backward needs a forward with gradients enabled");
    if (!grads) {
        grads = std::make_shared<DeviceBuffer>(host_params.size() * sizeof(float));
        check(cnn_memset(ctx(), grads->dev, 0, grads->bytes), "memset(bn grads)");  // moving-stat slots stay 0
    }
    std::shared_ptr<Slab> stage, used;
    float* d = batch_on_device(delta, stage, used);
    check(cnn_bn_backward(ctx(), d, in_slab->dev, xhat_slab->dev, params->f(), batch_stats->f(), batch_stats->f() + C,
                          grads->f(), grads->f() + C, B, C, in_slab->H, in_slab->W, eps), "cnn_bn_backward");
    used->device_written();
    if (stage) {
        used->to_host();
        for (int b = 0; b < B; ++b) std::memcpy(delta[b]->data, used->host + (size_t)b * used->per, used->per * sizeof(float));
    }
    return delta;
}

void BatchNorm2D::update_gradients(const data_type learning_rate) {
    assert(grads && "update_gradients before backward");
    check(cnn_sgd_step(ctx(), params->f(), grads->f(), (size_t)2 * out_channels, learning_rate), "cnn_sgd_step");
    host_params_stale = true;
}

void BatchNorm2D::save_weights(std::ofstream& writer) const {  // gamma, beta, moving_mean, moving_var (:168-174)
    download_params();
    writer.write(reinterpret_cast<const char*>(host_params.data()),
                 static_cast<std::streamsize>(sizeof(data_type) * host_params.size()));
}

void BatchNorm2D::load_weights(std::ifstream& reader) {
    reader.read(reinterpret_cast<char*>(host_params.data()),
                static_cast<std::streamsize>(sizeof(data_type) * host_params.size()));
    host_params_stale = false;
    device_params_stale = true;
}
