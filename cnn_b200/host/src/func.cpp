// func.cpp -- softmax / one_hot / cross_entroy_backward with the reference's signatures
// (cpu/include/func.h:8-18), computed by cnn_softmax_xent / cnn_xent_backward on the device.
// The tiny [B][classes] results are copied back immediately: cnn.cpp:92 and inference.cpp:69-70
// read probs[b]->data directly.
#include <cassert>
#include <sstream>

#include "backend.h"
#include "func.h"

using cnn_b200::batch_on_device;
using cnn_b200::check;
using cnn_b200::ctx;
using cnn_b200::make_slab;
using cnn_b200::make_views;
using cnn_b200::Slab;

std::vector<tensor> softmax(const std::vector<tensor>& input) {
    const int B = (int)input.size(), n = input[0]->get_length();
    std::shared_ptr<Slab> stage, used;
    const float* z = batch_on_device(input, stage, used);
    auto out = make_slab(B, n, 1, 1);
    check(cnn_softmax_xent(ctx(), z, nullptr, out->dev, nullptr, nullptr, nullptr, B, n), "cnn_softmax_xent");
    out->device_written();
    out->to_host();
    return make_views(out, "softmax_");
}

std::vector<tensor> one_hot(const std::vector<int>& labels, const int num_classes) {
    const int B = (int)labels.size();
    auto slab = make_slab(B, num_classes, 1, 1);
    for (int b = 0; b < B; ++b) {
        assert(labels[b] >= 0 && labels[b] < num_classes);
        for (int i = 0; i < num_classes; ++i) slab->host[(size_t)b * num_classes + i] = (i == labels[b]) ? 1.0f : 0.f;
    }
    return make_views(slab, "one_hot_");
}

std::pair<data_type, std::vector<tensor> > cross_entroy_backward(const std::vector<tensor>& probs,
                                                                  const std::vector<tensor>& labels) {
    const int B = (int)labels.size(), n = probs[0]->get_length();
    std::shared_ptr<Slab> sp, up, sl, ul;
    const float* p = batch_on_device(probs, sp, up);
    const float* y = batch_on_device(labels, sl, ul);
    auto delta = make_slab(B, n, 1, 1);
    cnn_b200::DeviceBuffer loss_dev(sizeof(float));
    check(cnn_xent_backward(ctx(), p, y, delta->dev, loss_dev.f(), B, n), "cnn_xent_backward");
    delta->device_written();
    float loss_sum = 0.f;
    check(cnn_d2h(ctx(), &loss_sum, loss_dev.dev, sizeof(float)), "d2h(loss)");
    delta->to_host();
    data_type loss_value = loss_sum;
    loss_value = loss_value * (-1.0) / B;  // func.cpp:71, evaluated in double like the reference
    return std::make_pair(loss_value, make_views(delta, "xent_delta_"));
}

std::string float_to_string(const float value, const int precision) {
    std::stringstream buffer;
    buffer.precision(precision);
    buffer.setf(std::ios::fixed);
    buffer << value;
    return buffer.str();
}
