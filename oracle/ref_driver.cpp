// ref_driver.cpp -- C entry points around the UNMODIFIED reference layer classes.
//
// TEST INFRASTRUCTURE ONLY.  Compiled (oracle/Makefile) together with the
// reference's own sources, in place from /root/reference/cpu/src, into
// oracle/_ref/libcnn_ref.so.  Nothing from the reference is copied into this
// repo; this file only *calls* its public API (Conv2D::forward/backward, ...,
// AlexNet, softmax, cross_entroy_backward) the way cnn.cpp:81-92 does.
// `private` is widened for this translation unit only so gradients held in
// private members (Conv2D::weights_gradients ...) can be read back.
// std / stub headers first (their include guards keep the widening off them)
#include <opencv2/core.hpp>
#include <cstdio>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <list>
#include <map>
#include <random>

#define private public
#include "architectures.h"
#include "func.h"
#undef private

using namespace architectures;

namespace {

std::vector<tensor> make_batch(const float* src, int B, int C, int H, int W) {
    std::vector<tensor> v;
    v.reserve(B);
    const size_t n = size_t(C) * H * W;
    for (int b = 0; b < B; ++b) {
        tensor t(new Tensor3D(C, H, W));
        if (src) std::memcpy(t->data, src + b * n, n * sizeof(float));
        v.emplace_back(std::move(t));
    }
    return v;
}

void read_batch(const std::vector<tensor>& v, float* dst, int B) {
    if (!dst) return;
    for (int b = 0; b < B; ++b) {
        const size_t n = size_t(v[b]->get_length());
        std::memcpy(dst + b * n, v[b]->data, n * sizeof(float));
    }
}

struct Quiet {  // the reference prints on save/load; keep test logs clean
    std::streambuf* old;
    Quiet() : old(std::cout.rdbuf(nullptr)) {}
    ~Quiet() { std::cout.rdbuf(old); }
};

struct Net {
    std::unique_ptr<AlexNet> alex;                 // the reference container, or
    std::list<std::shared_ptr<Layer>> layers;      // a list built from its layer classes
    std::list<std::shared_ptr<Layer>>& seq() { return alex ? alex->layers_sequence : layers; }
    std::vector<tensor> probs, delta;
};

}  // namespace

extern "C" {

// ---- single layers: forward (+ backward when delta != nullptr) ----------------

int ref_conv2d(int B, int Cin, int H, int W, int Cout, int k, int stride, const float* x,
               const float* w, const float* bias, const float* delta, float* y, float* dw,
               float* db, float* dx) {
    Conv2D conv("conv", Cin, Cout, k, stride);
    const int per = Cin * k * k;
    for (int o = 0; o < Cout; ++o) std::memcpy(conv.weights[o]->data, w + o * per, per * sizeof(float));
    std::memcpy(conv.bias.data(), bias, Cout * sizeof(float));
    auto in = make_batch(x, B, Cin, H, W);
    auto out = conv.forward(in);
    read_batch(out, y, B);
    if (!delta) return 0;
    auto d = make_batch(delta, B, Cout, out[0]->H, out[0]->W);
    auto back = conv.backward(d);
    read_batch(back, dx, B);
    for (int o = 0; o < Cout; ++o)
        std::memcpy(dw + o * per, conv.weights_gradients[o]->data, per * sizeof(float));
    std::memcpy(db, conv.bias_gradients.data(), Cout * sizeof(float));
    return 0;
}

int ref_maxpool(int B, int C, int H, int W, int k, int step, const float* x, const float* delta,
                float* y, int* mask, float* dx) {
    MaxPool2D pool("pool", k, step);
    auto in = make_batch(x, B, C, H, W);
    auto out = pool.forward(in);
    read_batch(out, y, B);
    const size_t n = size_t(out[0]->get_length());
    if (mask)
        for (int b = 0; b < B; ++b) std::memcpy(mask + b * n, pool.mask[b].data(), n * sizeof(int));
    if (!delta) return 0;
    auto d = make_batch(delta, B, C, out[0]->H, out[0]->W);
    auto back = pool.backward(d);
    read_batch(back, dx, B);
    return 0;
}

int ref_relu(int B, int C, int H, int W, const float* x, float* delta_inout, float* y) {
    ReLU relu("relu");
    auto in = make_batch(x, B, C, H, W);
    auto out = relu.forward(in);
    read_batch(out, y, B);
    if (!delta_inout) return 0;
    auto d = make_batch(delta_inout, B, C, H, W);
    auto back = relu.backward(d);
    read_batch(back, delta_inout, B);
    return 0;
}

int ref_linear(int B, int C, int H, int W, int out_ch, const float* x, const float* w,
               const float* bias, const float* delta, float* y, float* dw, float* db, float* dx) {
    const int in_ch = C * H * W;
    LinearLayer lin("linear", in_ch, out_ch);
    std::memcpy(lin.weights.data(), w, size_t(in_ch) * out_ch * sizeof(float));
    std::memcpy(lin.bias.data(), bias, out_ch * sizeof(float));
    auto in = make_batch(x, B, C, H, W);
    auto out = lin.forward(in);
    read_batch(out, y, B);
    if (!delta) return 0;
    auto d = make_batch(delta, B, out_ch, 1, 1);
    auto back = lin.backward(d);
    read_batch(back, dx, B);
    std::memcpy(dw, lin.weights_gradients.data(), size_t(in_ch) * out_ch * sizeof(float));
    std::memcpy(db, lin.bias_gradients.data(), out_ch * sizeof(float));
    return 0;
}

// eval != 0 runs the forward under WithoutGrad (moving statistics); no backward then.
int ref_batchnorm(int B, int C, int H, int W, int eval, const float* x, const float* gamma,
                  const float* beta, float* moving_mean, float* moving_var, float* delta_inout,
                  float* y, float* xhat, float* mean, float* var, float* dgamma, float* dbeta) {
    BatchNorm2D bn("bn", C);
    std::memcpy(bn.gamma.data(), gamma, C * sizeof(float));
    std::memcpy(bn.beta.data(), beta, C * sizeof(float));
    std::memcpy(bn.moving_mean.data(), moving_mean, C * sizeof(float));
    std::memcpy(bn.moving_var.data(), moving_var, C * sizeof(float));
    auto in = make_batch(x, B, C, H, W);
    std::vector<tensor> out;
    if (eval) {
        WithoutGrad guard;
        out = bn.forward(in);
    } else {
        out = bn.forward(in);
    }
    read_batch(out, y, B);
    read_batch(bn.normed_input, xhat, B);
    std::memcpy(moving_mean, bn.moving_mean.data(), C * sizeof(float));
    std::memcpy(moving_var, bn.moving_var.data(), C * sizeof(float));
    if (mean) std::memcpy(mean, bn.buffer_mean.data(), C * sizeof(float));
    if (var) std::memcpy(var, bn.buffer_var.data(), C * sizeof(float));
    if (eval || !delta_inout) return 0;
    auto d = make_batch(delta_inout, B, C, H, W);
    auto back = bn.backward(d);
    read_batch(back, delta_inout, B);
    std::memcpy(dgamma, bn.gamma_gradients.data(), C * sizeof(float));
    std::memcpy(dbeta, bn.beta_gradients.data(), C * sizeof(float));
    return 0;
}

// softmax + one_hot + cross_entroy_backward (func.cpp); argmax per row into pred.
int ref_softmax_xent(int B, int n, const float* logits, const int* labels, float* probs,
                     float* delta, float* loss, int* pred) {
    auto in = make_batch(logits, B, n, 1, 1);
    auto p = softmax(in);
    read_batch(p, probs, B);
    if (pred)
        for (int b = 0; b < B; ++b) pred[b] = p[b]->argmax();
    if (!labels) return 0;
    std::vector<int> lab(labels, labels + B);
    auto ld = cross_entroy_backward(p, one_hot(lab, n));
    *loss = ld.first;
    read_batch(ld.second, delta, B);
    return 0;
}

// ---- networks ----------------------------------------------------------------

void* ref_alexnet_create(int num_classes, int batch_norm) {
    Net* n = new Net;
    n->alex.reset(new AlexNet(num_classes, batch_norm != 0));
    return n;
}

// specs: n x 5 ints {type, a, b, c, d} with the oracle's layer codes
// (0 conv cin,cout,k,stride; 1 bn C; 2 relu; 3 pool k,step; 4 linear in,out).
void* ref_net_create(const int* specs, int n_layers) {
    Net* n = new Net;
    for (int i = 0; i < n_layers; ++i) {
        const int* s = specs + 5 * i;
        const std::string nm = "layer_" + std::to_string(i);
        switch (s[0]) {
            case 0: n->layers.emplace_back(new Conv2D(nm, s[1], s[2], s[3], s[4])); break;
            case 1: n->layers.emplace_back(new BatchNorm2D(nm, s[1])); break;
            case 2: n->layers.emplace_back(new ReLU(nm)); break;
            case 3: n->layers.emplace_back(new MaxPool2D(nm, s[1], s[2])); break;
            case 4: n->layers.emplace_back(new LinearLayer(nm, s[1], s[2])); break;
            default: delete n; return nullptr;
        }
    }
    return n;
}

void ref_net_destroy(void* h) { delete static_cast<Net*>(h); }

// Raw little-endian fp32 stream in layer order == the reference checkpoint format.
int ref_net_load_file(void* h, const char* path) {
    Net* n = static_cast<Net*>(h);
    std::ifstream reader(path, std::ios::binary);
    if (!reader) return -1;
    for (auto& l : n->seq()) l->load_weights(reader);
    return 0;
}
int ref_net_save_file(void* h, const char* path) {
    Net* n = static_cast<Net*>(h);
    std::ofstream writer(path, std::ios::binary);
    if (!writer) return -1;
    for (auto& l : n->seq()) l->save_weights(writer);
    return 0;
}

// Gradients in checkpoint order (BN moving statistics -> zeros).
long ref_net_get_grads(void* h, float* out) {
    Net* n = static_cast<Net*>(h);
    long pos = 0;
    auto put = [&](const float* p, long cnt) {
        if (out) std::memcpy(out + pos, p, cnt * sizeof(float));
        pos += cnt;
    };
    for (auto& l : n->seq()) {
        if (auto* c = dynamic_cast<Conv2D*>(l.get())) {
            for (auto& g : c->weights_gradients) put(g->data, g->get_length());
            put(c->bias_gradients.data(), long(c->bias_gradients.size()));
        } else if (auto* f = dynamic_cast<LinearLayer*>(l.get())) {
            put(f->weights_gradients.data(), long(f->weights_gradients.size()));
            put(f->bias_gradients.data(), long(f->bias_gradients.size()));
        } else if (auto* b = dynamic_cast<BatchNorm2D*>(l.get())) {
            put(b->gamma_gradients.data(), b->out_channels);
            put(b->beta_gradients.data(), b->out_channels);
            std::vector<float> z(2 * b->out_channels, 0.f);
            put(z.data(), long(z.size()));
        }
    }
    return pos;
}

// AlexNet::forward (alexnet.cpp:35-46) or the same loop over a custom list.
int ref_net_forward(void* h, int B, int C, int H, int W, const float* x, float* logits,
                    int no_grad_flag) {
    Net* n = static_cast<Net*>(h);
    auto in = make_batch(x, B, C, H, W);
    std::vector<tensor> out(in);
    auto run = [&]() {
        if (n->alex) out = n->alex->forward(in);
        else for (auto& l : n->layers) out = l->forward(out);
    };
    if (no_grad_flag) { WithoutGrad guard; run(); } else run();
    read_batch(out, logits, B);
    return int(out[0]->get_length());
}

// grad_cam.cpp:71-80 for ONE image: forward with gradients enabled, softmax, AlexNet::grad_cam(layer) ->
// the HxW class-activation map as the reference's 8-bit cv::Mat (alexnet.cpp:95-142).
int ref_alexnet_grad_cam(void* h, const float* x, const char* layer, unsigned char* cam, int cam_cap, float* probs) {
    Net* n = static_cast<Net*>(h);
    if (!n->alex) return -1;
    auto in = make_batch(x, 1, 3, 224, 224);
    const auto out = n->alex->forward(in);
    const auto p = softmax(out);
    if (probs) std::memcpy(probs, p[0]->data, sizeof(float) * p[0]->get_length());
    const cv::Mat m = n->alex->grad_cam(layer);
    const int cnt = m.rows * m.cols;
    if (cnt > cam_cap) return -2;
    std::memcpy(cam, m.data, cnt);
    return cnt;
}

// The body of the train loop, cnn.cpp:81-92.
float ref_net_train_step(void* h, int B, int C, int H, int W, const float* x, const int* labels,
                         float lr, float* probs, float* dx_image) {
    Net* n = static_cast<Net*>(h);
    auto in = make_batch(x, B, C, H, W);
    std::vector<tensor> out(in);
    if (n->alex) out = n->alex->forward(in);
    else for (auto& l : n->layers) out = l->forward(out);
    const int classes = out[0]->get_length();
    auto p = softmax(out);
    std::vector<int> lab(labels, labels + B);
    auto ld = cross_entroy_backward(p, one_hot(lab, classes));
    if (n->alex) n->alex->backward(ld.second);
    else for (auto l = n->layers.rbegin(); l != n->layers.rend(); ++l) ld.second = (*l)->backward(ld.second);
    if (n->alex) n->alex->update_gradients(lr);
    else for (auto& l : n->layers) l->update_gradients(lr);
    read_batch(p, probs, B);
    read_batch(ld.second, dx_image, B);
    return ld.first;
}

// Output of layer `idx` after the last forward (Layer::get_output, architectures.h:45).
long ref_net_layer_output(void* h, int idx, int B, float* out) {
    Net* n = static_cast<Net*>(h);
    auto it = n->seq().begin();
    std::advance(it, idx);
    auto v = (*it)->get_output();
    if (out) read_batch(v, out, B);
    return long(B) * v[0]->get_length();
}

// Initial parameters exactly as the reference constructors draw them
// (conv2d.cpp:22-30 seed 212, linear.cpp:14-18 seed 1998): serialise a fresh net.
int ref_alexnet_init_params_to_file(int num_classes, int batch_norm, const char* path) {
    AlexNet net(num_classes, batch_norm != 0);
    Quiet q;
    net.save_weights(path);
    return 0;
}

}  // extern "C"
