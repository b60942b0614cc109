// backend.cpp -- context singleton, device buffers, slab pairs (replaces the reference's
// per-image `new data_type[C*H*W]`, data_format.h:17-26 / data_format.cpp:152-158).
#include "backend.h"

#include <cstring>

namespace cnn_b200 {

cnn_ctx* ctx() {
    static cnn_ctx* c = [] {
        cnn_ctx* h = nullptr;
        const char* dev = std::getenv("CNN_B200_DEVICE");
        check(cnn_ctx_create(dev ? std::atoi(dev) : 0, nullptr, &h), "cnn_ctx_create");
        return h;
    }();
    return c;
}

DeviceBuffer::DeviceBuffer(size_t n) : bytes(n) { check(cnn_malloc(ctx(), n, &dev), "cnn_malloc"); }
DeviceBuffer::~DeviceBuffer() { cnn_free(ctx(), dev); }

Slab::Slab(int B_, int C_, int H_, int W_) : B(B_), C(C_), H(H_), W(W_), per((size_t)C_ * H_ * W_) {
    void* d = nullptr;
    void* h = nullptr;
    check(cnn_malloc(ctx(), count() * sizeof(float), &d), "cnn_malloc(slab)");
    check(cnn_host_alloc(ctx(), count() * sizeof(float), &h), "cnn_host_alloc(slab)");
    dev = static_cast<float*>(d);
    host = static_cast<float*>(h);
}

Slab::~Slab() {
    cnn_sync(ctx());
    cnn_free(ctx(), dev);
    cnn_host_free(ctx(), host);
}

void Slab::to_host() {
    if (host_valid) return;
    check(cnn_d2h(ctx(), host, dev, count() * sizeof(float)), "cnn_d2h(slab)");
    host_valid = true;
}

void Slab::to_device() {
    if (dev_valid) return;
    check(cnn_h2d(ctx(), dev, host, count() * sizeof(float)), "cnn_h2d(slab)");
    dev_valid = true;
}

std::shared_ptr<Slab> make_slab(int B, int C, int H, int W) { return std::make_shared<Slab>(B, C, H, W); }

std::vector<tensor> make_views(const std::shared_ptr<Slab>& s, const std::string& prefix) {
    std::vector<tensor> v;
    v.reserve(s->B);
    for (int b = 0; b < s->B; ++b)
        v.emplace_back(new Tensor3D(s->C, s->H, s->W, prefix + std::to_string(b), s, b));
    return v;
}

float* batch_on_device(const std::vector<tensor>& v, std::shared_ptr<Slab>& staging,
                       std::shared_ptr<Slab>& used) {
    const int n = (int)v.size();
    bool in_place = n > 0 && v[0]->slab && n <= v[0]->slab->B;
    for (int b = 0; in_place && b < n; ++b) in_place = v[b]->slab == v[0]->slab && v[b]->slab_index == b;
    if (in_place) {
        used = v[0]->slab;
        used->to_device();
        return used->dev;
    }
    const Tensor3D& t0 = *v[0];
    if (!staging || staging->B < n || staging->C != t0.C || staging->H != t0.H || staging->W != t0.W)
        staging = make_slab(n, t0.C, t0.H, t0.W);
    cnn_sync(ctx());  // the previous upload of this staging mirror must have completed
    for (int b = 0; b < n; ++b) {
        v[b]->sync_host();
        std::memcpy(staging->host + (size_t)b * staging->per, v[b]->data, staging->per * sizeof(float));
    }
    staging->host_valid = true;
    staging->dev_valid = false;
    staging->to_device();
    used = staging;
    return staging->dev;
}

}  // namespace cnn_b200
