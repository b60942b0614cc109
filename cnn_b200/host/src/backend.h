// backend.h -- internal glue between the reference-shaped C++ classes and the C ABI
// (include/cnn_b200.h).  Not installed; the public headers only forward-declare these types.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>

#include "cnn_b200.h"
#include "data_format.h"

namespace cnn_b200 {

cnn_ctx* ctx();  // process-wide context (device CNN_B200_DEVICE, default 0), created on first use

// The reference has no error channel (asserts / UB); the backend aborts with the C-ABI message.
inline void check(int rc, const char* what) {
    if (rc != 0) {
        std::fprintf(stderr, "cnn_b200: %s failed (%d): %s\n", what, rc, cnn_last_error());
        std::abort();
    }
}

struct DeviceBuffer {
    void* dev = nullptr;
    size_t bytes = 0;
    explicit DeviceBuffer(size_t n);
    ~DeviceBuffer();
    float* f() const { return static_cast<float*>(dev); }
    int32_t* i() const { return static_cast<int32_t*>(dev); }
};

// One batch: B images of C*H*W floats in ONE device allocation plus a pinned host mirror.
struct Slab {
    int B, C, H, W;
    size_t per;          // floats per image
    float* dev = nullptr;
    float* host = nullptr;
    bool host_valid = true;   // mirror holds the latest values
    bool dev_valid = false;   // device holds the latest values
    Slab(int B, int C, int H, int W);
    ~Slab();
    void to_host();     // D2H when the device is newer (blocks)
    void to_device();   // H2D when the host is newer (async on the context stream)
    void device_written() { dev_valid = true; host_valid = false; }
    size_t count() const { return per * (size_t)B; }
};

std::shared_ptr<Slab> make_slab(int B, int C, int H, int W);
// B view tensors "prefix_<b>" over a slab (Layer::output and friends)
std::vector<tensor> make_views(const std::shared_ptr<Slab>& s, const std::string& prefix);
// Device pointer of a batch.  Views 0..n-1 of one slab are used in place; anything else (tensors
// made with `new Tensor3D`, scattered views) is packed into `staging` and uploaded.  `used`
// receives the slab the pointer belongs to (keeps it alive, lets callers mark it written).
float* batch_on_device(const std::vector<tensor>& v, std::shared_ptr<Slab>& staging,
                       std::shared_ptr<Slab>& used);

}  // namespace cnn_b200
