// data_format.cpp -- Tensor3D methods (reference semantics: cpu/src/data_format.cpp) on top of the
// slab storage.  Reading methods refresh the host mirror first; mutating ones mark the device stale.
#include <cstring>
#include <iomanip>
#include <iostream>

#include "backend.h"

namespace {
data_type* plain_alloc(size_t n) { return new data_type[n]; }
}

Tensor3D::Tensor3D(const int _C, const int _H, const int _W, const std::string _name)
    : C(_C), H(_H), W(_W), data(plain_alloc((size_t)_C * _H * _W)), name(std::move(_name)) {}

Tensor3D::Tensor3D(const std::tuple<int, int, int>& shape, const std::string _name)
    : C(std::get<0>(shape)), H(std::get<1>(shape)), W(std::get<2>(shape)),
      data(plain_alloc((size_t)std::get<0>(shape) * std::get<1>(shape) * std::get<2>(shape))),
      name(std::move(_name)) {}

Tensor3D::Tensor3D(const int length, const std::string _name)
    : C(length), H(1), W(1), data(plain_alloc((size_t)length)), name(std::move(_name)) {}

Tensor3D::Tensor3D(const int _C, const int _H, const int _W, std::string _name,
                   std::shared_ptr<cnn_b200::Slab> s, int index)
    : C(_C), H(_H), W(_W), data(s->host + (size_t)index * s->per), name(std::move(_name)),
      slab(std::move(s)), slab_index(index) {}

Tensor3D::~Tensor3D() noexcept {
    if (!slab && data != nullptr) delete[] data;  // views do not own their bytes
    data = nullptr;
}

void Tensor3D::sync_host() const {
    if (slab) slab->to_host();
}

void Tensor3D::host_written() {
    // the whole B-image slab flips to "host is current": the OTHER images' host bytes must be current first,
    // otherwise the next upload would overwrite their device results with a stale mirror
    if (slab) { slab->to_host(); slab->host_valid = true; slab->dev_valid = false; }
}

// planes in OpenCV order B,G,R, value u8 * 1/255 (data_format.cpp:13-23)
void Tensor3D::read_from_opencv_mat(const uchar* const img_ptr) {
    sync_host();
    const int plane = H * W;
    for (int i = 0; i < plane; ++i) {
        const uchar* px = img_ptr + 3 * i;
        data[i] = px[0] * 1.f / 255;
        data[plane + i] = px[1] * 1.f / 255;
        data[2 * plane + i] = px[2] * 1.f / 255;
    }
    host_written();
}

void Tensor3D::set_zero() {
    sync_host();   // only this view is zeroed; the rest of the slab keeps its (device) contents
    std::memset(data, 0, sizeof(data_type) * (size_t)C * H * W);
    host_written();
}

int Tensor3D::argmax() const {  // first maximum wins (strict >), data_format.cpp:37-48
    if (data == nullptr) return 0;
    sync_host();
    const int n = C * H * W;
    int best = 0;
    for (int i = 1; i < n; ++i)
        if (data[i] > data[best]) best = i;
    return best;
}

int Tensor3D::argmin() const {
    if (data == nullptr) return 0;
    sync_host();
    const int n = C * H * W;
    int best = 0;
    for (int i = 1; i < n; ++i)
        if (data[i] < data[best]) best = i;
    return best;
}

data_type Tensor3D::max() const { return data[argmax()]; }
data_type Tensor3D::min() const { return data[argmin()]; }

void Tensor3D::div(const data_type times) {
    sync_host();
    const int n = C * H * W;
    for (int i = 0; i < n; ++i) data[i] /= times;
    host_written();
}

void Tensor3D::normalize(const std::vector<data_type> mean, const std::vector<data_type> std_div) {
    if (C != 3) return;
    sync_host();
    const int plane = H * W;
    for (int ch = 0; ch < C; ++ch)
        for (int i = 0; i < plane; ++i) data[ch * plane + i] = (data[ch * plane + i] - mean[ch]) / std_div[ch];
    host_written();
}

cv::Mat Tensor3D::opecv_mat(const int CH) const {
    sync_host();
    cv::Mat img;
    const int plane = H * W;
    if (CH == 3) {
        img = cv::Mat(H, W, CV_8UC3);
        for (int i = 0; i < plane; ++i)
            for (int ch = 0; ch < 3; ++ch)
                img.data[3 * i + ch] = cv::saturate_cast<uchar>(255 * data[i + ch * plane]);
    } else if (CH == 1) {
        img = cv::Mat(H, W, CV_8UC1);
        for (int i = 0; i < plane; ++i) img.data[i] = cv::saturate_cast<uchar>(255 * data[i]);
    }
    return img;
}

int Tensor3D::get_length() const { return C * H * W; }
std::tuple<int, int, int> Tensor3D::get_shape() const { return std::make_tuple(C, H, W); }

void Tensor3D::print_shape() const {
    std::cout << name << "  ==>  " << C << " x " << H << " x " << W << "\n";
}

void Tensor3D::print(const int _C) const {
    sync_host();
    std::cout << name << "  content is : ";
    for (int i = 0; i < H; ++i) {
        for (int j = 0; j < W; ++j)
            std::cout << std::setiosflags(std::ios::fixed) << std::setprecision(3) << data[(_C * H + i) * W + j] << "   ";
        std::cout << "\n";
    }
}

std::shared_ptr<Tensor3D> Tensor3D::rot180() const {
    sync_host();
    std::shared_ptr<Tensor3D> out(new Tensor3D(C, H, W, name + "_rot180"));
    const int plane = H * W;
    for (int c = 0; c < C; ++c)
        for (int i = 0; i < plane; ++i) out->data[c * plane + i] = data[c * plane + plane - 1 - i];
    return out;
}

std::shared_ptr<Tensor3D> Tensor3D::pad(const int padding) const {
    sync_host();
    const int PH = H + 2 * padding, PW = W + 2 * padding;
    std::shared_ptr<Tensor3D> out(new Tensor3D(C, PH, PW, name + "_rot180"));
    std::memset(out->data, 0, sizeof(data_type) * (size_t)C * PH * PW);
    for (int c = 0; c < C; ++c)
        for (int i = 0; i < H; ++i)
            std::memcpy(out->data + ((size_t)c * PH + padding + i) * PW + padding, data + ((size_t)c * H + i) * W,
                        W * sizeof(data_type));
    return out;
}
