// data_format.h -- Tensor3D of the B200 backend: same public surface as the reference's
// cpu/include/data_format.h:10-53 (C, H, W, data, name, every method, `tensor` alias), new storage.
//
// `data` is always a HOST pointer user code may dereference (pipeline.cpp fills it, func.cpp /
// cnn.cpp / alexnet.cpp:105-129 read it).  A tensor produced by a device layer is a VIEW of one
// image inside that layer's contiguous [B][C][H][W] slab pair (device + pinned host mirror); the
// mirror is refreshed lazily by sync_host() -- called by every reading method here and by
// Layer::get_output().  Tensors made with `new Tensor3D(...)` own plain host memory exactly like
// the reference, and are uploaded when a layer consumes them.
#ifndef CNN_DATA_FORMAT_H
#define CNN_DATA_FORMAT_H

#include <opencv2/core.hpp>

#include <memory>
#include <string>
#include <tuple>
#include <vector>

using data_type = float;

namespace cnn_b200 { struct Slab; }

class Tensor3D {
public:
    const int C, H, W;
    data_type* data;
    std::string name;
    Tensor3D(const int _C, const int _H, const int _W, const std::string _name = "pipeline");
    Tensor3D(const std::tuple<int, int, int>& shape, const std::string _name = "pipeline");
    Tensor3D(const int length, const std::string _name = "pipeline");
    void read_from_opencv_mat(const uchar* const img_ptr);
    void set_zero();
    data_type max() const;
    int argmax() const;
    data_type min() const;
    int argmin() const;
    void div(const data_type times);
    void normalize(const std::vector<data_type> mean = {0.406, 0.456, 0.485},
                   const std::vector<data_type> std_div = {0.225, 0.224, 0.229});
    cv::Mat opecv_mat(const int CH = 3) const;
    int get_length() const;
    std::tuple<int, int, int> get_shape() const;
    void print_shape() const;
    void print(const int _C = 0) const;
    std::shared_ptr<Tensor3D> rot180() const;
    std::shared_ptr<Tensor3D> pad(const int padding = 1) const;
    ~Tensor3D() noexcept;

    // ---- backend extension (not in the reference) ------------------------------------------
    // View of image `index` inside a slab; used by the layers only.
    Tensor3D(const int _C, const int _H, const int _W, std::string _name,
             std::shared_ptr<cnn_b200::Slab> slab, int index);
    void sync_host() const;                      // make `data` current (D2H of the slab if stale)
    void host_written();                         // host contents changed: device copy is stale
    std::shared_ptr<cnn_b200::Slab> slab;        // null for plain host tensors
    int slab_index = 0;
    Tensor3D(const Tensor3D&) = delete;
    Tensor3D& operator=(const Tensor3D&) = delete;
};
using tensor = std::shared_ptr<Tensor3D>;

#endif  // CNN_DATA_FORMAT_H
