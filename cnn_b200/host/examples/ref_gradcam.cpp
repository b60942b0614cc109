// ref_gradcam.cpp -- the per-image body of the reference's grad_cam.cpp:61-80 (read_from_opencv_mat -> forward ->
// softmax -> AlexNet::grad_cam("conv_layer_3")) with the reference's OWN, unmodified alexnet.cpp on top of the
// B200 layer classes.  Images arrive as raw interleaved u8 HWC bytes (what cv::imread + cv::resize hand to
// Tensor3D::read_from_opencv_mat); the 6x6 8-bit map of every image is printed.  Built by `make refcheck`.
//   ref_gradcam <checkpoint.model> <images.u8 raw: N x 224 x 224 x 3> <N>
#include <cstdio>
#include <fstream>
#include <vector>

#include "architectures.h"
#include "func.h"

int main(int argc, char** argv) {
    using namespace architectures;
    if (argc < 4) return 2;
    const int N = std::atoi(argv[3]);
    AlexNet network(3, false);
    network.load_weights(argv[1]);
    std::ifstream f(argv[2], std::ios::binary);
    std::vector<unsigned char> raw((size_t)N * 224 * 224 * 3);
    f.read(reinterpret_cast<char*>(raw.data()), (std::streamsize)raw.size());
    if (!f) return 3;
    no_grad = false;   // grad_cam.cpp:57
    std::vector<tensor> image_buffer({tensor(new Tensor3D(3, 224, 224))});
    for (int i = 0; i < N; ++i) {
        image_buffer[0]->read_from_opencv_mat(raw.data() + (size_t)i * 224 * 224 * 3);
        const auto output = network.forward(image_buffer);
        const auto prob = softmax(output);
        const int max_index = prob[0]->argmax();
        std::printf("image %d class %d prob %.6f cam", i, max_index, prob[0]->data[max_index]);
        const cv::Mat cam = network.grad_cam("conv_layer_3");
        for (int j = 0; j < cam.rows * cam.cols; ++j) std::printf(" %d", (int)cam.data[j]);
        std::printf("\n");
    }
    return 0;
}
