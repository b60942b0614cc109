// architectures.h -- the reference's layer API (cpu/include/architectures.h:12-215) backed by
// libcnn_b200: identical class names, constructor / method signatures, globals and the AlexNet
// declaration, so alexnet.cpp, cnn.cpp, inference.cpp and grad_cam.cpp compile unmodified.
// Private members are new: parameters, gradients and activations live in device slabs.
#ifndef CNN_ARCHITECTURES_H
#define CNN_ARCHITECTURES_H

#include <filesystem>
#include <fstream>
#include <list>
#include <random>

// This backend's Tensor3D must be seen first: the reference's pipeline.h includes "data_format.h"
// next to itself, and both headers share the include guard CNN_DATA_FORMAT_H.
#include "data_format.h"
#if __has_include("pipeline.h")
#include "pipeline.h"   // the reference's data pipeline stays as it is
#else
namespace pipeline {}
#endif

namespace cnn_b200 { struct DeviceBuffer; struct Slab; }

namespace architectures {
    using namespace pipeline;

    extern data_type random_times;   // divisor of every N(0,1) initial weight (architectures.cpp:6)
    extern bool no_grad;             // set by WithoutGrad (architectures.cpp:8)

    class WithoutGrad final {
    public:
        explicit WithoutGrad() { architectures::no_grad = true; }
        ~WithoutGrad() noexcept { architectures::no_grad = false; }
    };

    class Layer {
    public:
        const std::string name;
        std::vector<tensor> output;
    public:
        Layer(std::string& _name) : name(std::move(_name)) {}
        virtual ~Layer() = default;
        virtual std::vector<tensor> forward(const std::vector<tensor>& input) = 0;
        virtual std::vector<tensor> backward(std::vector<tensor>& delta) = 0;
        virtual void update_gradients(const data_type learning_rate = 1e-4) {}
        virtual void save_weights(std::ofstream& writer) const {}
        virtual void load_weights(std::ifstream& reader) {}
        // host-readable outputs: refreshes the pinned mirror of the device slab first
        virtual std::vector<tensor> get_output() const;
    };

    class Conv2D : public Layer {
    private:
        const int in_channels, out_channels, kernel_size, stride;
        const int params_for_one_kernel;
        std::default_random_engine seed;
        // [Cout*Cin*k*k filters | Cout biases] in checkpoint order, host copy + device copy
        mutable std::vector<data_type> host_params;
        std::shared_ptr<cnn_b200::DeviceBuffer> params, grads;
        mutable bool host_params_stale = false;
        bool device_params_stale = true;
        std::shared_ptr<cnn_b200::Slab> in_slab, out_slab, dx_slab;   // saved input / output / delta_output
        std::vector<tensor> delta_output;
        int in_H = 0, in_W = 0, in_B = 0;
        void upload_params();
        void download_params() const;
    public:
        Conv2D(std::string _name, const int _in_channels = 3, const int _out_channels = 16,
               const int _kernel_size = 3, const int _stride = 2);
        std::vector<tensor> forward(const std::vector<tensor>& input);
        std::vector<tensor> backward(std::vector<tensor>& delta);
        void update_gradients(const data_type learning_rate = 1e-4);
        void save_weights(std::ofstream& writer) const;
        void load_weights(std::ifstream& reader);
        int get_params_num() const;
    };

    class MaxPool2D : public Layer {
    private:
        const int kernel_size, step, padding;
        std::shared_ptr<cnn_b200::Slab> out_slab, dx_slab;
        std::shared_ptr<cnn_b200::DeviceBuffer> mask;
        std::vector<tensor> delta_output;
        int in_C = 0, in_H = 0, in_W = 0;
    public:
        MaxPool2D(std::string _name, const int _kernel_size = 2, const int _step = 2)
                : Layer(_name), kernel_size(_kernel_size), step(_step), padding(0) {}
        std::vector<tensor> forward(const std::vector<tensor>& input);
        std::vector<tensor> backward(std::vector<tensor>& delta);
    };

    class ReLU : public Layer {
    private:
        std::shared_ptr<cnn_b200::Slab> out_slab;
    public:
        ReLU(std::string _name) : Layer(_name) {}
        std::vector<tensor> forward(const std::vector<tensor>& input);
        std::vector<tensor> backward(std::vector<tensor>& delta);
    };

    class LinearLayer : public Layer {
    private:
        const int in_channels, out_channels;
        mutable std::vector<data_type> host_params;   // [in*out weights | out biases]
        std::shared_ptr<cnn_b200::DeviceBuffer> params, grads;
        mutable bool host_params_stale = false;
        bool device_params_stale = true;
        std::tuple<int, int, int> delta_shape;
        std::shared_ptr<cnn_b200::Slab> in_slab, out_slab, dx_slab;
        std::vector<tensor> delta_output;
        void upload_params();
        void download_params() const;
    public:
        LinearLayer(std::string _name, const int _in_channels, const int _out_channels);
        std::vector<tensor> forward(const std::vector<tensor>& input);
        std::vector<tensor> backward(std::vector<tensor>& delta);
        void update_gradients(const data_type learning_rate = 1e-4);
        void save_weights(std::ofstream& writer) const;
        void load_weights(std::ifstream& reader);
    };

    class BatchNorm2D : public Layer {
    private:
        const int out_channels;
        const data_type eps, momentum;
        mutable std::vector<data_type> host_params;   // [gamma | beta | moving_mean | moving_var]
        std::shared_ptr<cnn_b200::DeviceBuffer> params, grads, batch_stats;
        mutable bool host_params_stale = false;
        bool device_params_stale = true;
        std::shared_ptr<cnn_b200::Slab> in_slab, out_slab, xhat_slab;
        void upload_params();
        void download_params() const;
    public:
        BatchNorm2D(std::string _name, const int _out_channels, const data_type _eps = 1e-5,
                    const data_type _momentum = 0.1);
        std::vector<tensor> forward(const std::vector<tensor>& input);
        std::vector<tensor> backward(std::vector<tensor>& delta);
        void update_gradients(const data_type learning_rate = 1e-4);
        void save_weights(std::ofstream& writer) const;
        void load_weights(std::ifstream& reader);
    };

    // Declared exactly as the reference does (architectures.h:196-215); its definition is the
    // reference's own, unmodified alexnet.cpp (thin glue, stays).
    class AlexNet {
    public:
        bool print_info = false;
    private:
        std::list< std::shared_ptr<Layer> > layers_sequence;
    public:
        AlexNet(const int num_classes = 3, const bool batch_norm = false);
        std::vector<tensor> forward(const std::vector<tensor>& input);
        void backward(std::vector<tensor>& delta_start);
        void update_gradients(const data_type learning_rate = 1e-4);
        void save_weights(const std::filesystem::path& save_path) const;
        void load_weights(const std::filesystem::path& checkpoint_path);
        cv::Mat grad_cam(const std::string& layer_name) const;
    };
}

#endif  // CNN_ARCHITECTURES_H
