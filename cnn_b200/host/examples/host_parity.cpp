// host_parity.cpp -- drives the reference-shaped C++ layer classes (Conv2D, ReLU, MaxPool2D,
// LinearLayer, BatchNorm2D, softmax, cross_entroy_backward) exactly the way alexnet.cpp:12-65 and
// cnn.cpp:81-92 do, on synthetic input, and checks every step against the CPU oracle
// (oracle/cnn_oracle.c).  TEST PROGRAM: exit code 0 = parity within 1e-4.
//   host_parity [batch] [steps] [bn]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <list>
#include <vector>

#include "architectures.h"
#include "func.h"
#include "cnn_oracle.h"

using namespace architectures;

static float synth(unsigned long long seed, unsigned long long idx) {  // cnn_b200/synth.py
    unsigned long long z = (seed << 40) + idx + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (float)(z >> 40) * (1.0f / 16777216.0f);
}

static double rel_err(const float* a, const float* r, size_t n) {
    double num = 0, den = 1e-30;
    for (size_t i = 0; i < n; ++i) {
        num = std::fmax(num, std::fabs((double)a[i] - r[i]));
        den = std::fmax(den, std::fabs((double)r[i]));
    }
    return num / den;
}

int main(int argc, char** argv) {
    const int B = argc > 1 ? std::atoi(argv[1]) : 4, steps = argc > 2 ? std::atoi(argv[2]) : 2;
    const bool bn = argc > 3 && std::atoi(argv[3]) != 0;
    const int classes = 3;
    // the model of alexnet.cpp:12-31, built from the layer classes
    std::list<std::shared_ptr<Layer>> net;
    std::vector<orc_layer_spec> spec;
    const int chans[5] = {3, 16, 32, 64, 128};
    for (int i = 0; i < 4; ++i) {
        net.emplace_back(new Conv2D("conv_layer_" + std::to_string(i + 1), chans[i], chans[i + 1], 3));
        spec.push_back({ORC_CONV, chans[i], chans[i + 1], 3, 2});
        if (bn) {
            net.emplace_back(new BatchNorm2D("bn_layer_" + std::to_string(i + 1), chans[i + 1]));
            spec.push_back({ORC_BN, chans[i + 1], 0, 0, 0});
        }
        net.emplace_back(new ReLU("relu_layer_" + std::to_string(i + 1)));
        spec.push_back({ORC_RELU, 0, 0, 0, 0});
        if (i == 0) {
            net.emplace_back(new MaxPool2D("max_pool_1", 2, 2));
            spec.push_back({ORC_POOL, 2, 2, 0, 0});
        }
    }
    net.emplace_back(new LinearLayer("linear_1", 6 * 6 * 128, classes));
    spec.push_back({ORC_LINEAR, 6 * 6 * 128, classes, 0, 0});

    // initial parameters = what the constructors drew (reference seeds); hand them to the oracle
    // through the checkpoint format (also exercises save_weights)
    const char* ckpt = "/tmp/cnn_b200_host_parity.model";
    {
        std::ofstream w(ckpt, std::ios::binary);
        for (auto& l : net) l->save_weights(w);
    }
    orc_net* onet = orc_net_create(spec.data(), (int)spec.size(), B, 3, 224, 224);
    const long P = orc_net_param_count(onet);
    std::vector<float> params(P), oparams(P);
    {
        std::ifstream r(ckpt, std::ios::binary);
        r.read(reinterpret_cast<char*>(params.data()), P * sizeof(float));
        if (r.gcount() != (std::streamsize)(P * sizeof(float))) { std::printf("checkpoint size mismatch\n"); return 2; }
    }
    orc_net_set_params(onet, params.data());
    std::printf("params %ld, first conv weights %.10f %.10f %.10f\n", P, params[0], params[1], params[2]);

    const size_t per = 3 * 224 * 224;
    std::vector<float> x(B * per);
    std::vector<int> labels(B);
    std::vector<tensor> input;
    for (int b = 0; b < B; ++b) {
        tensor t(new Tensor3D(3, 224, 224));
        for (size_t i = 0; i < per; ++i) t->data[i] = x[b * per + i] = synth(1234, b * per + i);
        input.push_back(t);
        labels[b] = b % classes;
    }
    double worst = 0;
    std::vector<float> oprobs(B * classes), odx(B * per);
    for (int s = 0; s < steps; ++s) {
        std::vector<tensor> out(input);
        for (auto& l : net) out = l->forward(out);                        // alexnet.cpp:41-44
        const auto probs = softmax(out);                                     // cnn.cpp:83
        auto loss_delta = cross_entroy_backward(probs, one_hot(labels, classes));
        for (auto l = net.rbegin(); l != net.rend(); ++l) loss_delta.second = (*l)->backward(loss_delta.second);
        for (auto& l : net) l->update_gradients(1e-3f);                      // cnn.cpp:90
        const float oloss = orc_net_train_step(onet, x.data(), labels.data(), 1e-3f, oprobs.data(), odx.data());
        std::vector<float> p(B * classes), dx(B * per);
        for (int b = 0; b < B; ++b) {
            for (int i = 0; i < classes; ++i) p[b * classes + i] = probs[b]->data[i];
            loss_delta.second[b]->sync_host();
            std::copy(loss_delta.second[b]->data, loss_delta.second[b]->data + per, dx.begin() + b * per);
            if (probs[b]->argmax() != orc_argmax(&oprobs[b * classes], classes)) { std::printf("argmax mismatch\n"); return 1; }
        }
        const double e_p = rel_err(p.data(), oprobs.data(), p.size()), e_dx = rel_err(dx.data(), odx.data(), dx.size());
        const double e_l = std::fabs(loss_delta.first - oloss) / std::fmax(1.0, std::fabs(oloss));
        std::printf("step %d loss %.6f (oracle %.6f) rel.err loss %.2e probs %.2e image-grad %.2e\n", s, loss_delta.first,
                    oloss, e_l, e_p, e_dx);
        worst = std::fmax(worst, std::fmax(e_p, std::fmax(e_dx, e_l)));
    }
    {   // updated weights through save_weights vs the oracle's
        std::ofstream w(ckpt, std::ios::binary);
        for (auto& l : net) l->save_weights(w);
    }
    {
        std::ifstream r(ckpt, std::ios::binary);
        r.read(reinterpret_cast<char*>(params.data()), P * sizeof(float));
    }
    orc_net_get_params(onet, oparams.data());
    const double e_w = rel_err(params.data(), oparams.data(), P);
    std::printf("weights after %d steps rel.err %.2e\n", steps, e_w);
    worst = std::fmax(worst, e_w);
    // gradCAM-style access: Layer::get_output() of a pre-ReLU conv layer must be host-readable
    auto it = net.begin();
    const auto conv1 = (*it)->get_output();
    long cnt = 0;
    const float* oconv1 = orc_net_layer_output(onet, 0, &cnt);
    std::vector<float> c1(cnt);
    for (int b = 0; b < B; ++b) std::copy(conv1[b]->data, conv1[b]->data + conv1[b]->get_length(), c1.begin() + (size_t)b * conv1[b]->get_length());
    const double e_c = rel_err(c1.data(), oconv1, cnt);
    std::printf("get_output(conv_layer_1) rel.err %.2e\n", e_c);
    worst = std::fmax(worst, e_c);
    orc_net_destroy(onet);
    std::printf("%s (worst %.2e, bar 1e-4)\n", worst <= 1e-4 ? "PARITY OK" : "PARITY FAILED", worst);
    return worst <= 1e-4 ? 0 : 1;
}
