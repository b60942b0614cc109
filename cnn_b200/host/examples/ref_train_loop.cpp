// ref_train_loop.cpp -- the body of the reference's train loop (cpu/src/cnn.cpp:54,81-92) on
// synthetic tensors, using the reference's OWN, unmodified AlexNet container (alexnet.cpp) on top
// of the B200 layer classes.  Built by `make refcheck`; proves the drop-in boundary compiles and
// links.  (cnn.cpp itself additionally needs OpenCV imread/resize and the dataset.)
#include <cstdio>
#include <vector>

#include "architectures.h"
#include "func.h"
#include "metrics.h"

int main(int argc, char** argv) {
    using namespace architectures;
    const int B = 4, classes = 3, iters = argc > 1 ? std::atoi(argv[1]) : 3;
    AlexNet network(classes, false);
    std::vector<tensor> batch;
    std::vector<int> labels;
    for (int b = 0; b < B; ++b) {
        tensor t(new Tensor3D(3, 224, 224));
        for (int i = 0; i < t->get_length(); ++i) t->data[i] = ((i * 2654435761u + b * 40503u) >> 8 & 0xFFFF) / 65536.f;
        batch.push_back(t);
        labels.push_back(b % classes);
    }
    ClassificationEvaluator evaluator;
    std::vector<int> predict(B, -1);
    for (int iter = 1; iter <= iters; ++iter) {
        const auto output = network.forward(batch);
        const auto probs = softmax(output);
        auto loss_delta = cross_entroy_backward(probs, one_hot(labels, classes));
        network.backward(loss_delta.second);
        network.update_gradients(1e-3f);
        for (int b = 0; b < B; ++b) predict[b] = probs[b]->argmax();
        evaluator.compute(predict, labels);
        std::printf("Train===> [batch %d/%d] [loss %.3f] [Accuracy %4.3f]\n", iter, iters, loss_delta.first, evaluator.get());
    }
    network.save_weights("/tmp/cnn_b200_ref_loop.model");
    return 0;
}
