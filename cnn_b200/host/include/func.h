// func.h -- same free functions as the reference's cpu/include/func.h:8-18, run on the device
// through libcnn_b200 (cnn_softmax_xent / cnn_xent_backward); results come back host-readable.
#ifndef CNN_FUNC_H
#define CNN_FUNC_H

#include "data_format.h"

// row softmax with the reference's clamped exp (func.cpp:7-37)
std::vector<tensor> softmax(const std::vector<tensor>& input);

// labels -> one-hot rows (func.cpp:40-53)
std::vector<tensor> one_hot(const std::vector<int>& labels, const int num_classes);

// loss = -(1/B) sum log(p) y ; delta = p - y   (func.cpp:56-73; the name keeps the reference's typo)
std::pair<data_type, std::vector<tensor> > cross_entroy_backward(
        const std::vector<tensor>& probs, const std::vector<tensor>& labels);

std::string float_to_string(const float value, const int precision);

#endif  // CNN_FUNC_H
