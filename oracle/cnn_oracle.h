/*
 * cnn_oracle.h -- CPU oracle for the conv-net train-step hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's
 * (hermosayhl/CNN, cpu/src) arithmetic, loop order for loop order, so that its
 * results are bit-identical to the reference compiled with its canonical flags
 * (-std=c++17 -O2, no -march, no -ffast-math).  Only tests/, bench.py's
 * cpu_baseline / --impl reference leg and __graft_entry__.smoke() may use it;
 * the product path (cnn_b200/) never links or calls it.
 *
 * Parity pin: oracle/_ref (the reference's own sources compiled in place) is
 * compared against this file bit-for-bit in tests/test_oracle_vs_ref.py, and
 * the README inference known-answer (dog 0.850634 / panda 0.999978 /
 * bird 0.999998) is checked in tests/test_oracle_golden.py.
 *
 * Layout: a batch is one contiguous [B][C][H][W] fp32 array; image b is the
 * reference's b-th Tensor3D (CHW, index c*H*W + h*W + w, data_format.h:11-26).
 */
#ifndef CNN_ORACLE_H
#define CNN_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* Conv2D::forward, conv2d.cpp:34-94. w is [Cout][Cin][k][k], no padding. */
void orc_conv2d_forward(const float* x, const float* w, const float* bias, float* y,
                        int B, int Cin, int H, int W, int Cout, int k, int stride);

/* Conv2D::backward, conv2d.cpp:97-202: dw/db (mean over batch) and dx (scatter form). */
void orc_conv2d_backward(const float* x, const float* w, const float* delta,
                         float* dw, float* db, float* dx,
                         int B, int Cin, int H, int W, int Cout, int k, int stride);

/* MaxPool2D::forward, pool2d.cpp:7-89.  mask may be NULL (no_grad). */
void orc_maxpool_forward(const float* x, float* y, int* mask,
                         int B, int C, int H, int W, int k, int step);

/* MaxPool2D::backward, pool2d.cpp:92-109 (zero, then scatter-ASSIGN). */
void orc_maxpool_backward(const float* delta, const int* mask, float* dx,
                          int B, int C, int H, int W, int OH, int OW);

/* ReLU::forward relu.cpp:9-28 / ReLU::backward relu.cpp:30-44 (in place on delta). */
void orc_relu_forward(const float* x, float* y, long n);
void orc_relu_backward(float* delta, const float* y, long n);

/* LinearLayer::forward linear.cpp:22-45; w is [in][out]. */
void orc_linear_forward(const float* x, const float* w, const float* bias, float* y,
                        int B, int in, int out);
/* LinearLayer::backward linear.cpp:47-93. */
void orc_linear_backward(const float* x, const float* w, const float* delta,
                         float* dw, float* db, float* dx, int B, int in, int out);

/* BatchNorm2D::forward, batchnorm2d.cpp:24-97 (train branch :44-80, eval :81-95). */
void orc_bn_forward_train(const float* x, const float* gamma, const float* beta,
                          float* moving_mean, float* moving_var,
                          float* batch_mean, float* batch_var,
                          float* xhat, float* y,
                          int B, int C, int H, int W, float eps, float momentum);
void orc_bn_forward_eval(const float* x, const float* gamma, const float* beta,
                         const float* moving_mean, const float* moving_var,
                         float* xhat, float* y,
                         int B, int C, int H, int W, float eps);
/* BatchNorm2D::backward, batchnorm2d.cpp:100-158 (in place on delta). */
void orc_bn_backward(float* delta, const float* x, const float* xhat,
                     const float* gamma, const float* batch_mean, const float* batch_var,
                     float* dgamma, float* dbeta,
                     int B, int C, int H, int W, float eps);

/* softmax func.cpp:16-37, argmax data_format.cpp:37-48. */
void orc_softmax(const float* logits, float* probs, int B, int n);
int orc_argmax(const float* v, int n);
/* one_hot func.cpp:40-53 + cross_entroy_backward func.cpp:56-73. Returns loss. */
float orc_cross_entropy_backward(const float* probs, const int* labels, float* delta,
                                 int B, int n);

/* <Layer>::update_gradients: p -= lr*g (conv2d.cpp:205-217, linear.cpp:95-102,
 * batchnorm2d.cpp:161-166). */
void orc_sgd(float* p, const float* g, long n, float lr);

/* ---- whole-network runner (AlexNet::forward/backward/update_gradients,
 *      alexnet.cpp:35-65, driven like cnn.cpp:81-92) ---------------------- */
enum { ORC_CONV = 0, ORC_BN = 1, ORC_RELU = 2, ORC_POOL = 3, ORC_LINEAR = 4,
       ORC_PAD = 5, ORC_AVGPOOL = 6 /* extensions, see below */ };

/* Extensions that are NOT in the reference (items 7-8 of its TODO list, cnn.cpp:15-24): zero padding as a layer
 * (PAD: a = border width) and average pooling (AVGPOOL: a = k, b = step).  Checked against this definition only. */
void orc_pad_forward(const float* x, float* y, int B, int C, int H, int W, int pad);
void orc_pad_backward(const float* delta, float* dx, int B, int C, int H, int W, int pad);
void orc_avgpool_forward(const float* x, float* y, int B, int C, int H, int W, int k, int step);
void orc_avgpool_backward(const float* delta, float* dx, int B, int C, int H, int W, int k, int step);

typedef struct {
    int type;
    int a, b, c, d; /* CONV: cin,cout,k,stride; BN: channels; POOL: k,step; LINEAR: in,out */
} orc_layer_spec;

typedef struct orc_net orc_net;

/* Builds the net for inputs [B][C][H][W]; params start at 0 (caller loads them). */
orc_net* orc_net_create(const orc_layer_spec* specs, int n_layers, int B, int C, int H, int W);
void orc_net_destroy(orc_net* net);
long orc_net_param_count(const orc_net* net);   /* checkpoint float count (BN: 4 arrays) */
void orc_net_set_params(orc_net* net, const float* flat);
void orc_net_get_params(const orc_net* net, float* flat);
void orc_net_get_grads(const orc_net* net, float* flat); /* same order; BN moving stats -> 0 */
int orc_net_num_classes(const orc_net* net);
/* forward only; no_grad!=0 selects BN eval branch and skips pool masks. */
void orc_net_forward(orc_net* net, const float* x, float* logits, int no_grad);
/* one train step: forward, softmax, xent, backward, sgd. probs/dx_image may be NULL. */
float orc_net_train_step(orc_net* net, const float* x, const int* labels, float lr,
                         float* probs, float* dx_image);
/* pointer to layer i's output / incoming-delta buffers (for layer-wise parity). */
const float* orc_net_layer_output(const orc_net* net, int layer, long* count);

#ifdef __cplusplus
}
#endif
#endif
