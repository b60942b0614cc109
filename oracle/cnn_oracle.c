/*
 * cnn_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY, see cnn_oracle.h).
 *
 * Every function restates one reference function with the same loop nesting,
 * the same fp32 accumulation order and the same mixed float/double expression
 * shapes, so that with -O2 -ffp-contract=off it reproduces the reference's
 * bits.  Citations are file:line under /root/reference/cpu/src.
 */
#include "cnn_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ conv */

/* conv2d.cpp:34-94.  Window centres start at `radius` and step by stride while
 * < H-radius (:76-77); accumulate i-major then window, bias last (:78-86). */
void orc_conv2d_forward(const float* x, const float* w, const float* bias, float* y,
                        int B, int Cin, int H, int W, int Cout, int k, int stride) {
    const int OH = (H - k) / stride + 1, OW = (W - k) / stride + 1;
    const int radius = (k - 1) / 2, win = k * k, plane = H * W, oplane = OH * OW;
    int* off = (int*)malloc(sizeof(int) * win);
    int pos = 0;
    for (int dx = -radius; dx <= radius; ++dx)
        for (int dy = -radius; dy <= radius; ++dy) off[pos++] = dx * W + dy; /* :54-59 */
    for (int b = 0; b < B; ++b) {
        const float* img = x + (long)b * Cin * plane;
        for (int o = 0; o < Cout; ++o) {
            float* out = y + ((long)b * Cout + o) * oplane;
            const float* wo = w + (long)o * Cin * win;
            int cnt = 0;
            for (int r = radius; r < H - radius; r += stride)
                for (int c = radius; c < W - radius; c += stride) {
                    float sum = 0.f;
                    const int coord = r * W + c;
                    for (int i = 0; i < Cin; ++i) {
                        const int start = i * plane + coord, sw = i * win;
                        for (int t = 0; t < win; ++t) sum += img[start + off[t]] * wo[sw + t];
                    }
                    sum += bias[o];
                    out[cnt++] = sum;
                }
        }
    }
    free(off);
}

/* conv2d.cpp:97-202: wgrad :120-151 (per-image sum, /B, += over b), bgrad
 * :153-157, dgrad scatter :175-199 after zeroing :168. */
void orc_conv2d_backward(const float* x, const float* w, const float* delta,
                         float* dw, float* db, float* dx,
                         int B, int Cin, int H, int W, int Cout, int k, int stride) {
    const int OH = (H - k) / stride + 1, OW = (W - k) / stride + 1;
    const int win = k * k, plane = H * W, oplane = OH * OW;
    memset(dw, 0, sizeof(float) * (size_t)Cout * Cin * win);
    for (int o = 0; o < Cout; ++o) db[o] = 0;
    for (int b = 0; b < B; ++b)
        for (int o = 0; o < Cout; ++o) {
            const float* od = delta + ((long)b * Cout + o) * oplane;
            for (int i = 0; i < Cin; ++i) {
                const float* in = x + ((long)b * Cin + i) * plane;
                float* wp = dw + ((long)o * Cin + i) * win;
                for (int kx = 0; kx < k; ++kx)
                    for (int ky = 0; ky < k; ++ky) {
                        float sum = 0;
                        for (int r = 0; r < OH; ++r) {
                            const float* dp = od + r * OW;
                            const float* ip = in + (r * stride + kx) * W;
                            for (int c = 0; c < OW; ++c) sum += dp[c] * ip[c * stride + ky];
                        }
                        wp[kx * k + ky] += sum / B;
                    }
            }
            float sum = 0;
            for (int d = 0; d < oplane; ++d) sum += od[d];
            db[o] += sum / B;
        }
    if (!dx) return;
    memset(dx, 0, sizeof(float) * (size_t)B * Cin * plane);
    const int radius = (k - 1) / 2;
    int* off = (int*)malloc(sizeof(int) * win);
    int pos = 0;
    for (int a = -radius; a <= radius; ++a)
        for (int c = -radius; c <= radius; ++c) off[pos++] = a * W + c;
    for (int b = 0; b < B; ++b) {
        float* img = dx + (long)b * Cin * plane;
        for (int o = 0; o < Cout; ++o) {
            const float* od = delta + ((long)b * Cout + o) * oplane;
            const float* wo = w + (long)o * Cin * win;
            int cnt = 0;
            for (int r = radius; r < H - radius; r += stride)
                for (int c = radius; c < W - radius; c += stride) {
                    const int coord = r * W + c;
                    for (int i = 0; i < Cin; ++i) {
                        const int start = i * plane + coord, sw = i * win;
                        for (int t = 0; t < win; ++t) img[start + off[t]] += wo[sw + t] * od[cnt];
                    }
                    ++cnt;
                }
        }
    }
    free(off);
}

/* ------------------------------------------------------------------ pool */

/* pool2d.cpp:53-87: first element seeds the max, strict '<' (first max wins,
 * NaN never replaces), mask = flat CHW index into the image (:79-82). */
void orc_maxpool_forward(const float* x, float* y, int* mask,
                         int B, int C, int H, int W, int k, int step) {
    const int OH = (H - k) / step + 1, OW = (W - k) / step + 1;
    const int plane = H * W, oplane = OH * OW, win = k * k;
    int* off = (int*)malloc(sizeof(int) * win);
    int pos = 0;
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j) off[pos++] = i * W + j; /* :35-38 */
    for (int b = 0; b < B; ++b)
        for (int ch = 0; ch < C; ++ch) {
            const float* img = x + ((long)b * C + ch) * plane;
            float* out = y + ((long)b * C + ch) * oplane;
            int* mp = mask ? mask + ((long)b * C + ch) * oplane : NULL;
            int cnt = 0;
            for (int r = 0; r <= H - k; r += step) {
                const float* row = img + r * W;
                for (int c = 0; c <= W - k; c += step) {
                    float mv = row[c];
                    int mi = 0;
                    for (int t = 1; t < win; ++t) {
                        const float comp = row[c + off[t]];
                        if (mv < comp) { mv = comp; mi = off[t]; }
                    }
                    out[cnt] = mv;
                    if (mp) mp[cnt] = ch * plane + mi + r * W + c;
                    ++cnt;
                }
            }
        }
    free(off);
}

/* pool2d.cpp:92-109 */
void orc_maxpool_backward(const float* delta, const int* mask, float* dx,
                          int B, int C, int H, int W, int OH, int OW) {
    const long in_len = (long)C * H * W, out_len = (long)C * OH * OW;
    memset(dx, 0, sizeof(float) * (size_t)B * in_len);
    for (int b = 0; b < B; ++b) {
        const int* mp = mask + b * out_len;
        const float* src = delta + b * out_len;
        float* res = dx + b * in_len;
        for (long i = 0; i < out_len; ++i) res[mp[i]] = src[i];
    }
}

/* ------------------------------------------------------------------ relu */

void orc_relu_forward(const float* x, float* y, long n) { /* relu.cpp:25 */
    for (long i = 0; i < n; ++i) y[i] = x[i] >= 0 ? x[i] : 0;
}
void orc_relu_backward(float* delta, const float* y, long n) { /* relu.cpp:39 */
    for (long i = 0; i < n; ++i) delta[i] = y[i] <= 0 ? 0 : delta[i];
}

/* ---------------------------------------------------------------- linear */

void orc_linear_forward(const float* x, const float* w, const float* bias, float* y,
                        int B, int in, int out) { /* linear.cpp:33-43 */
    for (int b = 0; b < B; ++b) {
        const float* src = x + (long)b * in;
        float* res = y + (long)b * out;
        for (int i = 0; i < out; ++i) {
            float sum = 0;
            for (int j = 0; j < in; ++j) sum += src[j] * w[(long)j * out + i];
            res[i] = sum + bias[i];
        }
    }
}

void orc_linear_backward(const float* x, const float* w, const float* delta,
                         float* dw, float* db, float* dx, int B, int in, int out) {
    for (int i = 0; i < in; ++i) { /* linear.cpp:56-64 */
        float* wp = dw + (long)i * out;
        for (int j = 0; j < out; ++j) {
            float sum = 0;
            for (int b = 0; b < B; ++b) sum += x[(long)b * in + i] * delta[(long)b * out + j];
            wp[j] = sum / B;
        }
    }
    for (int i = 0; i < out; ++i) { /* :66-71 */
        float sum = 0;
        for (int b = 0; b < B; ++b) sum += delta[(long)b * out + i];
        db[i] = sum / B;
    }
    if (!dx) return;
    for (int b = 0; b < B; ++b) { /* :80-90 */
        const float* src = delta + (long)b * out;
        float* res = dx + (long)b * in;
        for (int i = 0; i < in; ++i) {
            float sum = 0;
            const float* wp = w + (long)i * out;
            for (int j = 0; j < out; ++j) sum += src[j] * wp[j];
            res[i] = sum;
        }
    }
}

/* -------------------------------------------------------------- batchnorm */

static inline float sq(float v) { return v * v; } /* batchnorm2d.cpp:11-13 */

void orc_bn_forward_train(const float* x, const float* gamma, const float* beta,
                          float* moving_mean, float* moving_var,
                          float* batch_mean, float* batch_var,
                          float* xhat, float* y,
                          int B, int C, int H, int W, float eps, float momentum) {
    const int fl = H * W, ol = B * fl;
    for (int o = 0; o < C; ++o) {
        float u = 0; /* :46-53 */
        for (int b = 0; b < B; ++b) {
            const float* s = x + ((long)b * C + o) * fl;
            for (int i = 0; i < fl; ++i) u += s[i];
        }
        u = u / ol;
        float var = 0; /* :55-61 */
        for (int b = 0; b < B; ++b) {
            const float* s = x + ((long)b * C + o) * fl;
            for (int i = 0; i < fl; ++i) var += sq(s[i] - u);
        }
        var = var / ol;
        batch_mean[o] = u;
        batch_var[o] = var;
        const float var_inv = 1. / sqrtf(var + eps); /* :67, double divide -> float */
        for (int b = 0; b < B; ++b) {
            const float* s = x + ((long)b * C + o) * fl;
            float* n = xhat + ((long)b * C + o) * fl;
            float* d = y + ((long)b * C + o) * fl;
            for (int i = 0; i < fl; ++i) {
                n[i] = (s[i] - u) * var_inv;
                d[i] = gamma[o] * n[i] + beta[o];
            }
        }
        moving_mean[o] = (1 - momentum) * moving_mean[o] + momentum * u; /* :78-79 */
        moving_var[o] = (1 - momentum) * moving_var[o] + momentum * var;
    }
}

void orc_bn_forward_eval(const float* x, const float* gamma, const float* beta,
                         const float* moving_mean, const float* moving_var,
                         float* xhat, float* y,
                         int B, int C, int H, int W, float eps) { /* :81-94 */
    const int fl = H * W;
    for (int o = 0; o < C; ++o) {
        const float u = moving_mean[o];
        const float var_inv = 1. / sqrtf(moving_var[o] + eps);
        for (int b = 0; b < B; ++b) {
            const float* s = x + ((long)b * C + o) * fl;
            float* n = xhat + ((long)b * C + o) * fl;
            float* d = y + ((long)b * C + o) * fl;
            for (int i = 0; i < fl; ++i) {
                n[i] = (s[i] - u) * var_inv;
                d[i] = gamma[o] * n[i] + beta[o];
            }
        }
    }
}

void orc_bn_backward(float* delta, const float* x, const float* xhat,
                     const float* gamma, const float* batch_mean, const float* batch_var,
                     float* dgamma, float* dbeta,
                     int B, int C, int H, int W, float eps) {
    const int fl = H * W, ol = B * fl;
    float* ng = (float*)malloc(sizeof(float) * (size_t)ol); /* norm_gradients :109 */
    for (int o = 0; o < C; ++o) dgamma[o] = dbeta[o] = 0;
    for (int o = 0; o < C; ++o) {
        memset(ng, 0, sizeof(float) * (size_t)ol);
        for (int b = 0; b < B; ++b) { /* :118-127 */
            const float* dp = delta + ((long)b * C + o) * fl;
            const float* np = xhat + ((long)b * C + o) * fl;
            float* gp = ng + (long)b * fl;
            for (int i = 0; i < fl; ++i) {
                dgamma[o] += dp[i] * np[i];
                dbeta[o] += dp[i];
                gp[i] += dp[i] * gamma[o];
            }
        }
        float var_gradient = 0; /* :129-138: the -0.5 literal makes this a double expr */
        const float u = batch_mean[o];
        const float var_inv = 1. / sqrtf(batch_var[o] + eps);
        const float var_inv_3 = var_inv * var_inv * var_inv;
        for (int b = 0; b < B; ++b) {
            const float* s = x + ((long)b * C + o) * fl;
            const float* gp = ng + (long)b * fl;
            for (int i = 0; i < fl; ++i)
                var_gradient += gp[i] * (s[i] - u) * (-0.5) * var_inv_3;
        }
        float u_gradient = 0; /* :140-147 */
        const float inv = var_gradient / ol;
        for (int b = 0; b < B; ++b) {
            const float* s = x + ((long)b * C + o) * fl;
            const float* gp = ng + (long)b * fl;
            for (int i = 0; i < fl; ++i)
                u_gradient += gp[i] * (-var_inv) + inv * (-2) * (s[i] - u);
        }
        for (int b = 0; b < B; ++b) { /* :149-155 */
            const float* s = x + ((long)b * C + o) * fl;
            const float* gp = ng + (long)b * fl;
            float* bp = delta + ((long)b * C + o) * fl;
            for (int i = 0; i < fl; ++i)
                bp[i] = gp[i] * var_inv + inv * 2 * (s[i] - u) + u_gradient / ol;
        }
    }
    free(ng);
}

/* ------------------------------------------------------- softmax / xent */

int orc_argmax(const float* v, int n) { /* data_format.cpp:37-48 */
    float mv = v[0];
    int mi = 0;
    for (int i = 1; i < n; ++i)
        if (v[i] > mv) { mv = v[i]; mi = i; }
    return mi;
}

static inline float clamped_exp(float v) { /* func.cpp:7-11 */
    if (v >= 88) return FLT_MAX;
    else if (v <= -50) return 0.f;
    return expf(v);
}

void orc_softmax(const float* logits, float* probs, int B, int n) { /* func.cpp:16-37 */
    for (int b = 0; b < B; ++b) {
        const float* in = logits + (long)b * n;
        float* p = probs + (long)b * n;
        const float mv = in[orc_argmax(in, n)];
        float sum = 0;
        for (int i = 0; i < n; ++i) {
            p[i] = clamped_exp(in[i] - mv);
            sum += p[i];
        }
        for (int i = 0; i < n; ++i) p[i] /= sum;
        for (int i = 0; i < n; ++i)
            if (isnan(p[i])) p[i] = 0.f;
    }
}

float orc_cross_entropy_backward(const float* probs, const int* labels, float* delta,
                                 int B, int n) { /* func.cpp:40-73 */
    float loss = 0;
    for (int b = 0; b < B; ++b)
        for (int i = 0; i < n; ++i) {
            const float yv = (labels[b] == i) ? 1.0f : 0.0f;
            delta[(long)b * n + i] = probs[(long)b * n + i] - yv;
            loss += logf(probs[(long)b * n + i]) * yv; /* 0*log(0) = NaN, as in the reference */
        }
    loss = loss * (-1.0) / B;
    return loss;
}

void orc_sgd(float* p, const float* g, long n, float lr) {
    for (long i = 0; i < n; ++i) p[i] -= lr * g[i];
}

/* --------------------------------------------------------- extensions
 * NOT in the reference: items 7-8 of its TODO list (cnn.cpp:15-24, "padding", "AvgPool / global pool").  The
 * definitions the CUDA kernels are checked against: zero padding as a layer in front of a convolution, and the
 * window mean (fp32 sum in scan order, then one division by k*k).  Parity for these two is against this
 * definition only ("parity unpinned": the reference has nothing to pin them to). */
void orc_pad_forward(const float* x, float* y, int B, int C, int H, int W, int pad) {
    const int PH = H + 2 * pad, PW = W + 2 * pad;
    memset(y, 0, sizeof(float) * (size_t)B * C * PH * PW);
    for (long bc = 0; bc < (long)B * C; ++bc)
        for (int r = 0; r < H; ++r)
            memcpy(y + (bc * PH + r + pad) * PW + pad, x + (bc * H + r) * W, sizeof(float) * (size_t)W);
}
void orc_pad_backward(const float* delta, float* dx, int B, int C, int H, int W, int pad) {
    const int PH = H + 2 * pad, PW = W + 2 * pad;
    for (long bc = 0; bc < (long)B * C; ++bc)
        for (int r = 0; r < H; ++r)
            memcpy(dx + (bc * H + r) * W, delta + (bc * PH + r + pad) * PW + pad, sizeof(float) * (size_t)W);
}
void orc_avgpool_forward(const float* x, float* y, int B, int C, int H, int W, int k, int step) {
    const int OH = (H - k) / step + 1, OW = (W - k) / step + 1;
    for (long bc = 0; bc < (long)B * C; ++bc)
        for (int oy = 0; oy < OH; ++oy)
            for (int ox = 0; ox < OW; ++ox) {
                float s = 0.f;
                for (int a = 0; a < k; ++a)
                    for (int b = 0; b < k; ++b) s += x[(bc * H + oy * step + a) * W + ox * step + b];
                y[(bc * OH + oy) * OW + ox] = s / (float)(k * k);
            }
}
/* every input cell gathers the deltas of the windows that cover it (highest window first), then one multiply */
void orc_avgpool_backward(const float* delta, float* dx, int B, int C, int H, int W, int k, int step) {
    const int OH = (H - k) / step + 1, OW = (W - k) / step + 1;
    const float inv = 1.f / (float)(k * k);
    for (long bc = 0; bc < (long)B * C; ++bc)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                float s = 0.f;
                int oy1 = y / step, ox1 = x / step;
                if (oy1 > OH - 1) oy1 = OH - 1;
                if (ox1 > OW - 1) ox1 = OW - 1;
                for (int oy = oy1; oy >= 0 && oy * step + k > y; --oy)
                    for (int ox = ox1; ox >= 0 && ox * step + k > x; --ox) s += delta[(bc * OH + oy) * OW + ox];
                dx[(bc * H + y) * W + x] = s * inv;
            }
}

/* ----------------------------------------------------------------- net */

typedef struct {
    orc_layer_spec spec;
    int C, H, W;       /* input shape */
    int OC, OH, OW;    /* output shape */
    float *w, *b, *dw, *db;          /* conv / linear / bn(gamma,beta) */
    float *mm, *mv, *bm, *bv, *xhat; /* bn */
    int* mask;                       /* pool */
    const float* in;                 /* saved input (aliases previous output) */
    float* out;                      /* persistent output [B][OC][OH][OW] */
    float* dxbuf;                    /* delta_output (conv, pool, linear) */
} orc_layer;

struct orc_net {
    int n, B;
    orc_layer* L;
    float *probs, *delta0;
    int classes;
};

static long lay_wcount(const orc_layer* l) {
    switch (l->spec.type) {
        case ORC_CONV: return (long)l->spec.b * l->spec.a * l->spec.c * l->spec.c;
        case ORC_LINEAR: return (long)l->spec.a * l->spec.b;
        case ORC_BN: return l->spec.a;
        default: return 0;
    }
}
static long lay_bcount(const orc_layer* l) {
    switch (l->spec.type) {
        case ORC_CONV: return l->spec.b;
        case ORC_LINEAR: return l->spec.b;
        case ORC_BN: return l->spec.a;
        default: return 0;
    }
}

orc_net* orc_net_create(const orc_layer_spec* specs, int n_layers, int B, int C, int H, int W) {
    orc_net* net = (orc_net*)calloc(1, sizeof(orc_net));
    net->n = n_layers;
    net->B = B;
    net->L = (orc_layer*)calloc((size_t)n_layers, sizeof(orc_layer));
    for (int i = 0; i < n_layers; ++i) {
        orc_layer* l = &net->L[i];
        l->spec = specs[i];
        l->C = C; l->H = H; l->W = W;
        const long in_len = (long)B * C * H * W;
        switch (l->spec.type) {
            case ORC_CONV:
                l->OC = l->spec.b;
                l->OH = (H - l->spec.c) / l->spec.d + 1;
                l->OW = (W - l->spec.c) / l->spec.d + 1;
                l->dxbuf = (float*)calloc((size_t)in_len, sizeof(float));
                break;
            case ORC_POOL:
                l->OC = C;
                l->OH = (H - l->spec.a) / l->spec.b + 1;
                l->OW = (W - l->spec.a) / l->spec.b + 1;
                l->dxbuf = (float*)calloc((size_t)in_len, sizeof(float));
                l->mask = (int*)calloc((size_t)B * l->OC * l->OH * l->OW, sizeof(int));
                break;
            case ORC_LINEAR:
                l->OC = l->spec.b; l->OH = 1; l->OW = 1;
                l->dxbuf = (float*)calloc((size_t)in_len, sizeof(float));
                break;
            case ORC_PAD:
                l->OC = C; l->OH = H + 2 * l->spec.a; l->OW = W + 2 * l->spec.a;
                l->dxbuf = (float*)calloc((size_t)in_len, sizeof(float));
                break;
            case ORC_AVGPOOL:
                l->OC = C;
                l->OH = (H - l->spec.a) / l->spec.b + 1;
                l->OW = (W - l->spec.a) / l->spec.b + 1;
                l->dxbuf = (float*)calloc((size_t)in_len, sizeof(float));
                break;
            case ORC_BN:
                l->OC = C; l->OH = H; l->OW = W;
                l->mm = (float*)calloc((size_t)C, sizeof(float));
                l->mv = (float*)calloc((size_t)C, sizeof(float));
                l->bm = (float*)calloc((size_t)C, sizeof(float));
                l->bv = (float*)calloc((size_t)C, sizeof(float));
                l->xhat = (float*)calloc((size_t)in_len, sizeof(float));
                break;
            default: /* RELU */
                l->OC = C; l->OH = H; l->OW = W;
        }
        const long wc = lay_wcount(l), bc = lay_bcount(l);
        if (wc) {
            l->w = (float*)calloc((size_t)wc, sizeof(float));
            l->dw = (float*)calloc((size_t)wc, sizeof(float));
            l->b = (float*)calloc((size_t)bc, sizeof(float));
            l->db = (float*)calloc((size_t)bc, sizeof(float));
        }
        l->out = (float*)calloc((size_t)B * l->OC * l->OH * l->OW, sizeof(float));
        C = l->OC; H = l->OH; W = l->OW;
    }
    net->classes = C * H * W;
    net->probs = (float*)calloc((size_t)B * net->classes, sizeof(float));
    net->delta0 = (float*)calloc((size_t)B * net->classes, sizeof(float));
    return net;
}

void orc_net_destroy(orc_net* net) {
    if (!net) return;
    for (int i = 0; i < net->n; ++i) {
        orc_layer* l = &net->L[i];
        free(l->w); free(l->b); free(l->dw); free(l->db);
        free(l->mm); free(l->mv); free(l->bm); free(l->bv); free(l->xhat);
        free(l->mask); free(l->out); free(l->dxbuf);
    }
    free(net->L); free(net->probs); free(net->delta0); free(net);
}

long orc_net_param_count(const orc_net* net) {
    long n = 0;
    for (int i = 0; i < net->n; ++i) {
        const orc_layer* l = &net->L[i];
        n += lay_wcount(l) + lay_bcount(l);
        if (l->spec.type == ORC_BN) n += 2L * l->spec.a;
    }
    return n;
}
int orc_net_num_classes(const orc_net* net) { return net->classes; }

/* Checkpoint order (alexnet.cpp:69-77): conv W then bias (conv2d.cpp:220-226),
 * linear W then bias (linear.cpp:105-108), BN gamma,beta,moving_mean,moving_var
 * (batchnorm2d.cpp:168-174). */
void orc_net_set_params(orc_net* net, const float* f) {
    for (int i = 0; i < net->n; ++i) {
        orc_layer* l = &net->L[i];
        const long wc = lay_wcount(l), bc = lay_bcount(l);
        if (!wc) continue;
        memcpy(l->w, f, sizeof(float) * (size_t)wc); f += wc;
        memcpy(l->b, f, sizeof(float) * (size_t)bc); f += bc;
        if (l->spec.type == ORC_BN) {
            memcpy(l->mm, f, sizeof(float) * (size_t)bc); f += bc;
            memcpy(l->mv, f, sizeof(float) * (size_t)bc); f += bc;
        }
    }
}
void orc_net_get_params(const orc_net* net, float* f) {
    for (int i = 0; i < net->n; ++i) {
        const orc_layer* l = &net->L[i];
        const long wc = lay_wcount(l), bc = lay_bcount(l);
        if (!wc) continue;
        memcpy(f, l->w, sizeof(float) * (size_t)wc); f += wc;
        memcpy(f, l->b, sizeof(float) * (size_t)bc); f += bc;
        if (l->spec.type == ORC_BN) {
            memcpy(f, l->mm, sizeof(float) * (size_t)bc); f += bc;
            memcpy(f, l->mv, sizeof(float) * (size_t)bc); f += bc;
        }
    }
}
void orc_net_get_grads(const orc_net* net, float* f) {
    for (int i = 0; i < net->n; ++i) {
        const orc_layer* l = &net->L[i];
        const long wc = lay_wcount(l), bc = lay_bcount(l);
        if (!wc) continue;
        memcpy(f, l->dw, sizeof(float) * (size_t)wc); f += wc;
        memcpy(f, l->db, sizeof(float) * (size_t)bc); f += bc;
        if (l->spec.type == ORC_BN) {
            memset(f, 0, sizeof(float) * 2 * (size_t)bc); f += 2 * bc;
        }
    }
}

const float* orc_net_layer_output(const orc_net* net, int layer, long* count) {
    const orc_layer* l = &net->L[layer];
    if (count) *count = (long)net->B * l->OC * l->OH * l->OW;
    return l->out;
}

void orc_net_forward(orc_net* net, const float* x, float* logits, int no_grad) {
    const int B = net->B;
    const float* cur = x;
    for (int i = 0; i < net->n; ++i) { /* alexnet.cpp:41-44 */
        orc_layer* l = &net->L[i];
        l->in = cur;
        const orc_layer_spec* s = &l->spec;
        switch (s->type) {
            case ORC_CONV:
                orc_conv2d_forward(cur, l->w, l->b, l->out, B, l->C, l->H, l->W, s->b, s->c, s->d);
                break;
            case ORC_BN:
                if (!no_grad)
                    orc_bn_forward_train(cur, l->w, l->b, l->mm, l->mv, l->bm, l->bv, l->xhat,
                                         l->out, B, l->C, l->H, l->W, 1e-5, 0.1);
                else
                    orc_bn_forward_eval(cur, l->w, l->b, l->mm, l->mv, l->xhat, l->out, B, l->C,
                                        l->H, l->W, 1e-5);
                break;
            case ORC_RELU:
                orc_relu_forward(cur, l->out, (long)B * l->C * l->H * l->W);
                break;
            case ORC_POOL:
                orc_maxpool_forward(cur, l->out, no_grad ? NULL : l->mask, B, l->C, l->H, l->W,
                                    s->a, s->b);
                break;
            case ORC_LINEAR:
                orc_linear_forward(cur, l->w, l->b, l->out, B, s->a, s->b);
                break;
            case ORC_PAD:
                orc_pad_forward(cur, l->out, B, l->C, l->H, l->W, s->a);
                break;
            case ORC_AVGPOOL:
                orc_avgpool_forward(cur, l->out, B, l->C, l->H, l->W, s->a, s->b);
                break;
        }
        cur = l->out;
    }
    if (logits) memcpy(logits, cur, sizeof(float) * (size_t)B * net->classes);
}

/* cnn.cpp:81-90: forward, softmax, cross_entroy_backward, backward, update. */
float orc_net_train_step(orc_net* net, const float* x, const int* labels, float lr,
                         float* probs, float* dx_image) {
    const int B = net->B;
    orc_net_forward(net, x, NULL, 0);
    const float* logits = net->L[net->n - 1].out;
    orc_softmax(logits, net->probs, B, net->classes);
    const float loss = orc_cross_entropy_backward(net->probs, labels, net->delta0, B, net->classes);
    if (probs) memcpy(probs, net->probs, sizeof(float) * (size_t)B * net->classes);
    float* delta = net->delta0;
    for (int i = net->n - 1; i >= 0; --i) { /* alexnet.cpp:53-58 */
        orc_layer* l = &net->L[i];
        const orc_layer_spec* s = &l->spec;
        switch (s->type) {
            case ORC_CONV:
                orc_conv2d_backward(l->in, l->w, delta, l->dw, l->db, l->dxbuf, B, l->C, l->H,
                                    l->W, s->b, s->c, s->d);
                delta = l->dxbuf;
                break;
            case ORC_BN:
                orc_bn_backward(delta, l->in, l->xhat, l->w, l->bm, l->bv, l->dw, l->db, B, l->C,
                                l->H, l->W, 1e-5);
                break;
            case ORC_RELU:
                orc_relu_backward(delta, l->out, (long)B * l->C * l->H * l->W);
                break;
            case ORC_POOL:
                orc_maxpool_backward(delta, l->mask, l->dxbuf, B, l->C, l->H, l->W, l->OH, l->OW);
                delta = l->dxbuf;
                break;
            case ORC_LINEAR:
                orc_linear_backward(l->in, l->w, delta, l->dw, l->db, l->dxbuf, B, s->a, s->b);
                delta = l->dxbuf;
                break;
            case ORC_PAD:
                orc_pad_backward(delta, l->dxbuf, B, l->C, l->H, l->W, s->a);
                delta = l->dxbuf;
                break;
            case ORC_AVGPOOL:
                orc_avgpool_backward(delta, l->dxbuf, B, l->C, l->H, l->W, s->a, s->b);
                delta = l->dxbuf;
                break;
        }
    }
    if (dx_image)
        memcpy(dx_image, delta,
               sizeof(float) * (size_t)B * net->L[0].C * net->L[0].H * net->L[0].W);
    for (int i = 0; i < net->n; ++i) { /* alexnet.cpp:62-65 */
        orc_layer* l = &net->L[i];
        const long wc = lay_wcount(l), bc = lay_bcount(l);
        if (!wc) continue;
        orc_sgd(l->w, l->dw, wc, lr);
        orc_sgd(l->b, l->db, bc, lr);
    }
    return loss;
}
