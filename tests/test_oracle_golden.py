"""The plain-C oracle (oracle/cnn_oracle.c) against the committed golden fixtures, which were
produced by the reference's own code (tests/golden/make_golden.py).  Bit-exact everywhere:
the oracle restates the reference loop for loop.  Also the README inference known-answer."""
import os

import numpy as np

from conftest import GOLDEN
from oracle import port
from cnn_b200.nets import alexnet_lite
from cnn_b200.synth import synth_images, synth_labels


def eq(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


def test_conv_golden(ops_golden):
    g = ops_golden
    for tag in ("conv_a", "conv_b", "conv_c", "conv_d", "conv_e"):
        s = int(g[f"{tag}.cfg"][6])
        y = port.conv2d_forward(g[f"{tag}.x"], g[f"{tag}.w"], g[f"{tag}.b"], s)
        assert eq(y, g[f"{tag}.y"]), tag
        dw, db, dx = port.conv2d_backward(g[f"{tag}.x"], g[f"{tag}.w"], g[f"{tag}.delta"], s)
        assert eq(dw, g[f"{tag}.dw"]) and eq(db, g[f"{tag}.db"]) and eq(dx, g[f"{tag}.dx"]), tag


def test_dgrad_leaves_uncovered_border_zero(ops_golden):
    dx = ops_golden["conv_a.dx"]  # H=20,W=22,k=3,s=2 -> last row / col never touched
    assert not dx[:, :, -1, :].any() and not dx[:, :, :, -1].any()


def test_pool_golden(ops_golden):
    g = ops_golden
    for tag in ("pool_a", "pool_b", "pool_c"):
        _, _, _, _, k, st = (int(v) for v in g[f"{tag}.cfg"])
        y, mask = port.maxpool_forward(g[f"{tag}.x"], k, st)
        assert eq(y, g[f"{tag}.y"]) and np.array_equal(mask, g[f"{tag}.mask"]), tag
        dx = port.maxpool_backward(g[f"{tag}.delta"], mask, g[f"{tag}.x"].shape)
        assert eq(dx, g[f"{tag}.dx"]), tag


def test_relu_golden(ops_golden):
    g = ops_golden
    y = port.relu_forward(g["relu.x"])
    assert eq(y, g["relu.y"])
    assert np.signbit(y.flat[1])  # -0.0 passes through (relu.cpp:25: x >= 0)
    assert y.flat[2] == 0         # NaN -> 0
    assert eq(port.relu_backward(g["relu.delta"], y), g["relu.dx"])


def test_linear_golden(ops_golden):
    g = ops_golden
    x = g["linear.x"].reshape(3, -1)
    assert eq(port.linear_forward(x, g["linear.w"], g["linear.b"]), g["linear.y"])
    dw, db, dx = port.linear_backward(x, g["linear.w"], g["linear.delta"])
    assert eq(dw, g["linear.dw"]) and eq(db, g["linear.db"])
    assert eq(dx.reshape(g["linear.dx"].shape), g["linear.dx"])


def test_bn_golden(ops_golden):
    g = ops_golden
    r = port.bn_forward_train(g["bn.x"], g["bn.gamma"], g["bn.beta"], g["bn.mm"], g["bn.mv"])
    for k_, gk in (("y", "bn.y"), ("xhat", "bn.xhat"), ("mean", "bn.mean"), ("var", "bn.var"),
                   ("moving_mean", "bn.mm_out"), ("moving_var", "bn.mv_out")):
        assert eq(r[k_], g[gk]), k_
    dx, dg, db = port.bn_backward(g["bn.delta"], g["bn.x"], r["xhat"], g["bn.gamma"], r["mean"], r["var"])
    assert eq(dx, g["bn.dx"]) and eq(dg, g["bn.dgamma"]) and eq(db, g["bn.dbeta"])
    e = port.bn_forward_eval(g["bn.x"], g["bn.gamma"], g["bn.beta"], r["moving_mean"], r["moving_var"])
    assert eq(e["y"], g["bn.y_eval"])


def test_softmax_xent_golden(ops_golden):
    g = ops_golden
    for t in ("xent", "xent2"):
        p = port.softmax(g[f"{t}.z"])
        assert eq(p, g[f"{t}.p"])
        assert np.array_equal(port.argmax_rows(p), g[f"{t}.pred"])
        loss, d = port.cross_entropy_backward(p, g[f"{t}.labels"])
        assert eq(d, g[f"{t}.delta"])
        assert eq(np.float32(loss), g[f"{t}.loss"])
    assert np.isnan(g["xent.loss"])  # the reference's 0*log(0) quirk is reproduced


def test_readme_inference_known_answer():
    """imgs/image-20230208213627060.png: dog 0.850634, panda 0.999978, bird 0.999998
    (inference.cpp:35,55-70 with checkpoint AlexNet_aug_1e-3/iter_395000)."""
    params = np.fromfile(os.path.join(GOLDEN, "kat_checkpoint.model"), np.float32)
    u8 = np.load(os.path.join(GOLDEN, "kat_images_u8.npy"))          # [3,224,224,3] BGR HWC
    x = (u8.astype(np.float32) * np.float32(1.0) / np.float32(255)).transpose(0, 3, 1, 2)  # data_format.cpp:17-21
    net = port.Net(alexnet_lite(3), 1, 3, 224, 224)
    net.set_params(params)
    want = [(0, "0.850634"), (1, "0.999978"), (2, "0.999998")]
    for i, (cls, txt) in enumerate(want):
        p = port.softmax(net.forward(x[i:i + 1], no_grad=True))[0]
        assert int(np.argmax(p)) == cls
        assert f"{p[cls]:.6g}" == txt.rstrip("0") or f"{p[cls]:.6f}" == txt, (i, p)


def test_train_trajectory_golden(train_golden):
    g = train_golden
    init = np.fromfile(os.path.join(GOLDEN, "alexnet_init.model"), np.float32)
    # libstdc++ draws of the reference constructors (conv2d.cpp:23-30): first conv1 weights
    assert np.allclose(init[:3], [-0.0405017957, 0.0594193935, 0.151970685], rtol=0, atol=1e-9)
    B = 4
    x = synth_images(B, 3, 224, 224, seed=1234)
    lab = synth_labels(B, 3)
    assert np.array_equal(lab, g["labels"])
    net = port.Net(alexnet_lite(3), B, 3, 224, 224)
    assert net.n_params == 111267 == init.size
    net.set_params(init)
    for step in range(3):
        loss, probs, dx = net.train_step(x, lab, 1e-3, want_dx=True)
        assert eq(np.float32(loss), g[f"loss{step}"]) and eq(probs, g[f"probs{step}"])
        if step == 0:
            assert eq(net.get_grads(), g["grads0"])
            assert eq(dx[:, :, ::7, ::5], g["dx_image0_sample"])
            for li in (0, 2, 3, 9):
                assert eq(net.layer_output(li)[::97], g[f"layer{li}_out_sample"])
    assert eq(net.get_params(), g["params3"])


def test_train_trajectory_bn_golden(train_golden):
    g = train_golden
    B = 4
    x = synth_images(B, 3, 224, 224, seed=1234)
    lab = synth_labels(B, 3)
    net = port.Net(alexnet_lite(3, batch_norm=True), B, 3, 224, 224)
    init = np.fromfile(os.path.join(GOLDEN, "alexnet_init.model"), np.float32)
    from cnn_b200.nets import insert_bn_params
    net.set_params(insert_bn_params(alexnet_lite(3, batch_norm=True), init))
    for step in range(2):
        loss, probs, _ = net.train_step(x, lab, 1e-3)
        assert eq(np.float32(loss), g[f"bn_loss{step}"]) and eq(probs, g[f"bn_probs{step}"])
    assert eq(net.get_params()[::13], g["bn_params2_sample"])


def test_extension_layers_against_torch_autograd():
    """PAD / AVGPOOL are not in the reference (its TODO items 7-8, cnn.cpp:15-24), so the oracle's definition of them is
    pinned to an independent one: torch's fp64 pad / avg_pool2d and their autograd gradients, plus one train step of a
    small 'same'-padded net against torch autograd of the same graph (fp64)."""
    import torch
    import torch.nn.functional as F
    from oracle import port
    rng = np.random.default_rng(3)
    for (B, C, H, W, k, step, pad) in [(2, 3, 9, 8, 3, 2, 1), (1, 2, 6, 6, 6, 6, 2), (2, 2, 7, 7, 3, 1, 0)]:
        x = rng.standard_normal((B, C, H, W)).astype(np.float32)
        xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
        yp = F.pad(xt, (pad, pad, pad, pad))
        np.testing.assert_array_equal(port.pad_forward(x, pad), yp.detach().numpy().astype(np.float32))
        dp = rng.standard_normal(tuple(yp.shape)).astype(np.float32)
        yp.backward(torch.tensor(dp, dtype=torch.float64))
        np.testing.assert_array_equal(port.pad_backward(dp, x.shape, pad), xt.grad.numpy().astype(np.float32))
        xt.grad = None
        ya = F.avg_pool2d(xt, k, step)
        assert np.abs(port.avgpool_forward(x, k, step) - ya.detach().numpy()).max() <= 1e-6
        da = rng.standard_normal(tuple(ya.shape)).astype(np.float32)
        ya.backward(torch.tensor(da, dtype=torch.float64))
        assert np.abs(port.avgpool_backward(da, x.shape, k, step) - xt.grad.numpy()).max() <= 1e-6
    # whole step: PAD 1 -> conv 3x3 s1 -> ReLU -> AVGPOOL (global) -> Linear -> softmax-xent, batch-mean gradients
    spec = [(5, 1, 0, 0, 0), (0, 3, 4, 3, 1), (2, 0, 0, 0, 0), (6, 8, 8, 0, 0), (4, 4, 3, 0, 0)]
    B = 3
    x = rng.random((B, 3, 8, 8)).astype(np.float32)
    lab = np.array([0, 2, 1], np.int32)
    w1 = (rng.standard_normal((4, 3, 3, 3)) / 3).astype(np.float32)
    b1 = (rng.standard_normal(4) / 10).astype(np.float32)
    w2 = rng.standard_normal((4, 3)).astype(np.float32)     # LinearLayer weights are [in][out] (linear.cpp:33-45)
    b2 = (rng.standard_normal(3) / 10).astype(np.float32)
    o = port.Net(spec, B, 3, 8, 8)
    o.set_params(np.concatenate([w1.ravel(), b1, w2.ravel(), b2]))
    loss, probs, dx = o.train_step(x, lab, 0.0, want_dx=True)
    T = lambda a: torch.tensor(a, dtype=torch.float64, requires_grad=True)
    xt, W1, B1, W2, B2 = T(x), T(w1), T(b1), T(w2), T(b2)
    h = F.avg_pool2d(F.relu(F.conv2d(F.pad(xt, (1, 1, 1, 1)), W1, B1)), 8).flatten(1)
    logits = h @ W2 + B2
    logp = F.log_softmax(logits, dim=1)
    # cross_entroy_backward (func.cpp:56-73) returns delta = probs - one_hot per image; the layers divide by the batch
    tl = -logp[torch.arange(B), torch.tensor(lab, dtype=torch.long)].sum() / B
    tl.backward()
    g = o.get_grads()
    ref = np.concatenate([W1.grad.numpy().ravel(), B1.grad.numpy(), W2.grad.numpy().ravel(), B2.grad.numpy()])
    assert np.abs(probs - logp.exp().detach().numpy()).max() <= 1e-6
    assert np.abs(g - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max())
    # the image gradient is per image (not batch-averaged): the layers only average the weight gradients
    assert np.abs(dx - B * xt.grad.numpy()).max() <= 1e-5 * max(1.0, np.abs(dx).max())


def test_extension_net_spec_shapes_and_layout():
    """nets.padded_resnet_shaped: 'same' convolutions keep the resolution, the global average pool leaves one value per
    channel, and the oracle engine agrees with nets.shapes / nets.param_layout on sizes."""
    from cnn_b200 import nets
    from oracle import port
    spec = nets.padded_resnet_shaped(3, width=8, in_hw=16)
    shp = nets.shapes(spec, 3, 16, 16)
    assert shp[0] == (3, 18, 18) and shp[1] == (8, 16, 16)          # PAD 1, then 3x3 stride-1: back to 16x16
    assert shp[-2] == (16, 1, 1) and shp[-1] == (3, 1, 1)            # global average pool, classifier
    _, total = nets.param_layout(spec)
    o = port.Net(spec, 2, 3, 16, 16)
    assert o.n_params == total and o.classes == 3
    for li, (c, h, w) in enumerate(shp):
        assert o.layer_output(li).size == 2 * c * h * w, li
