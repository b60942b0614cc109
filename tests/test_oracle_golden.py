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
