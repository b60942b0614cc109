"""Pins the plain-C oracle against oracle/_ref (the reference's own sources compiled in place)
on fresh seeded inputs -- bit for bit.  Skipped where libcnn_ref.so was never built."""
import numpy as np
import pytest

from oracle import port, ref
from cnn_b200.nets import alexnet_lite, insert_bn_params

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")


def eq(a, b):
    return np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("cfg", [(2, 3, 31, 29, 8, 3, 2), (1, 6, 16, 16, 12, 3, 1), (3, 2, 17, 19, 5, 5, 3),
                                 (1, 16, 27, 27, 32, 3, 2), (2, 1, 7, 7, 1, 7, 1)])
def test_conv(cfg):
    B, Cin, H, W, Cout, k, s = cfg
    rng = np.random.default_rng(sum(cfg))
    x = rng.random((B, Cin, H, W), dtype=np.float32)
    w = (rng.standard_normal((Cout, Cin, k, k)) / 10).astype(np.float32)
    b = (rng.standard_normal(Cout) / 10).astype(np.float32)
    y = port.conv2d_forward(x, w, b, s)
    d = rng.standard_normal(y.shape).astype(np.float32)
    yr, dwr, dbr, dxr = ref.conv2d(x, w, b, s, d)
    dw, db, dx = port.conv2d_backward(x, w, d, s)
    assert eq(y, yr) and eq(dw, dwr) and eq(db, dbr) and eq(dx, dxr)


@pytest.mark.parametrize("cfg", [(2, 5, 111, 111, 2, 2), (1, 3, 13, 10, 3, 2), (2, 2, 9, 9, 3, 1), (1, 1, 6, 6, 2, 3)])
def test_pool(cfg):
    B, C, H, W, k, st = cfg
    rng = np.random.default_rng(sum(cfg))
    x = (np.round(rng.standard_normal((B, C, H, W)) * 3) / 3).astype(np.float32)
    y, mask = port.maxpool_forward(x, k, st)
    d = rng.standard_normal(y.shape).astype(np.float32)
    yr, maskr, dxr = ref.maxpool(x, k, st, d)
    assert eq(y, yr) and np.array_equal(mask, maskr)
    assert eq(port.maxpool_backward(d, mask, x.shape), dxr)


def test_relu_linear_bn_xent():
    rng = np.random.default_rng(7)
    x = rng.standard_normal((3, 4, 6, 5)).astype(np.float32)
    d = rng.standard_normal(x.shape).astype(np.float32)
    yr, dr = ref.relu(x, d)
    y = port.relu_forward(x)
    assert eq(y, yr) and eq(port.relu_backward(d, y), dr)

    w = (rng.standard_normal((120, 7)) / 10).astype(np.float32)
    b = (rng.standard_normal(7) / 10).astype(np.float32)
    dl = rng.standard_normal((3, 7)).astype(np.float32)
    yr, dwr, dbr, dxr = ref.linear(x, w, b, dl)
    xf = x.reshape(3, -1)
    dw, db, dx = port.linear_backward(xf, w, dl)
    assert eq(port.linear_forward(xf, w, b), yr) and eq(dw, dwr) and eq(db, dbr) and eq(dx.reshape(x.shape), dxr)

    g = (1 + rng.standard_normal(4) / 3).astype(np.float32)
    bt = rng.standard_normal(4).astype(np.float32)
    z4 = np.zeros(4, np.float32)
    r = ref.batchnorm(x, g, bt, z4, z4, d)
    p = port.bn_forward_train(x, g, bt, z4, z4)
    assert all(eq(p[k], r[k]) for k in ("y", "xhat", "mean", "var", "moving_mean", "moving_var"))
    dx, dg, dbt = port.bn_backward(d, x, p["xhat"], g, p["mean"], p["var"])
    assert eq(dx, r["dx"]) and eq(dg, r["dgamma"]) and eq(dbt, r["dbeta"])

    z = (rng.standard_normal((16, 10)) * 20).astype(np.float32)
    lab = rng.integers(0, 10, 16).astype(np.int32)
    pr, predr, lossr, dlr = ref.softmax_xent(z, lab)
    pp = port.softmax(z)
    loss, dd = port.cross_entropy_backward(pp, lab)
    assert eq(pp, pr) and np.array_equal(port.argmax_rows(pp), predr) and eq(dd, dlr)
    assert eq(np.float32(loss), np.float32(lossr))


@pytest.mark.parametrize("bn", [False, True])
def test_alexnet_container_train_steps(bn):
    """The reference AlexNet container (alexnet.cpp) vs the oracle's net runner, B=2, 2 steps."""
    rng = np.random.default_rng(11)
    B = 2
    x = rng.random((B, 3, 224, 224), dtype=np.float32)
    lab = np.array([2, 0], np.int32)
    init = ref.alexnet_init_params(3, bn)
    spec = alexnet_lite(3, batch_norm=bn)
    if bn:  # reference init of a BN net == no-BN draws + constructor-state BN params
        assert eq(init, insert_bn_params(spec, ref.alexnet_init_params(3, False)))
    rnet = ref.Net(batch_norm=bn)
    rnet.set_params(init)
    onet = port.Net(spec, B, 3, 224, 224)
    onet.set_params(init)
    for _ in range(2):
        lr_, pr, dxr = rnet.train_step(x, lab, 1e-3, want_dx=True)
        lo, po, dxo = onet.train_step(x, lab, 1e-3, want_dx=True)
        assert eq(np.float32(lo), np.float32(lr_)) and eq(po, pr) and eq(dxo, dxr)
        assert eq(onet.get_grads(), rnet.get_grads())
    assert eq(onet.get_params(), rnet.get_params())
    # eval-mode forward (WithoutGrad): BN uses moving statistics
    assert eq(onet.forward(x, no_grad=True), rnet.forward(x, 3, no_grad=True))
