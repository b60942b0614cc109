#!/usr/bin/env python
"""Generates the committed golden fixtures in tests/golden/ by RUNNING THE REFERENCE ITSELF
(oracle/_ref/libcnn_ref.so = /root/reference/cpu/src compiled in place) in the build
container.  The GPU box has no /root/reference, so tests there use these files.

    python tests/golden/make_golden.py        # needs /root/reference and cv2

Outputs (all inputs are seeded, so the run is reproducible):
  kat_checkpoint.model    copy of cpu/checkpoints/AlexNet_aug_1e-3/iter_395000_train_0.918_valid_0.913.model
  kat_images_u8.npy       dog/panda/bird.jpg decoded by cv2.imread and cv2.resize(224,224) (uint8 HWC, BGR)
  gradcam_kat.npz         AlexNet::grad_cam("conv_layer_3") of those images by the reference (6x6 u8 maps, probabilities)
                          -- exactly what inference.cpp:55-63 feeds read_from_opencv_mat
  alexnet_init.model      parameters drawn by the reference constructors (seeds 212 / 1998)
  ops_golden.npz          per-layer forward/backward vectors from the reference classes
  train_golden.npz        3 AlexNet-lite train steps (B=4) from alexnet_init.model on synthetic input
"""
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref  # noqa: E402
from cnn_b200.synth import synth_images, synth_labels  # noqa: E402

REF = "/root/reference"


def kat():
    import cv2
    ck = os.path.join(REF, "cpu/checkpoints/AlexNet_aug_1e-3/iter_395000_train_0.918_valid_0.913.model")
    shutil.copyfile(ck, os.path.join(HERE, "kat_checkpoint.model"))
    imgs = []
    for n in ("dog", "panda", "bird"):
        im = cv2.imread(os.path.join(REF, f"datasets/images/{n}.jpg"))
        imgs.append(cv2.resize(im, (224, 224)))
    np.save(os.path.join(HERE, "kat_images_u8.npy"), np.stack(imgs).astype(np.uint8))


def gradcam():
    """grad_cam.cpp:71-80 / AlexNet::grad_cam (alexnet.cpp:95-142) run by the reference itself on the three README
    images with the shipped checkpoint: the 6x6 8-bit class-activation maps + probabilities."""
    u8 = np.load(os.path.join(HERE, "kat_images_u8.npy"))
    x = np.ascontiguousarray((u8.astype(np.float32) * np.float32(1.0) / np.float32(255)).transpose(0, 3, 1, 2))
    net = ref.Net()
    net.load_file(os.path.join(HERE, "kat_checkpoint.model"))
    cams, probs = [], []
    for i in range(3):
        c, p = net.grad_cam(x[i])
        cams.append(c)
        probs.append(p)
    np.savez_compressed(os.path.join(HERE, "gradcam_kat.npz"), cam=np.stack(cams), probs=np.stack(probs))


def ops():
    rng = np.random.default_rng(20260101)
    out = {}

    def rnd(*s):
        return rng.standard_normal(s).astype(np.float32)

    # conv: (B, Cin, H, W, Cout, k, stride)
    for tag, (B, Cin, H, W, Cout, k, s) in {
        "conv_a": (2, 3, 20, 22, 16, 3, 2),    # AlexNet-lite conv1 shape class (even H -> uncovered border)
        "conv_b": (2, 16, 13, 13, 32, 3, 2),
        "conv_c": (1, 8, 12, 11, 16, 3, 1),    # VGG-style stride 1
        "conv_d": (2, 4, 15, 14, 8, 5, 2),     # k=5
        "conv_e": (1, 5, 9, 9, 7, 3, 3),       # odd channel counts, stride 3
    }.items():
        x = rng.random((B, Cin, H, W), dtype=np.float32)
        w = (rnd(Cout, Cin, k, k) / 10).astype(np.float32)
        b = (rnd(Cout) / 10).astype(np.float32)
        y = ref.conv2d(x, w, b, s)
        d = rnd(*y.shape)
        y, dw, db, dx = ref.conv2d(x, w, b, s, d)
        out.update({f"{tag}.cfg": np.array([B, Cin, H, W, Cout, k, s]), f"{tag}.x": x, f"{tag}.w": w,
                    f"{tag}.b": b, f"{tag}.delta": d, f"{tag}.y": y, f"{tag}.dw": dw,
                    f"{tag}.db": db, f"{tag}.dx": dx})

    # pool: (B, C, H, W, k, step); values quantised so that ties occur
    for tag, (B, C, H, W, k, st) in {"pool_a": (2, 4, 11, 11, 2, 2), "pool_b": (1, 3, 10, 9, 3, 2),
                                     "pool_c": (2, 2, 8, 8, 2, 1)}.items():
        x = (np.round(rnd(B, C, H, W) * 2) / 2).astype(np.float32)
        x[0, 0, 0, :3] = [np.nan, 1.0, -0.0]
        y, mask, _ = ref.maxpool(x, k, st)
        d = rnd(*y.shape)
        y, mask, dx = ref.maxpool(x, k, st, d)
        out.update({f"{tag}.cfg": np.array([B, C, H, W, k, st]), f"{tag}.x": x, f"{tag}.delta": d,
                    f"{tag}.y": y, f"{tag}.mask": mask, f"{tag}.dx": dx})

    # relu incl. -0.0, NaN, inf
    x = rnd(2, 3, 5, 7)
    x.flat[:6] = [0.0, -0.0, np.nan, np.inf, -np.inf, -1e-38]
    d = rnd(*x.shape)
    y, dd = ref.relu(x, d)
    out.update({"relu.x": x, "relu.delta": d, "relu.y": y, "relu.dx": dd})

    # linear (input keeps a 3-D shape, linear.cpp:25)
    x = rnd(3, 8, 3, 3)
    w = (rnd(72, 5) / 10).astype(np.float32)
    b = (rnd(5) / 10).astype(np.float32)
    d = rnd(3, 5)
    y, dw, db, dx = ref.linear(x, w, b, d)
    out.update({"linear.x": x, "linear.w": w, "linear.b": b, "linear.delta": d, "linear.y": y,
                "linear.dw": dw, "linear.db": db, "linear.dx": dx})

    # batchnorm: train forward+backward, then eval forward with the updated moving stats
    x = (rnd(3, 6, 7, 5) * 2 + 1).astype(np.float32)
    g = (1 + rnd(6) / 4).astype(np.float32)
    bt = (rnd(6) / 4).astype(np.float32)
    mm = (rnd(6) / 4).astype(np.float32)
    mv = np.abs(rnd(6)).astype(np.float32)
    d = rnd(*x.shape)
    r = ref.batchnorm(x, g, bt, mm, mv, d)
    e = ref.batchnorm(x, g, bt, r["moving_mean"], r["moving_var"], eval_mode=True)
    out.update({"bn.x": x, "bn.gamma": g, "bn.beta": bt, "bn.mm": mm, "bn.mv": mv, "bn.delta": d,
                "bn.y": r["y"], "bn.xhat": r["xhat"], "bn.mean": r["mean"], "bn.var": r["var"],
                "bn.mm_out": r["moving_mean"], "bn.mv_out": r["moving_var"], "bn.dx": r["dx"],
                "bn.dgamma": r["dgamma"], "bn.dbeta": r["dbeta"], "bn.y_eval": e["y"]})

    # softmax / xent / argmax incl. clamp thresholds, ties and the 0*log(0)=NaN loss quirk
    z = rnd(6, 3) * 3
    z[1] = [0.0, 60.0, 0.0]      # p underflows to 0 -> loss NaN (func.cpp:67)
    z[2] = [100.0, 5.0, -100.0]  # exp clamp branches
    z[3] = [2.0, 2.0, 1.0]       # argmax tie -> first
    lab = np.array([0, 1, 2, 0, 1, 2], np.int32)
    p, pred, loss, dl = ref.softmax_xent(z, lab)
    z2 = z[[0, 2, 3, 4, 5]]
    lab2 = lab[[0, 2, 3, 4, 5]]
    p2, pred2, loss2, dl2 = ref.softmax_xent(z2, lab2)
    out.update({"xent.z": z, "xent.labels": lab, "xent.p": p, "xent.pred": pred,
                "xent.loss": np.float32(loss), "xent.delta": dl,
                "xent2.z": z2, "xent2.labels": lab2, "xent2.p": p2, "xent2.pred": pred2,
                "xent2.loss": np.float32(loss2), "xent2.delta": dl2})
    np.savez_compressed(os.path.join(HERE, "ops_golden.npz"), **out)


def train():
    init = ref.alexnet_init_params(3, False)
    init.tofile(os.path.join(HERE, "alexnet_init.model"))
    B = 4
    x = synth_images(B, 3, 224, 224, seed=1234)
    lab = synth_labels(B, 3)
    net = ref.Net()  # the reference AlexNet container itself
    net.set_params(init)
    out = {"labels": lab, "lr": np.float32(1e-3)}
    for step in range(3):
        loss, probs, dx = net.train_step(x, lab, 1e-3, want_dx=True)
        out[f"loss{step}"] = np.float32(loss)
        out[f"probs{step}"] = probs
        if step == 0:
            out["grads0"] = net.get_grads()
            out["dx_image0_sample"] = dx[:, :, ::7, ::5].copy()
            for li in (0, 2, 3, 9):  # conv1, pool1, conv2, linear outputs (subsampled)
                out[f"layer{li}_out_sample"] = net.layer_output(li, B)[::97].copy()
    out["params3"] = net.get_params()
    # BN variant, 2 steps, to pin BatchNorm inside the full loop
    initbn = ref.alexnet_init_params(3, True)
    netbn = ref.Net(batch_norm=True)
    netbn.set_params(initbn)
    for step in range(2):
        loss, probs, _ = netbn.train_step(x, lab, 1e-3)
        out[f"bn_loss{step}"] = np.float32(loss)
        out[f"bn_probs{step}"] = probs
    out["bn_params2_sample"] = netbn.get_params()[::13].copy()
    np.savez_compressed(os.path.join(HERE, "train_golden.npz"), **out)


if __name__ == "__main__":
    ref.build()
    kat()
    gradcam()
    ops()
    train()
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
