"""fp64 evaluation of one train step (checker only: the fp64 tie-breaker of SURVEY 8c) by torch autograd on the GPU."""
import numpy as np
import torch

from cnn_b200 import nets


def fp64_step(spec, params, x, lab):
    """The same train step evaluated in fp64 by torch autograd on the GPU (checker only): loss = -(1/B) sum log p[label],
    whose parameter gradients are exactly the reference's (1/B inside conv / linear gradients, func.cpp:56-73)."""
    import torch.nn.functional as F
    lay, _ = nets.param_layout(spec)
    P = {}
    for li, kind, off, n in lay:
        P[(li, kind)] = torch.tensor(params[off:off + n], dtype=torch.float64, device="cuda", requires_grad=True)
    h = torch.tensor(x, dtype=torch.float64, device="cuda")
    for li, (t, a, b, c, d) in enumerate(spec):
        if t == nets.CONV:
            h = F.conv2d(h, P[(li, "w")].view(b, a, c, c), P[(li, "b")], stride=d)
        elif t == nets.RELU:
            h = torch.relu(h)
        elif t == nets.POOL:
            h = F.max_pool2d(h, a, b)
        elif t == nets.LINEAR:
            h = h.reshape(h.shape[0], -1) @ P[(li, "w")].view(a, b) + P[(li, "b")]
    logp = torch.log_softmax(h, dim=1)
    loss = -logp[torch.arange(h.shape[0]), torch.tensor(lab, dtype=torch.long, device="cuda")].mean()
    loss.backward()
    g = np.zeros(len(params), np.float64)
    for li, kind, off, n in lay:
        g[off:off + n] = P[(li, kind)].grad.reshape(-1).cpu().numpy()
    return float(loss.item()), h.detach().cpu().numpy(), g


