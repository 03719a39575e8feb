import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with `-m gpu` on the GPU box)")


@pytest.fixture(scope="session")
def weights_default():
    from erweitern_55_b200 import weights as W
    return W.init_weights(seed=55, perturbed=False)


@pytest.fixture(scope="session")
def weights_perturbed():
    from erweitern_55_b200 import weights as W
    return W.init_weights(seed=55, perturbed=True)


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "forward_golden.npz"))


def calibrate_trained_like(seed=77, blocks=7, filters=128, logit_std=3.0, batch=256):
    """A weight set with the dynamic range of a trained network instead of Keras' initialisers: every BatchNorm's moving
    statistics are the actual statistics of its input on a batch of synthetic positions (so they are far from the
    identity: means of a few tenths, variances from 0.01 to 10), gammas in [0.5, 2], betas ~ N(0, 0.5), biases
    ~ N(0, 0.2), and the policy 1x1 kernel scaled until the logits have standard deviation `logit_std` (range beyond
    +-10).  Built layer by layer with torch on the CPU from erweitern_55_b200.weights' own layout."""
    import numpy as np
    import torch
    import torch.nn.functional as F
    from erweitern_55_b200 import synth, weights as W
    rng = np.random.default_rng(seed)
    names = [n for n, _ in W.weight_spec(blocks, filters)]
    ws = dict(zip(names, W.init_weights(blocks, filters, seed=seed, perturbed=True)))
    x = torch.from_numpy(synth.synthetic_planes(batch, seed=seed + 1))

    def conv(t, layer):
        ws[layer + "/bias"] = rng.normal(0.0, 0.2, size=ws[layer + "/bias"].shape).astype(np.float32)
        k = torch.from_numpy(ws[layer + "/kernel"]).permute(3, 2, 0, 1).contiguous()
        return F.conv2d(t, k, torch.from_numpy(ws[layer + "/bias"]), padding=k.shape[2] // 2)

    def bn(t, layer):
        c = t.shape[1]
        mean = t.mean(dim=(0, 2, 3)).numpy()
        var = t.var(dim=(0, 2, 3), unbiased=False).numpy() + 1e-4
        ws[layer + "/moving_mean"], ws[layer + "/moving_variance"] = mean.astype(np.float32), var.astype(np.float32)
        ws[layer + "/gamma"] = rng.uniform(0.5, 2.0, size=c).astype(np.float32)
        ws[layer + "/beta"] = rng.normal(0.0, 0.5, size=c).astype(np.float32)
        s = torch.from_numpy(ws[layer + "/gamma"] / np.sqrt(ws[layer + "/moving_variance"] + 1e-3))
        return (t - torch.from_numpy(ws[layer + "/moving_mean"]).view(1, -1, 1, 1)) * s.view(1, -1, 1, 1) + torch.from_numpy(ws[layer + "/beta"]).view(1, -1, 1, 1)

    conv_names = [n.rsplit("/", 1)[0] for n in names if n.endswith("/kernel")]
    bn_names = [n.rsplit("/", 1)[0] for n in names if n.endswith("/gamma")]
    trunk_convs, trunk_bns = conv_names[:1 + 2 * blocks], bn_names[:1 + 2 * blocks]
    with torch.no_grad():
        h = F.relu(bn(conv(x, trunk_convs[0]), trunk_bns[0]))
        for b in range(blocks):
            y = F.relu(bn(conv(h, trunk_convs[1 + 2 * b]), trunk_bns[1 + 2 * b]))
            h = F.relu(bn(conv(y, trunk_convs[2 + 2 * b]), trunk_bns[2 + 2 * b]) + h)
        head_convs = {tuple(ws[n + "/kernel"].shape): n for n in conv_names[1 + 2 * blocks:]}
        head_bns = {ws[n + "/gamma"].shape[0]: n for n in bn_names[1 + 2 * blocks:]}
        p = F.relu(bn(conv(h, head_convs[(3, 3, filters, filters)]), head_bns[filters]))
        pol = head_convs[(1, 1, filters, 69)]
        logits = conv(p, pol)
        ws[pol + "/kernel"] = (ws[pol + "/kernel"] * (logit_std / float(logits.std()))).astype(np.float32)
        bn(conv(h, head_convs[(1, 1, filters, 1)]), head_bns[1])
        ws[head_bns[1] + "/beta"][:] = 0.5        # keep the single value-head channel alive
    return [ws[n] for n in names]


@pytest.fixture(scope="session")
def weights_trained_like():
    return calibrate_trained_like()
