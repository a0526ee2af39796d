"""The numpy restatement of the reference's photometric loss (L1 + fused SSIM 'valid', src/training/trainer.cpp:103-131,
src/training/kernels/ssim.cu) against an independent torch implementation (conv2d + autograd, float64)."""
import numpy as np
import pytest
import torch

import oracle as O


def _torch_loss(image, target, lam):
    x = torch.tensor(image, dtype=torch.float64, requires_grad=True)
    y = torch.tensor(target, dtype=torch.float64)
    X = torch.clamp(x, 0.0, 1.0).permute(2, 0, 1)[None]
    Y = y.permute(2, 0, 1)[None]
    g = torch.tensor(O._SSIM_G, dtype=torch.float64)
    k2 = torch.outer(g, g)[None, None].repeat(3, 1, 1, 1)
    conv = lambda t: torch.nn.functional.conv2d(t, k2, padding=5, groups=3)  # noqa: E731
    mu1, mu2 = conv(X), conv(Y)
    s1, s2, s12 = conv(X * X) - mu1 * mu1, conv(Y * Y) - mu2 * mu2, conv(X * Y) - mu1 * mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    m = ((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (s1 + s2 + C2))
    H, W = m.shape[2:]
    if H > 10 and W > 10:
        m = m[:, :, 5:H - 5, 5:W - 5]
    loss = (1 - lam) * (X - Y).abs().mean() + lam * (1 - m.mean())
    loss.backward()
    return float(loss), x.grad.numpy()


@pytest.mark.parametrize("hw", [(40, 56), (9, 30)])
def test_photometric_loss_matches_torch_autograd(hw):
    rng = np.random.RandomState(3)
    h, w = hw
    image = rng.uniform(-0.1, 1.1, size=(h, w, 3))  # some values outside [0,1]: the clamp must cut their gradient
    target = rng.uniform(0, 1, size=(h, w, 3))
    for lam in (0.2, 0.0, 1.0):
        loss, grad, _ = O.photometric_loss(image, target, lam)
        tl, tg = _torch_loss(image, target, lam)
        assert abs(loss - tl) <= 1e-12 * max(1.0, abs(tl))
        assert np.abs(grad - tg).max() <= 1e-12 + 1e-9 * np.abs(tg).max()
