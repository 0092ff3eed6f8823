"""Shim for the PyPI `hadamard-transform` package (un-pinned, absent from /root/reference).

Only call site: /root/reference/quantization/quant_layer.py:19 -- an orthonormal (1/sqrt(n))
Walsh-Hadamard transform along the LAST dim, batched over leading dims, self-inverse
(pinned by the self-check at quant_layer.py:94-100).  Restated as the Sylvester-ordered
butterfly; scipy.linalg.hadamard(n)/sqrt(n) is the ground truth used in tests.
"""
import math
import torch
import torch.nn.functional as F


def hadamard_transform(x: torch.Tensor, normalize: bool = True) -> torch.Tensor:
    n = x.shape[-1]
    assert n & (n - 1) == 0, "last dim must be a power of two"
    shape = x.shape
    y = x.reshape(-1, n)
    h = 1
    while h < n:
        y = y.view(-1, n // (2 * h), 2, h)
        a, b = y[:, :, 0, :], y[:, :, 1, :]
        y = torch.stack((a + b, a - b), dim=2).reshape(-1, n)
        h *= 2
    if normalize:
        y = y / math.sqrt(n)
    return y.view(shape)


def pad_to_power_of_2(x: torch.Tensor) -> torch.Tensor:
    n = x.shape[-1]
    m = 1 if n == 0 else 2 ** math.ceil(math.log2(n))
    return F.pad(x, (0, m - n))
