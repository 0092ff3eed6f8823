"""Golden vectors for the 'mse' / 'l1' / 'gaussian' scale initialisers (quantizer.py:170-222) from the UNMODIFIED
reference.  Run in the build container only:   python tests/golden/make_init_golden.py  -> tests/golden/scale_inits.npz"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_shims"))
sys.path.insert(0, "/root/reference")

from quantization.quantizer import UniformAffineQuantizer  # noqa: E402

if __name__ == "__main__":
    g = torch.Generator().manual_seed(903)
    w = torch.randn(24, 11, 3, 3, generator=g) * 0.2
    w[2] = w[2].abs() + 0.01        # one-sided channel: the search does NOT clamp the range to zero
    w[5] *= 30.0                    # a wide channel
    w[7, 0, 0, 0] = 4.0             # an outlier: the shrunk ranges win
    w[9] = 0.0                      # all-zero channel: delta = eps
    b = torch.randn(24, generator=g) * 0.1
    out = {"w": w.numpy(), "b": b.numpy()}
    for method in ("mse", "l1", "gaussian"):
        for bits in (2, 4, 6, 8):
            for name, x in (("w", w), ("b", b)):
                q = UniformAffineQuantizer(n_bits=8, channel_wise=True, scale_method=method)
                q.bitwidth_refactor(bits)
                y = q(x)
                out[f"{method}{bits}_{name}_delta"] = q.delta.detach().numpy().copy()
                out[f"{method}{bits}_{name}_zp"] = np.asarray(q.zero_point.detach().numpy()).copy()
                out[f"{method}{bits}_{name}_deq"] = y.detach().numpy().copy()
    np.savez_compressed(os.path.join(HERE, "scale_inits.npz"), **out)
    print("scale_inits", len(out))
