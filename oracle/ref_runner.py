"""Drives the UNMODIFIED reference staged under oracle/_ref/ (oracle/make_ref.py) through its own public API on the host
cores: `quantization.QuantModel` + `quantization.model_reconstruction` for calibration iterations, `QuantModel.forward`
for the quantised decode.  TEST / BENCH INFRASTRUCTURE: imported only by bench.py's CPU legs (`--impl reference`,
`cpu_baseline`) and by tests.  Nothing of `neuroquant_b200` is imported here.

The reference's loop is monolithic (calib_model.py:92-240), so K timed iterations are obtained by handing it a loader of
W + K mini-batches with iters = W + K: int(0.05 * iters / len(gt)) = 0 step-size epochs and exactly one AdaRound epoch of
W + K iterations (:144, :203-206); the rounding regulariser switches on after 20 % of them, as in a real run.  Iteration
boundaries are time-stamped from a wrapper around `LossFunction.__call__` (the reference's code is not edited)."""
from __future__ import annotations

import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "quantization", "calib_model.py"))


def _import_reference():
    for p in (REF, os.path.join(HERE, "ref_shims")):
        if p not in sys.path:
            sys.path.insert(0, p)
    for name in ("models", "quantization", "utils"):
        m = sys.modules.get(name)
        if m is not None and not str(getattr(m, "__file__", "")).startswith(REF):
            raise ImportError(f"module {name!r} is already imported from {getattr(m, '__file__', '?')}: the reference arm needs its own process")
    import models
    import quantization
    import quantization.calib_model as cm
    from utils import get_config
    return models, quantization, cm, get_config


class ListLoader(list):
    """Stand-in for the DataLoader `gt` of model_reconstruction: len() + iteration of sample dicts."""


def build(workload: str, bits, hadamard: bool, seed: int = 903):
    """Reference model of `workload` with the seeded decoder weights of oracle.workloads.random_stages, wrapped in the
    reference's QuantModel with scales initialised (calibrate_network.py:218-238)."""
    from . import workloads as W
    models, quantization, cm, get_config = _import_reference()
    arch, cfg_dict = W.WORKLOADS[workload]
    if workload in W.REFERENCE_YAML:
        cfg = get_config(os.path.join(REF, W.REFERENCE_YAML[workload]))
    else:
        cfg = dict(cfg_dict)
    model = (models.HNeRV if arch == "hnerv" else models.NeRV)(cfg)
    convs = [model.decoder[0]] + [blk.conv[0] for blk in list(model.decoder)[1:]] + [model.head_layer]
    with torch.no_grad():
        for c, st in zip(convs, W.random_stages(cfg_dict, arch, seed)):
            c.weight.copy_(st.weight)
            c.bias.copy_(st.bias)
    import copy
    build.fp_model = copy.deepcopy(model)  # the full-precision network, before the (in-place) module surgery
    qnn = quantization.QuantModel(model=model, hadamard=hadamard,
                                  weight_quant_params={"n_bits": 8, "channel_wise": True, "scale_method": "max"})
    qnn.set_bitwidth(list(bits))
    qnn.eval()
    qnn.set_quant_state(True)
    return qnn, arch, cfg_dict, (quantization, cm)


def time_calibration(workload: str, bits, hadamard: bool, batch: int, steps: int, warmup: int, hyper: dict, threads: int):
    """Returns (seconds per iteration over the `steps` timed AdaRound iterations, decode seconds per batch, qnn)."""
    from . import workloads as W
    torch.set_num_threads(threads)
    qnn, arch, cfg, (quantization, cm) = build(workload, bits, hadamard)
    gen = torch.Generator().manual_seed(903)
    c, h, w = W.embed_shape(cfg, arch)
    F = 8
    embeds = torch.randn(F, c, h, w, generator=gen)
    frames = torch.rand(F, 3, cfg["crop_h"], cfg["crop_w"], generator=gen)
    with torch.no_grad():
        qnn(embeds[:batch])  # first quantised forward initialises the step sizes (quantizer.py:112-115)
    n = warmup + steps
    samples = []
    for i in range(n):
        o = (i * batch) % (F - batch + 1)
        idx = torch.arange(o, o + batch)
        samples.append({"img": frames[o:o + batch], "norm_idx": idx.float() / F, "idx": idx})
    stamps = []
    _call = cm.LossFunction.__call__

    def stamped(self, pred, tgt, grad=None):
        stamps.append(time.perf_counter())  # forward of iteration len(stamps) done; loss, backward, Adam follow
        return _call(self, pred, tgt, grad)

    cm.LossFunction.__call__ = stamped
    try:
        quantization.model_reconstruction(qnn, cali_data=embeds, gt=ListLoader(samples), arch=arch, batch_size=batch, iters=n,
                                          weight=hyper["weight"], opt_mode="mse", hadamard=hadamard, b_range=hyper["b_range"],
                                          warmup=hyper["warmup"], p=hyper["p"], lr=hyper["lr"])
    finally:
        cm.LossFunction.__call__ = _call
    t_end = time.perf_counter()
    assert len(stamps) == n, (len(stamps), n)
    # iteration i spans stamp[i] .. stamp[i+1] (same phase of consecutive iterations); the last one is closed by the
    # return of model_reconstruction minus nothing measurable (the final soft_targets switch)
    stamps.append(t_end)
    per_iter = (stamps[n] - stamps[warmup]) / steps
    with torch.no_grad():
        qnn(embeds[:batch])
        t0 = time.perf_counter()
        reps = 2
        for _ in range(reps):
            qnn(embeds[:batch])  # hard-rounded weights after calibration (calib_model.py:231-240)
        dec = (time.perf_counter() - t0) / reps
    return per_iter, dec, qnn
