"""Golden fixtures for block-wise reconstruction from the UNMODIFIED reference.

Run in the build container only:   python tests/golden/make_block_golden.py
Writes tests/golden/block_*.npz.  The reference's block_reconstruction (calib_block.py:91-183) draws its mini-batches
with an unseeded torch.randperm and its QDrop masks with torch.rand_like; both are recorded here (the functions are
wrapped for the duration of the call) so that the oracle and the CUDA path can replay the same draws.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_shims"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, HERE)

import models  # noqa: E402  (reference)
import quantization  # noqa: E402  (reference)
import quantization.calib_block as cb  # noqa: E402
from quantization.quant_layer import QuantModule  # noqa: E402
from make_golden import TINY_HNERV, TINY_NERV, npy  # noqa: E402

torch.set_num_threads(8)


def repaired_layer_reconstruction():
    """calib_layer.layer_reconstruction stops at calib_layer.py:130 (`opt_params +=` with no prior assignment).  Compile
    the reference's OWN source with the one missing statement inserted, in the module's namespace, in memory only."""
    import inspect
    import quantization.calib_layer as cl
    src = inspect.getsource(cl.layer_reconstruction)
    needle = "    opt_params += [layer.weight_quantizer.alpha]"
    assert src.count(needle) == 1
    src = src.replace(needle, "    opt_params = []\n" + needle)
    ns = {}
    exec(compile(src, "<calib_layer.layer_reconstruction + opt_params = []>", "exec"), cl.__dict__, ns)
    return cl, ns["layer_reconstruction"]


def run_block_case(tag, arch, cfg, bits, hadamard, block_idx, asym, input_prob, iters=60, n_frames=20, bsz=2, opt_mode="mse",
                   layer=None):
    """layer: None = block_reconstruction on decoder[block_idx]; 'conv' = the repaired layer_reconstruction on that
    block's convolution; 'head' / 'stem' = on the head layer / decoder[0]."""
    global cb
    cb_block = cb
    if layer is not None:
        cb, layer_fn = repaired_layer_reconstruction()
    try:
        _run_case(tag, arch, cfg, bits, hadamard, block_idx, asym, input_prob, iters, n_frames, bsz, opt_mode, layer,
                  layer_fn if layer is not None else None)
    finally:
        cb = cb_block


def _run_case(tag, arch, cfg, bits, hadamard, block_idx, asym, input_prob, iters, n_frames, bsz, opt_mode, layer, layer_fn):
    torch.manual_seed(903)
    model = (models.HNeRV if arch == "hnerv" else models.NeRV)(cfg)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "encoder" not in n and p.dim() == 1:
                p.add_(torch.randn_like(p) * 0.05)
    g = torch.Generator().manual_seed(11)
    frames = torch.rand(n_frames, 3, cfg["crop_h"], cfg["crop_w"], generator=g)
    with torch.no_grad():
        cali = model.encode(frames) * 3.0 if arch == "hnerv" else model.encode(torch.arange(n_frames).float() / n_frames)
    out = {"cali": npy(cali), "bits": np.array(bits), "hadamard": np.array(hadamard), "block_idx": np.array(block_idx),
           "asym": np.array(asym), "input_prob": np.array(input_prob), "iters": np.array(iters), "bsz": np.array(bsz)}
    for k, v in model.state_dict().items():
        if "encoder" not in k:
            out["sd/" + k] = npy(v)
    qnn = quantization.QuantModel(model=model, hadamard=hadamard,
                                  weight_quant_params={"n_bits": 8, "channel_wise": True, "scale_method": "max"})
    qnn.set_bitwidth(bits)
    qnn.eval()
    qnn.set_quant_state(True)
    with torch.no_grad():
        qnn(cali[:bsz])  # first quantised forward initialises every step size
    block = {None: lambda: qnn.model.decoder[block_idx], "conv": lambda: qnn.model.decoder[block_idx].conv,
             "head": lambda: qnn.model.head_layer, "stem": lambda: qnn.model.decoder[0]}[layer]()
    conv = [m for m in block.modules() if isinstance(m, QuantModule)][0]
    assert layer is None or block is conv
    out["layer"] = np.array(layer or "")
    out["stage"] = np.array([i for i, m in enumerate(x for x in qnn.model.modules() if isinstance(x, QuantModule)) if m is conv][0])

    idx_log, mask_log, traj, cache = [], [], [], {}
    _randperm, _rand_like, _call, _save = torch.randperm, torch.rand_like, cb.LossFunction.__call__, cb.save_inp_oup_data
    _save_grad = cb.save_grad_data
    import quantization.data_utils as du
    _get_grad = du.GetLayerGrad.__call__
    raw = []

    def get_grad(self, x):
        r = _get_grad(self, x)
        raw.append(r.detach().cpu().clone())  # before |g| + 1 swallows it (data_utils.py:113)
        return r

    def save_grad(*a, **k):
        r = _save_grad(*a, **k)
        cache["grad"] = r.detach().cpu().clone()
        return r

    def randperm(n, *a, **k):
        r = _randperm(n, *a, **k)
        idx_log.append(r[:bsz].clone())
        return r

    def rand_like(x, *a, **k):
        r = _rand_like(x, *a, **k)
        mask_log.append(r.clone())
        return r

    def rec_call(self, pred, tgt, grad=None):
        saved = (self.count, self.round)
        self.round = "none"  # the reference's own reconstruction term alone (total = 0 + rec_loss, calib_block.py:75-85)
        with torch.no_grad():
            rec = float(_call(self, pred, tgt, grad))
        self.count, self.round = saved
        tot = _call(self, pred, tgt, grad)
        traj.append((self.count, float(tot), rec, float(self.round_loss)))
        return tot

    def save(*a, **k):
        r = _save(*a, **k)
        cache["inp"], cache["sym"], cache["out"] = r[0][0], r[0][1], r[1]
        return r

    import logging
    logging.getLogger().setLevel(logging.WARNING)
    torch.randperm, torch.rand_like, cb.LossFunction.__call__, cb.save_inp_oup_data = randperm, rand_like, rec_call, save
    cb.save_grad_data = save_grad
    du.GetLayerGrad.__call__ = get_grad
    try:
        torch.manual_seed(5)
        (layer_fn or cb.block_reconstruction)(qnn, block, cali, batch_size=bsz, iters=iters, weight=0.01, opt_mode=opt_mode,
                                              asym=asym, b_range=(20, 2), warmup=0.2, input_prob=input_prob, p=2.0, lr=0.003)
    finally:
        torch.randperm, torch.rand_like, cb.LossFunction.__call__, cb.save_inp_oup_data = _randperm, _rand_like, _call, _save
        cb.save_grad_data = _save_grad
        du.GetLayerGrad.__call__ = _get_grad
    out["opt_mode"] = np.array(opt_mode)
    if "grad" in cache:
        # |g| + 1 in fp32 keeps ~3 digits of g; store the cache as the reference holds it
        out["cache_grad"] = npy(cache["grad"])
        out["raw_grad"] = npy(torch.cat(raw))
    out["idx"] = np.stack([npy(i) for i in idx_log])
    if mask_log:
        out["masks"] = np.stack([npy(m) for m in mask_log]).astype(np.float32)
    out["traj"] = np.array(traj, dtype=np.float64)
    out["cache_inp"], out["cache_sym"], out["cache_out"] = npy(cache["inp"]), npy(cache["sym"]), npy(cache["out"])
    wq, bq = conv.weight_quantizer, conv.bias_quantizer
    out["final/alpha_w"], out["final/alpha_b"] = npy(wq.alpha), npy(bq.alpha)
    out["final/delta_w"], out["final/zp_w"] = npy(wq.delta), npy(wq.zero_point)
    out["final/delta_b"], out["final/zp_b"] = npy(bq.delta), npy(bq.zero_point)
    assert wq.soft_targets is False and bq.soft_targets is False  # calib_block.py:180-183: both go hard
    with torch.no_grad():
        y = block(cache["inp"][:bsz])
    out["final/block_out"] = npy(y)
    out["final/codes_w"], out["final/codes_b"] = npy(wq.x_quant), npy(bq.x_quant)
    np.savez_compressed(os.path.join(HERE, f"{tag}.npz"), **out)
    print(tag, "block", type(block).__name__, tuple(conv.weight.shape), "iters", len(traj), "first/last rec",
          traj[0][2], traj[-1][2], "masks", len(mask_log))


if __name__ == "__main__":
    only = sys.argv[1:]
    if only:
        _all = run_block_case
        run_block_case = lambda tag, *a, **k: _all(tag, *a, **k) if tag in only else None  # noqa: E731
    run_block_case("block_tiny_hnerv", "hnerv", TINY_HNERV, [6, 5, 4, 5, 5, 6, 6], False, 3, False, 1.0)
    run_block_case("block_tiny_hnerv_qdrop", "hnerv", TINY_HNERV, [6, 5, 4, 5, 5, 6, 6], False, 2, True, 0.5)
    # (hadamard=True cannot run in the reference: calib_block.py:125 builds alpha from the UNROTATED org_weight while the
    #  forward quantises the rotated, channel-padded copy -- shape mismatch at quantizer.py:291)
    run_block_case("block_tiny_nerv", "nerv", TINY_NERV, [5, 6, 3, 4, 5, 4, 3], False, 2, False, 1.0)
    # coarse predecessors so that the output gradients reach the resolution of fp32 (|g| + 1): at 5-6 bits the cache is 1.0
    run_block_case("block_tiny_hnerv_fdiag", "hnerv", TINY_HNERV, [2, 2, 2, 3, 5, 6, 6], False, 3, False, 1.0, opt_mode="fisher_diag")
    run_block_case("block_tiny_hnerv_ffull", "hnerv", TINY_HNERV, [6, 5, 4, 5, 5, 6, 6], False, 2, True, 0.5, opt_mode="fisher_full")
    # layer_reconstruction, repaired in memory (see repaired_layer_reconstruction)
    run_block_case("layer_tiny_hnerv_conv", "hnerv", TINY_HNERV, [6, 5, 4, 5, 5, 6, 6], False, 3, True, 0.5, layer="conv", n_frames=10)
    run_block_case("layer_tiny_hnerv_head", "hnerv", TINY_HNERV, [6, 5, 4, 5, 5, 6, 4], False, 0, False, 1.0, layer="head",
                   opt_mode="fisher_diag", n_frames=10)
    run_block_case("layer_tiny_nerv_stem", "nerv", TINY_NERV, [4, 6, 3, 4, 5, 4, 3], False, 0, False, 1.0, layer="stem")
