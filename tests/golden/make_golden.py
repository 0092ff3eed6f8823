"""Generate golden fixtures from the UNMODIFIED reference at /root/reference.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
Writes tests/golden/*.npz.  The reference is imported with the three shim modules under
oracle/ref_shims (timm.models.layers, pytorch_msssim, hadamard_transform); no reference source
is copied.  All inputs are seeded; the mini-batch order is injected because the reference never
seeds its shuffle (calibrate_network.py:161).
"""
import copy
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_shims"))
sys.path.insert(0, "/root/reference")

import models  # noqa: E402  (reference)
import quantization  # noqa: E402  (reference)
from quantization.quantizer import UniformAffineQuantizer, AdaRoundQuantizer, lp_loss  # noqa: E402
from quantization.quant_layer import hadamard_along_channel_weight, QuantModule  # noqa: E402
from quantization.data_utils import LinearTempDecay  # noqa: E402
from utils import get_config, psnr_fn_single  # noqa: E402

torch.set_num_threads(8)

TINY_HNERV = dict(crop_h=32, crop_w=64, diff_enc=False, stage_block=1,
                  enc_strides=[2, 2, 2, 2, 1], enc_channel=[8, 8, 8, 8, 4],
                  channel_reduce=1.2, channel_lbound=6, dec_in_channel=20,
                  dec_kernels=[1, 3, 5, 5, 3], dec_strides=[2, 2, 2, 2, 1],
                  dec_norm="none", dec_acts="gelu", out_bias="tanh")
TINY_NERV = dict(crop_h=32, crop_w=64, diff_enc=False, base=1.25, level=10,
                 channel_reduce=2, channel_lbound=6, dec_in_channel=18,
                 dec_kernels=[3, 3, 3, 3, 3], dec_strides=[2, 2, 2, 2, 1],
                 dec_norm="none", dec_acts="gelu", out_bias="tanh")


def npy(t):
    return t.detach().cpu().numpy().copy()


def quantizer_kats():
    """UAQ 'max' init + forward, AdaRound init/soft/hard, lp_loss, b-schedule."""
    out = {}
    g = torch.Generator().manual_seed(903)
    w = torch.randn(12, 7, 3, 3, generator=g) * 0.3
    w[3] = 0.0  # an all-zero channel: delta clamps to eps
    w[5] = w[5].abs()  # one-sided channel: zero_point 0
    b = torch.randn(12, generator=g) * 0.1
    out["w"], out["b"] = npy(w), npy(b)
    for bits in (2, 3, 4, 6, 8):
        for name, x in (("w", w), ("b", b)):
            q = UniformAffineQuantizer(n_bits=8, channel_wise=True, scale_method="max")
            q.bitwidth_refactor(bits)
            y = q(x)
            out[f"uaq{bits}_{name}_delta"] = npy(q.delta)
            out[f"uaq{bits}_{name}_zp"] = npy(q.zero_point)
            out[f"uaq{bits}_{name}_deq"] = npy(y)
            # gradient of sum(y * r) wrt delta (STE)
            r = torch.randn(x.shape, generator=torch.Generator().manual_seed(bits))
            (y * r).sum().backward()
            out[f"uaq{bits}_{name}_r"] = npy(r)
            out[f"uaq{bits}_{name}_ddelta"] = npy(q.delta.grad)
            if name == "w" and bits == 3:
                continue  # all-zero channel -> delta 1e-8 -> fp16 0 -> NaN alpha; keep one case
            xa = x.clone()
            if name == "w":
                xa[3] = w[4] * 0.5  # avoid the NaN channel in the AdaRound KATs
                q2 = UniformAffineQuantizer(n_bits=8, channel_wise=True, scale_method="max")
                q2.bitwidth_refactor(bits)
                q2(xa)
            else:
                q2 = q
            a = AdaRoundQuantizer(uaq=q2, round_mode="learned_hard_sigmoid", weight_tensor=xa.data)
            out[f"ada{bits}_{name}_x"] = npy(xa)
            out[f"ada{bits}_{name}_delta"] = npy(a.delta)
            out[f"ada{bits}_{name}_zp"] = npy(a.zero_point)
            out[f"ada{bits}_{name}_alpha0"] = npy(a.alpha)
            # perturb alpha so soft targets cover the clamp regions
            with torch.no_grad():
                a.alpha.add_(torch.randn(a.alpha.shape, generator=torch.Generator().manual_seed(7 + bits)) * 2.0)
            out[f"ada{bits}_{name}_alpha"] = npy(a.alpha)
            a.soft_targets = True
            y = a(xa)
            out[f"ada{bits}_{name}_soft_codes"] = npy(a.x_quant)
            out[f"ada{bits}_{name}_soft_deq"] = npy(y)
            reg = (1 - ((a.get_soft_targets() - .5).abs() * 2).pow(7.5)).sum()
            ((y * r).sum() + 0.01 * reg).backward()
            out[f"ada{bits}_{name}_reg_b7.5"] = npy(reg)
            out[f"ada{bits}_{name}_dalpha"] = npy(a.alpha.grad)
            a.soft_targets = False
            y = a(xa)
            out[f"ada{bits}_{name}_hard_codes"] = npy(a.x_quant)
            out[f"ada{bits}_{name}_hard_deq"] = npy(y)
    # rotation
    wr = torch.randn(5, 8, 3, 3, generator=g)
    out["had_in"] = npy(wr)
    out["had_out"] = npy(hadamard_along_channel_weight(wr))
    # lp_loss
    p_, t_ = torch.rand(2, 3, 5, 7, generator=g), torch.rand(2, 3, 5, 7, generator=g)
    out["lp_pred"], out["lp_tgt"] = npy(p_), npy(t_)
    out["lp_p2"] = npy(lp_loss(p_, t_, p=2.0))
    out["lp_p24"] = npy(lp_loss(p_, t_, p=2.4))
    # b schedule as calibrate_network uses it (iters 21000, warmup .2, b 20->2)
    td = LinearTempDecay(21000, rel_start_decay=0.2, start_b=20, end_b=2)
    ts = np.array([1, 4199, 4200, 4500, 10000, 19500, 19998, 21000])
    out["b_t"] = ts
    out["b_val"] = np.array([td(int(t)) for t in ts], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "quantizer_kats.npz"), **out)
    print("quantizer_kats", len(out))


class ListLoader(list):
    """Stand-in for the DataLoader `gt` of model_reconstruction: len() + iteration of dicts."""


def run_model_case(tag, arch, cfg, bits, hadamard, iters, n_frames=8, bsz=2, omega=True, only_layers=False, channel_wise=True):
    torch.manual_seed(903)
    model = (models.HNeRV if arch == "hnerv" else models.NeRV)(cfg)
    with torch.no_grad():  # non-degenerate random decoder weights (default init is fine) + biases
        for n, p in model.named_parameters():
            if "encoder" not in n and p.dim() == 1:
                p.add_(torch.randn_like(p) * 0.05)
    g = torch.Generator().manual_seed(11)
    frames = torch.rand(n_frames, 3, cfg["crop_h"], cfg["crop_w"], generator=g)
    if arch == "hnerv":
        with torch.no_grad():
            cali = model.encode(frames) * 3.0
    else:
        with torch.no_grad():
            cali = model.encode(torch.arange(n_frames).float() / n_frames)
    out = {"frames": npy(frames), "cali": npy(cali), "bits": np.array(bits), "hadamard": np.array(hadamard)}
    for k, v in model.state_dict().items():
        if "encoder" not in k:
            out["sd/" + k] = npy(v)
    with torch.no_grad():
        fp_out, emb, _ = model.decode(cali[:bsz])
    out["fp_out"] = npy(fp_out)
    for i, e in enumerate(emb):
        out[f"fp_embed{i}"] = npy(e)

    fp_model = copy.deepcopy(model)
    qnn = quantization.QuantModel(model=model, hadamard=hadamard,
                                  weight_quant_params={"n_bits": 8, "channel_wise": channel_wise, "scale_method": "max"})
    out["channel_wise"] = np.array(channel_wise)
    out["avg_bits"] = np.array(qnn.set_bitwidth(bits), dtype=np.float64)
    qnn.eval()
    qnn.set_quant_state(True)
    with torch.no_grad():
        q_out, _, _ = qnn(cali[:bsz])
    out["uaq_out"] = npy(q_out)
    mods = [m for m in qnn.model.modules() if isinstance(m, QuantModule)]
    for i, m in enumerate(mods):
        out[f"init/{i}/delta_w"] = npy(m.weight_quantizer.delta)
        out[f"init/{i}/zp_w"] = npy(m.weight_quantizer.zero_point)
        out[f"init/{i}/delta_b"] = npy(m.bias_quantizer.delta)
        out[f"init/{i}/zp_b"] = npy(m.bias_quantizer.zero_point)
    pert = qnn.get_perturbation()
    for i, v in enumerate(pert):
        out[f"pert/{i}"] = npy(v)

    if omega:
        # Omega on the FP model (bit_assign.py:171-203), CPU: neutralise the hard-coded .cuda()
        import importlib.util
        spec = importlib.util.spec_from_file_location("ref_bit_assign", "/root/reference/methods/bit_assign.py")
        ba = importlib.util.module_from_spec(spec)
        _cuda = torch.Tensor.cuda
        torch.Tensor.cuda = lambda self, *a, **k: self
        try:
            spec.loader.exec_module(ba)
            loader = [{"img": frames[i:i + bsz], "norm_idx": torch.arange(i, i + bsz).float() / n_frames,
                       "idx": torch.arange(i, i + bsz)} for i in range(0, n_frames, bsz)]
            net_o, net_f = copy.deepcopy(fp_model), copy.deepcopy(fp_model)
            om = ba.sensitivity_criterion("omega", arch, net_o, qnn, loader, use_cuda=False)
            out["omega"] = np.array(float(om), dtype=np.float64)
            fd = ba.sensitivity_criterion("fisher_diag", arch, net_f, qnn, loader, use_cuda=False)
            out["fisher_diag"] = np.array(float(fd), dtype=np.float64)
            # the per-layer terms the reference logs (bit_assign.py:194-200, :208-214), at full precision: H v and the
            # accumulated gradient are still in the nets' .grad (gradtensor_to_vec, :38-55)
            side = {"omega": out["omega"], "fisher_diag": out["fisher_diag"],
                    "omega_layers": np.array([float((gg * v).sum()) for gg, v in zip(ba.gradtensor_to_vec(net_o), pert)]),
                    "fisher_layers": np.array([float((v.pow(2) * gg.pow(2)).sum()) for gg, v in zip(ba.gradtensor_to_vec(net_f), pert)])}
            np.savez_compressed(os.path.join(HERE, f"{tag}_sens_layers.npz"), **side)
            print(tag, "omega layers", side["omega_layers"], "sum", side["omega_layers"].sum(), "omega", float(om))
        finally:
            torch.Tensor.cuda = _cuda

    if only_layers:
        return
    # calibration with an injected, fixed batch order
    order = [[(2 * j + 3 * k) % n_frames for k in range(bsz)] for j in range(n_frames // bsz)]
    order = [[0, 5], [3, 6], [1, 4], [7, 2]][: n_frames // bsz]
    out["order"] = np.array(order)
    loader = ListLoader([{"img": frames[torch.tensor(ix)], "norm_idx": torch.tensor(ix).float() / n_frames,
                          "idx": torch.tensor(ix)} for ix in order])
    import logging
    logging.getLogger().setLevel(logging.WARNING)
    # record the loss trajectory by wrapping the reference LossFunction
    import quantization.calib_model as cm
    traj = []
    _call = cm.LossFunction.__call__

    def rec_call(self, pred, tgt, grad=None):
        tot = _call(self, pred, tgt, grad)
        traj.append((self.count, float(tot), float(lp_loss(pred, tgt, p=self.p)), float(self.round_loss)))
        return tot

    cm.LossFunction.__call__ = rec_call
    try:
        quantization.model_reconstruction(qnn, cali_data=cali, gt=loader, arch=arch, batch_size=bsz,
                                          iters=iters, weight=0.01, opt_mode="mse", hadamard=hadamard,
                                          b_range=(20, 2), warmup=0.2, p=2.0, lr=0.003)
    finally:
        cm.LossFunction.__call__ = _call
    out["traj"] = np.array(traj, dtype=np.float64)
    with torch.no_grad():
        c_out, _, _ = qnn(cali[:bsz])
    out["calib_out"] = npy(c_out)
    out["calib_psnr"] = npy(psnr_fn_single(c_out, frames[:bsz]))
    codes = qnn.get_quantized_param()
    for i, m in enumerate(mods):
        out[f"final/{i}/delta_w"] = npy(m.weight_quantizer.delta)
        out[f"final/{i}/zp_w"] = npy(m.weight_quantizer.zero_point)
        out[f"final/{i}/alpha_w"] = npy(m.weight_quantizer.alpha)
        out[f"final/{i}/delta_b"] = npy(m.bias_quantizer.delta)
        out[f"final/{i}/zp_b"] = npy(m.bias_quantizer.zero_point)
        out[f"final/{i}/alpha_b"] = npy(m.bias_quantizer.alpha)
        out[f"final/{i}/codes_w"] = npy(codes[2 * i])
        out[f"final/{i}/codes_b"] = npy(codes[2 * i + 1])
    np.savez_compressed(os.path.join(HERE, f"{tag}.npz"), **out)
    print(tag, "avg_bits", float(out["avg_bits"]), "traj", len(traj), "psnr", out["calib_psnr"],
          "omega", out.get("omega"))


def bookkeeping():
    """Data-independent numbers the reference logs: average bit-widths at full size."""
    out = {}
    for arch, path, cases in (
        ("hnerv", "/root/reference/configs/HNeRV/Bunny_1280x640_3M.yaml", ([6, 5, 4, 5, 5, 6, 6], [2, 3, 4, 6, 4, 4, 2], [6] * 7)),
        ("nerv", "/root/reference/configs/NeRV/Bunny_1280x640_3M.yaml", ([6, 5, 4, 5, 5, 6, 6],)),
    ):
        cfg = get_config(path)
        torch.manual_seed(903)
        model = (models.HNeRV if arch == "hnerv" else models.NeRV)(cfg)
        qnn = quantization.QuantModel(model=model, hadamard=False,
                                      weight_quant_params={"n_bits": 8, "channel_wise": True, "scale_method": "max"})
        for bits in cases:
            out[f"{arch}_" + "".join(map(str, bits))] = np.array(qnn.set_bitwidth(bits), dtype=np.float64)
        shapes = [tuple(m.weight.shape) for m in qnn.model.modules() if isinstance(m, QuantModule)]
        out[f"{arch}_shapes"] = np.array(shapes)
    np.savez_compressed(os.path.join(HERE, "bookkeeping.npz"), **out)
    print({k: v.tolist() for k, v in out.items()})


if __name__ == "__main__":
    if not sys.argv[1:]:
        quantizer_kats()
        bookkeeping()
    if sys.argv[1:] == ["layerwise"]:  # per-tensor scales (the command line without --channel_wise): tiny_hnerv_lw.npz only
        run_model_case("tiny_hnerv_lw", "hnerv", TINY_HNERV, [6, 5, 4, 5, 5, 6, 6], False, iters=80, omega=False, channel_wise=False)
        sys.exit(0)
    if sys.argv[1:] == ["sens_layers"]:  # only the per-layer sensitivity side files (tiny_*_sens_layers.npz)
        run_model_case("tiny_hnerv", "hnerv", TINY_HNERV, [6, 5, 4, 5, 5, 6, 6], False, iters=80, only_layers=True)
        run_model_case("tiny_hnerv_had", "hnerv", TINY_HNERV, [4, 5, 4, 6, 5, 6, 8], True, iters=80, only_layers=True)
        run_model_case("tiny_nerv", "nerv", TINY_NERV, [6, 5, 4, 5, 5, 6, 6], False, iters=80, only_layers=True)
        sys.exit(0)
    run_model_case("tiny_hnerv", "hnerv", TINY_HNERV, [6, 5, 4, 5, 5, 6, 6], False, iters=80)
    run_model_case("tiny_hnerv_had", "hnerv", TINY_HNERV, [4, 5, 4, 6, 5, 6, 8], True, iters=80)
    run_model_case("tiny_nerv", "nerv", TINY_NERV, [6, 5, 4, 5, 5, 6, 6], False, iters=80)
    run_model_case("tiny_nerv_had", "nerv", TINY_NERV, [5, 6, 3, 4, 5, 4, 3], True, iters=80, omega=False)
