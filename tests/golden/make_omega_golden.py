"""Omega of several bit-width configurations from the UNMODIFIED reference's sensitivity_criterion
(methods/bit_assign.py:171-203, double-backward Hessian-vector products), on the tiny models of make_golden.py:

    python tests/golden/make_omega_golden.py      # writes tests/golden/tiny_{hnerv,nerv}_omega_configs.npz

The fixture pins the Gram-table evaluation of Omega (neuroquant_b200.sensitivity.OmegaTable + nq_omega_search): the
table must reproduce the reference's score of ARBITRARY configurations, not only of the two it was built to rank.  The
models are rebuilt exactly as make_golden.run_model_case builds them and checked against the weights stored in
tiny_*.npz."""
import copy
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (puts the shims and /root/reference on sys.path, imports the reference)

CONFIGS = [[2, 2, 2, 2, 2, 2, 2], [8, 8, 8, 8, 8, 8, 8], [6, 5, 4, 5, 5, 6, 6], [2, 3, 4, 6, 4, 4, 2], [3, 7, 2, 8, 5, 4, 6],
           [8, 2, 8, 2, 8, 2, 8], [4, 4, 4, 4, 4, 4, 4], [5, 6, 3, 4, 5, 4, 3], [7, 3, 5, 2, 6, 8, 4]]


def run(tag, arch, cfg, n_frames=8, bsz=2):
    models, quantization = mg.models, mg.quantization
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
    stored = np.load(os.path.join(HERE, tag + ".npz"))
    for k, v in model.state_dict().items():
        if "encoder" not in k:
            assert np.array_equal(stored["sd/" + k], v.numpy()), k
    assert np.array_equal(stored["frames"], frames.numpy()) and np.array_equal(stored["cali"], cali.numpy())
    spec = importlib.util.spec_from_file_location("ref_bit_assign", "/root/reference/methods/bit_assign.py")
    ba = importlib.util.module_from_spec(spec)
    _cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self   # bit_assign.py:110 hard-codes .cuda()
    out = {"configs": np.array(CONFIGS), "omega": np.zeros(len(CONFIGS))}
    try:
        spec.loader.exec_module(ba)
        loader = [{"img": frames[i:i + bsz], "norm_idx": torch.arange(i, i + bsz).float() / n_frames, "idx": torch.arange(i, i + bsz)}
                  for i in range(0, n_frames, bsz)]
        for ci, bits in enumerate(CONFIGS):
            qnn = quantization.QuantModel(model=copy.deepcopy(model), hadamard=False,
                                          weight_quant_params={"n_bits": 8, "channel_wise": True, "scale_method": "max"})
            qnn.set_bitwidth(bits)
            qnn.eval()
            qnn.set_quant_state(True)
            with torch.no_grad():
                qnn(cali[:bsz])
            om = ba.sensitivity_criterion("omega", arch, copy.deepcopy(model), qnn, loader, use_cuda=False)
            out["omega"][ci] = float(om)
            print(tag, bits, float(om), flush=True)
    finally:
        torch.Tensor.cuda = _cuda
    np.savez_compressed(os.path.join(HERE, f"{tag}_omega_configs.npz"), **out)


if __name__ == "__main__":
    import logging
    logging.getLogger().setLevel(logging.WARNING)
    run("tiny_hnerv", "hnerv", mg.TINY_HNERV)
    run("tiny_nerv", "nerv", mg.TINY_NERV)
