"""Host-side logic that needs no GPU: the learning-rate schedule, the prefix quantisation of the Fisher gradient pass,
the runner's per-stage state, the packed-stream size arithmetic."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import nq_oracle as O
from tests.helpers import TINY_HNERV, load


def test_adjust_lr_matches_reference_sequences():
    """utils.adjust_lr (product) against the learning rates the reference's adjust_lr produced (regress fixtures) and
    against the oracle's restatement, for the cosine and the hybrid schedule."""
    from neuroquant_b200.utils import adjust_lr
    for tag in ("regress_tiny_nerv", "regress_tiny_nerv_l1"):
        g = load(tag)
        args = SimpleNamespace(lr=float(g["lr"]), lr_type=str(g["lr_type"]))
        opt = SimpleNamespace(param_groups=[{"lr": 0.0}, {"lr": 0.0}])
        epochs, n = int(g["epochs"]), len(g["lr_seq"])
        per = n // epochs
        got = []
        for it in range(n):
            e, i = divmod(it, per)
            got.append(adjust_lr(opt, (e + float(i) / per) / epochs, args))
            assert opt.param_groups[0]["lr"] == opt.param_groups[1]["lr"] == got[-1]
            assert got[-1] == O.adjust_lr(args.lr, (e + float(i) / per) / epochs, args.lr_type)
        assert np.allclose(got, g["lr_seq"], rtol=1e-12)
    with pytest.raises(NotImplementedError):
        adjust_lr(SimpleNamespace(param_groups=[]), 0.5, SimpleNamespace(lr=1.0, lr_type="step_0.1"))


def test_quantize_model_till_quantises_a_prefix():
    """data_utils.py:261-272: every layer / block up to and including the given one, in module order."""
    from neuroquant_b200.models import HNeRV
    from neuroquant_b200.quantization import QuantModel, QuantModule
    from neuroquant_b200.quantization.calib_block import quantize_model_till
    torch.manual_seed(0)
    qnn = QuantModel(HNeRV(TINY_HNERV), hadamard=False, weight_quant_params={"n_bits": 8, "channel_wise": True, "scale_method": "max"})
    mods = [m for m in qnn.model.modules() if isinstance(m, QuantModule)]
    assert len(mods) == 7
    for target, n_on in ((qnn.model.decoder[3], 4), (qnn.model.decoder[3].conv, 4), (qnn.model.decoder[0], 1), (qnn.model.head_layer, 7)):
        qnn.set_quant_state(True)
        quantize_model_till(qnn, target)
        assert [m.use_weight_quant for m in mods] == [True] * n_on + [False] * (7 - n_on)


def test_layer_and_block_entry_points_reject_wrong_modules():
    from neuroquant_b200.models import HNeRV
    from neuroquant_b200.quantization import QuantModel, block_reconstruction, layer_reconstruction
    qnn = QuantModel(HNeRV(TINY_HNERV), hadamard=False, weight_quant_params={"n_bits": 8, "channel_wise": True, "scale_method": "max"})
    cali = torch.zeros(10, 4, 2, 4)
    with pytest.raises(ValueError):
        block_reconstruction(qnn, qnn.model.head_layer, cali)          # a lone layer is not a block
    with pytest.raises(ValueError):
        layer_reconstruction(qnn, qnn.model.decoder[2], cali)          # a block is not a layer
    with pytest.raises(ValueError):
        block_reconstruction(qnn, qnn.model.decoder[2], cali, opt_mode="hessian")


def test_reference_module_paths_resolve_after_aliasing():
    """compat.install_reference_aliases: the reference's import paths, the block- and layer-wise modules included.  Run
    in a subprocess so that the aliases do not leak into this test session."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from neuroquant_b200.compat import install_reference_aliases\n"
            "install_reference_aliases()\n"
            "from quantization import QuantModel, model_reconstruction, block_reconstruction, layer_reconstruction\n"
            "from quantization.calib_layer import layer_reconstruction as a\n"
            "from quantization.calib_block import block_reconstruction as b\n"
            "from quantization.quantizer import AdaRoundQuantizer, UniformAffineQuantizer\n"
            "from models import HNeRV, NeRV\n"
            "assert a is layer_reconstruction and b is block_reconstruction\n"
            "assert QuantModel.__module__ == 'quantization.quant_model' and HNeRV.__module__ == 'models.HNeRV'\n"
            "print('aliases ok')\n") % root
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp", timeout=300)
    assert out.returncode == 0 and "aliases ok" in out.stdout, out.stderr[-2000:]


def test_oracle_workloads_agree_with_the_package():
    """bench.py's CPU legs take their configurations and seeded weights from oracle/workloads.py (so that the reference arm
    maps no product library); they must be the package's."""
    import torch
    from neuroquant_b200 import workloads as P
    from oracle import workloads as W
    assert set(P.WORKLOADS) == set(W.WORKLOADS)
    for name in P.WORKLOADS:
        assert P.WORKLOADS[name] == W.WORKLOADS[name]
        arch, cfg = P.WORKLOADS[name]
        assert P.embed_shape(cfg, arch) == W.embed_shape(cfg, arch)
        geoms, _ = (P.geometry_from_cfg(cfg, arch), None)
        _, h0, w0 = P.embed_shape(cfg, arch)
        assert P.conv_flops(geoms, h0, w0, 2) == W.conv_flops(cfg, arch, 2)
    arch, cfg = P.WORKLOADS["nerv-bunny-3m"]
    _, params = P.random_decoder(cfg, arch, 903)
    for (w, b), st in zip(params, W.random_stages(cfg, arch, 903)):
        assert torch.equal(w, st.weight) and torch.equal(b, st.bias)


def test_reference_staging_recipe_is_byte_exact():
    """oracle/make_ref.py copies the reference's product Python byte for byte (MANIFEST.json holds both hashes); the
    staged tree is what bench.py's `--impl reference` runs on the GPU box.  Skipped where nothing was staged."""
    import hashlib, json, os
    import pytest
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
    man = os.path.join(root, "MANIFEST.json")
    if not os.path.exists(man):
        pytest.skip("oracle/_ref not staged (no /root/reference at build time)")
    files = json.load(open(man))["files"]
    assert "quantization/calib_model.py" in files and "models/HNeRV.py" in files
    for rel, h in files.items():
        assert hashlib.sha256(open(os.path.join(root, rel), "rb").read()).hexdigest() == h["sha256"] == h["source_sha256"], rel
