"""GPU parity of the FP32 regression step (neuroquant_b200.methods.regress.DecoderTrainer) against the
reference's own model classes, loss_fn, adjust_lr and torch Adam replayed on seeded frames
(tests/golden/make_regress_golden.py)."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from tests.helpers import TINY_HNERV, TINY_NERV, load, t

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag,arch,cfg", [("regress_tiny_nerv", "nerv", TINY_NERV), ("regress_tiny_hnerv", "hnerv", TINY_HNERV),
                                          ("regress_tiny_nerv_l1", "nerv", TINY_NERV)])
@pytest.mark.parametrize("conv", ["tc", "simt"])
def test_regression_training_matches_reference(tag, arch, cfg, conv, monkeypatch):
    monkeypatch.setenv("NQ_CONV", conv)
    from neuroquant_b200.methods.regress import DecoderTrainer
    from neuroquant_b200.models import HNeRV, NeRV
    from neuroquant_b200.utils import adjust_lr
    g = load(tag)
    model = (HNeRV if arch == "hnerv" else NeRV)(cfg)
    sd0 = {k[4:]: t(g[k]) for k in g.files if k.startswith("sd0/")}
    missing, unexpected = model.load_state_dict(sd0, strict=True)
    assert not missing and not unexpected
    model = model.cuda().train()
    frames, norm_idx = t(g["frames"]).cuda(), t(g["norm_idx"]).cuda()
    epochs, order = int(g["epochs"]), g["order"]
    per_epoch = len(order) // epochs
    args = SimpleNamespace(lr=float(g["lr"]), lr_type=str(g["lr_type"]))
    trainer = DecoderTrainer(model, arch, args.lr, str(g["loss_type"]) if "loss_type" in g.files else "l2")
    losses, lrs, psnrs = [], [], []
    for it, idx in enumerate(order):
        epoch, i = divmod(it, per_epoch)
        lrs.append(adjust_lr(trainer, (epoch + float(i) / per_epoch) / epochs, args))     # regress.py:252-253
        idx = torch.as_tensor(idx).cuda()
        img = frames[idx]
        loss, img_out = trainer.step(img if arch == "hnerv" else norm_idx[idx], img)
        losses.append(float(loss))
        psnrs.append((-10 * torch.log10(((img_out - img) ** 2).flatten(1).mean(1) + 1e-9)).cpu().numpy())
    assert trainer.launches > 0
    assert np.allclose(lrs, g["lr_seq"], rtol=1e-12)
    assert np.allclose(losses, g["loss"], rtol=2e-5), (losses, g["loss"].tolist())
    assert np.allclose(np.stack(psnrs), g["psnr"], atol=2e-3)
    # every trained tensor, decoder (engine kernels) and encoder (torch autograd fed by the engine's d_embed): Adam's
    # update is g / (|g| + 1e-8)-like for the first steps, so an entry whose gradient is at rounding level may move by up
    # to one learning rate differently -- bounded, and rare
    sd1 = {k[4:]: g[k] for k in g.files if k.startswith("sd1/")}
    now = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    assert set(now) == set(sd1)
    moved = 0.0
    for k, ref in sd1.items():
        diff = np.abs(now[k] - ref)
        moved = max(moved, float(np.abs(ref - g["sd0/" + k]).max()))
        assert diff.max() <= 3 * float(max(g["lr_seq"])), (k, diff.max())
        assert (diff > 2e-5).mean() < 0.01, (k, (diff > 2e-5).mean(), diff.max())
    assert moved > 1e-4  # the fixture does train
