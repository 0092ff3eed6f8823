"""Golden fixtures for FP32 regression training from the UNMODIFIED reference.

Run in the build container only:   python tests/golden/make_regress_golden.py
Writes tests/golden/regress_*.npz.  regress.train (methods/regress.py:151-322) needs a PNG data set, DataLoader workers
and TensorBoard; its inner loop (:249-271) is replayed here on seeded synthetic frames with the reference's OWN model
classes, loss_fn and adjust_lr and torch.optim.Adam(model.parameters(), weight_decay=0.) as regress.py:239 builds it.
The mini-batch order (the reference shuffles unseeded) is recorded.
"""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_shims"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, HERE)

import models  # noqa: E402  (reference)
from utils import adjust_lr, loss_fn, psnr_fn_single  # noqa: E402  (reference)
from make_golden import TINY_HNERV, TINY_NERV, npy  # noqa: E402

torch.set_num_threads(8)


def run_case(tag, arch, cfg, epochs=3, n_frames=6, bsz=2, lr=2e-3, lr_type="cosine_0.1_1_0.1", loss="l2"):
    torch.manual_seed(903)
    model = (models.HNeRV if arch == "hnerv" else models.NeRV)(cfg)
    g = torch.Generator().manual_seed(11)
    frames = torch.rand(n_frames, 3, cfg["crop_h"], cfg["crop_w"], generator=g)
    norm_idx = torch.arange(n_frames).float() / n_frames
    out = {"frames": npy(frames), "norm_idx": npy(norm_idx), "epochs": np.array(epochs), "bsz": np.array(bsz), "lr": np.array(lr),
           "lr_type": np.array(lr_type), "loss_type": np.array(loss)}
    for k, v in model.state_dict().items():
        out["sd0/" + k] = npy(v)
    if arch == "nerv":
        with torch.no_grad():
            out["embed"] = npy(model.encode(norm_idx))  # fixed positional encoding: the decoder is all that trains
    args = SimpleNamespace(lr=lr, lr_type=lr_type)
    optimizer = torch.optim.Adam(model.parameters(), weight_decay=0.)  # regress.py:239
    order, losses, lrs, psnrs = [], [], [], []
    model.train()
    for epoch in range(epochs):
        perm = torch.randperm(n_frames, generator=g)
        batches = [perm[i:i + bsz] for i in range(0, n_frames - bsz + 1, bsz)]  # drop_last=True (regress.py:169)
        for i, idx in enumerate(batches):
            cur_epoch = (epoch + float(i) / len(batches)) / epochs           # regress.py:252
            cur_lr = adjust_lr(optimizer, cur_epoch, args)
            img = frames[idx]
            img_out, _, _ = model(img) if arch == "hnerv" else model(norm_idx[idx])
            final_loss = loss_fn(img_out, img, loss)
            optimizer.zero_grad()
            final_loss.backward()
            optimizer.step()
            order.append(npy(idx)); losses.append(float(final_loss)); lrs.append(cur_lr)
            psnrs.append(npy(psnr_fn_single(img_out.detach(), img)))
    out["order"], out["loss"], out["lr_seq"], out["psnr"] = np.stack(order), np.array(losses), np.array(lrs), np.stack(psnrs)
    for k, v in model.state_dict().items():
        out["sd1/" + k] = npy(v)
    np.savez_compressed(os.path.join(HERE, f"{tag}.npz"), **out)
    print(tag, "steps", len(losses), "loss", losses[0], "->", losses[-1], "lr", lrs[0], lrs[-1])


if __name__ == "__main__":
    run_case("regress_tiny_nerv", "nerv", TINY_NERV)
    run_case("regress_tiny_hnerv", "hnerv", TINY_HNERV)
    run_case("regress_tiny_nerv_l1", "nerv", TINY_NERV, loss="l1", lr_type="hybrid_0.2_1_2_0.1_0.05")
