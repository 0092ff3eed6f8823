"""End-to-end run of the two command lines on a tiny synthetic clip (GPU)."""
import glob
import os

import pytest
import torch
import yaml

from tests.helpers import TINY_HNERV, TINY_NERV

pytestmark = pytest.mark.gpu


def make_clip(tmp_path, cfg, arch, n_frames=8):
    from torchvision.io import write_png
    from neuroquant_b200.models import HNeRV, NeRV
    data = tmp_path / "clip"
    data.mkdir()
    g = torch.Generator().manual_seed(3)
    base = torch.rand(3, cfg["crop_h"] + 8, cfg["crop_w"] + 8, generator=g)
    for i in range(n_frames):
        frame = (base.roll(i, 2) * 255).to(torch.uint8)
        write_png(frame, str(data / f"{i + 1:04d}.png"))
    full = dict(cfg, loss="l2", epoch=3, workers=0, eval_freq=1, batch_size=1, learning_rate=0.0005)
    cfg_path = tmp_path / "cfg.yaml"
    cfg_path.write_text(yaml.safe_dump(full))
    torch.manual_seed(11)
    model = (HNeRV if arch == "hnerv" else NeRV)(full)
    ckpt = tmp_path / "fp.pth"
    torch.save(model.state_dict(), str(ckpt))
    return str(data), str(cfg_path), str(ckpt)


@pytest.mark.parametrize("arch,cfg,hadamard", [("hnerv", TINY_HNERV, False), ("nerv", TINY_NERV, True)])
def test_calibrate_network_cli(tmp_path, monkeypatch, arch, cfg, hadamard):
    from neuroquant_b200.methods import calibrate_network
    data, cfg_path, ckpt = make_clip(tmp_path, cfg, arch)
    monkeypatch.chdir(tmp_path)
    argv = ["--config", cfg_path, "--arch", arch, "--data_path", data, "--vid", "Clip", "--batch_size", "2",
            "--precision", "6", "5", "4", "5", "5", "6", "6", "--channel_wise", "--iters_w", "40", "--weight", "0.01",
            "--b_start", "20", "--b_end", "2", "--warmup", "0.2", "--lr", "0.003", "--ckpt", ckpt, "--outf", "t"]
    if hadamard:
        argv.append("--hadamard")
    calibrate_network.main(argv)
    files = glob.glob(os.path.join("results", "t", "**", f"{arch}_W*_prob1.0_max-init_CW.pth"), recursive=True)
    assert len(files) == 1 and f"hadamard-{hadamard}_max-init_batch2_CW_weight0.01_brange20-2_warmup0.2_lr0.003" in files[0]
    qnn = torch.load(files[0], weights_only=False)
    assert type(qnn).__module__ == "quantization.quant_model"  # the reference's pickle path
    codes = qnn.get_quantized_param()
    assert all(torch.equal(c, c.round()) for c in codes[0::2])  # weight codes are integers after calibration
    logs = glob.glob(os.path.join(os.path.dirname(files[0]), "*.log"))
    text = open(logs[0]).read()
    assert "Weight quantization model w/ opt" in text and "Training complete in" in text


def test_bit_assign_cli(tmp_path, monkeypatch):
    from neuroquant_b200.methods import bit_assign
    data, cfg_path, ckpt = make_clip(tmp_path, TINY_HNERV, "hnerv")
    monkeypatch.chdir(tmp_path)
    for mode in ("omega", "fisher_diag"):
        best, bits, score = bit_assign.assign(*_args(bit_assign, cfg_path, data, ckpt, mode))
        assert best in bit_assign.hnerv_candidate and bits == bit_assign.hnerv_candidate[best] and score == score
        assert best == "candidate2"  # the higher-precision toy candidate perturbs less (as in the reference's logs)


def _args(mod, cfg_path, data, ckpt, mode):
    from neuroquant_b200.utils import get_config
    args = mod.parse_args(["--config", cfg_path, "--arch", "hnerv", "--data_path", data, "--vid", "Clip", "--batch_size", "2",
                           "--channel_wise", "--init", "max", "--mode", mode, "--ckpt", ckpt, "--outf", "t"])
    cfg = get_config(cfg_path)
    args.outf = os.path.join("results", args.outf, "x_" + mode)
    return args, cfg


def test_calibrate_network_cli_mse_init_layerwise_scales(tmp_path, monkeypatch):
    """--init mse without --channel_wise: per-tensor scales from the range search, checkpoint suffix LW."""
    from neuroquant_b200.methods import calibrate_network
    data, cfg_path, ckpt = make_clip(tmp_path, TINY_HNERV, "hnerv")
    monkeypatch.chdir(tmp_path)
    calibrate_network.main(["--config", cfg_path, "--arch", "hnerv", "--data_path", data, "--vid", "Clip", "--batch_size", "2",
                            "--precision", "6", "5", "4", "5", "5", "6", "6", "--init", "mse", "--iters_w", "40", "--weight", "0.01",
                            "--b_start", "20", "--b_end", "2", "--warmup", "0.2", "--lr", "0.003", "--ckpt", ckpt, "--outf", "t"])
    files = glob.glob(os.path.join("results", "t", "**", "hnerv_W*_prob1.0_mse-init_LW.pth"), recursive=True)
    assert len(files) == 1
    qnn = torch.load(files[0], weights_only=False)
    from neuroquant_b200.quantization import QuantModule
    mods = [m for m in qnn.modules() if isinstance(m, QuantModule)]
    assert all(m.weight_quantizer.delta.numel() == 1 for m in mods)  # one step size per tensor
    assert all(torch.equal(c, c.round()) for c in qnn.get_quantized_param()[0::2])


@pytest.mark.parametrize("arch,cfg", [("nerv", TINY_NERV), ("hnerv", TINY_HNERV)])
def test_regress_cli_trains_and_writes_reference_checkpoints(tmp_path, monkeypatch, arch, cfg):
    """methods/regress.py: three epochs on the tiny clip; the loss falls, the checkpoints are plain state_dicts with the
    reference's keys, and the calibration command line starts from them."""
    from neuroquant_b200.methods import regress
    data, cfg_path, _ = make_clip(tmp_path, cfg, arch)
    monkeypatch.chdir(tmp_path)
    regress.main(["--config", cfg_path, "--arch", arch, "--data_path", data, "--vid", "Clip", "--outf", "r", "-p", "1"])
    ck = glob.glob(os.path.join("results", "r", "**", "epoch3.pth"), recursive=True)
    latest = glob.glob(os.path.join("results", "r", "**", "model_latest.pth"), recursive=True)
    assert len(ck) == 1 and len(latest) == 1
    sd = torch.load(ck[0], map_location="cpu")
    assert "decoder.0.weight" in sd and "decoder.1.conv.0.weight" in sd and "head_layer.bias" in sd
    assert any(k.startswith("encoder.") for k in sd) == (arch == "hnerv")
    text = open(glob.glob(os.path.join(os.path.dirname(ck[0]), "*.log"))[0]).read()
    psnrs = [float(l.split("pred_PSNR:")[1].split()[0].strip(",")) for l in text.splitlines() if "pred_PSNR:" in l]
    assert len(psnrs) >= 6 and psnrs[-1] > psnrs[0]          # it learns
    assert "Training complete in" in text
