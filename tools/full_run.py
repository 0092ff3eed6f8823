#!/usr/bin/env python
"""The north-star job, end to end through the command lines, with its record written to a JSON file:

    python tools/full_run.py --arch hnerv --out gpurun_out/r02_full_run_hnerv.json            # 1 GPU
    torchrun --nproc-per-node 2 tools/full_run.py --arch hnerv --skip-train --work /tmp/nq_full ...   # data parallel

  1. a synthetic 132-frame 1280x720 clip (smooth moving pattern, uint8 PNGs; the calibration crops 640x1280 out of it as
     the reference does with the Bunny frames) and the reference's YAML for the 3M model;
  2. `methods/regress.py` for a few epochs: a full-precision checkpoint with non-degenerate weights (the reference's
     epoch300.pth is not in its repository);
  3. `methods/calibrate_network.py --iters_w 21000 --precision 6 5 4 5 5 6 6 --channel_wise --batch_size 2 ...`: the
     reference's documented command (readme.md:103-108), i.e. 990 step-size + 19 998 AdaRound iterations;
  4. the record: wall time, the reference's `count=` log lines (b = 19.68 @ 4500, 3.61 @ 19500), PSNR full precision ->
     nearest rounding -> calibrated, the checkpoint reloaded (whole-object pickle) and decoding identically, and the final
     integer codes re-derived bit-exactly by the ORACLE quantiser from the saved V (alpha) and step sizes.
"""
import argparse
import glob
import json
import os
import re
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import yaml  # noqa: E402

CFGS = {
    "hnerv": dict(crop_h=640, crop_w=1280, diff_enc=False, stage_block=1, enc_strides=[5, 4, 4, 2, 2],
                  enc_channel=[64, 64, 64, 64, 16], channel_reduce=1.2, channel_lbound=12, dec_in_channel=92,
                  dec_kernels=[1, 3, 5, 5, 5], dec_strides=[5, 4, 4, 2, 2], dec_norm="none", dec_acts="gelu", out_bias="tanh",
                  loss="l2", epoch=300, workers=0, eval_freq=30, batch_size=1, learning_rate=0.0005),
    "nerv": dict(crop_h=640, crop_w=1280, diff_enc=False, base=1.25, level=80, channel_reduce=2, channel_lbound=24,
                 dec_in_channel=145, dec_kernels=[3, 3, 3, 3, 3], dec_strides=[5, 4, 4, 2, 2], dec_norm="none", dec_acts="gelu",
                 out_bias="tanh", loss="l2", epoch=300, workers=0, eval_freq=30, batch_size=1, learning_rate=0.0005),
}


def make_clip(path, n_frames=132, h=720, w=1280):
    """Smooth synthetic video: three drifting sinusoid fields per channel + two moving soft discs."""
    from torchvision.io import write_png
    os.makedirs(path, exist_ok=True)
    dev = "cuda"
    yy, xx = torch.meshgrid(torch.linspace(0, 1, h, device=dev), torch.linspace(0, 1, w, device=dev), indexing="ij")
    for t in range(n_frames):
        s = t / n_frames
        chans = []
        for c in range(3):
            f = 0.5 + 0.22 * torch.sin(6.0 * xx * (1 + 0.3 * c) + 4.0 * yy + 6.28 * s + c) \
                + 0.15 * torch.sin(17.0 * yy * (1 + 0.2 * c) - 9.0 * xx + 12.56 * s) \
                + 0.08 * torch.sin(40.0 * (xx + yy) + 3.0 * c + 18.8 * s)
            for k, (cx, cy, r) in enumerate(((0.2 + 0.6 * s, 0.4, 0.12), (0.7 - 0.4 * s, 0.65, 0.08))):
                d2 = ((xx - cx) * w / h) ** 2 + (yy - cy) ** 2
                f = f + (0.25 if (c + k) % 2 == 0 else -0.2) * torch.exp(-d2 / (2 * r * r))
            chans.append(f)
        img = (torch.stack(chans).clamp(0, 1) * 255).round().to(torch.uint8).cpu()
        write_png(img, os.path.join(path, f"{t + 1:04d}.png"))


def parse_log(path):
    text = open(path).read()
    counts = []
    for m in re.finditer(r"Total loss:\s+([\d.eE+-]+) \(rec:([\d.eE+-]+), round:([\d.eE+-]+)\)\s+b=([\d.]+)\s+count=(\d+)", text):
        counts.append({"count": int(m.group(5)), "total": float(m.group(1)), "rec": float(m.group(2)), "round": float(m.group(3)),
                       "b": float(m.group(4))})
    psnr = [float(m.group(1)) for m in re.finditer(r"best_pred_seen_psnr: ([\d.]+)", text)]
    fps = [float(m.group(1)) for m in re.finditer(r"FPS ([\d.]+)", text)]
    took = re.search(r"Training complete in: (\S+)", text)
    avg = re.search(r"qbits: ([\d.]+)", text)
    return {"count_lines": counts, "psnr_fp_quantoff_nearest_calibrated": psnr, "decode_fps_log": fps[-1] if fps else None,
            "training_complete_in": took.group(1) if took else None, "avg_bits": float(avg.group(1)) if avg else None}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="hnerv", choices=["hnerv", "nerv"])
    ap.add_argument("--work", default="/tmp/nq_full")
    ap.add_argument("--out", default="gpurun_out/r02_full_run.json")
    ap.add_argument("--iters", type=int, default=21000)
    ap.add_argument("--fp-epochs", type=int, default=20)
    ap.add_argument("--batch", type=int, default=2, help="global mini-batch (the reference's --batch_size)")
    ap.add_argument("--precision", type=int, nargs="+", default=[6, 5, 4, 5, 5, 6, 6])
    ap.add_argument("--hadamard", action="store_true")
    ap.add_argument("--skip-train", action="store_true", help="reuse the clip and checkpoint under --work")
    ap.add_argument("--train-only", action="store_true", help="make the clip and the full-precision checkpoint, then stop")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    from neuroquant_b200.methods import calibrate_network, regress

    work = os.path.join(a.work, a.arch)
    clip, cfg_path = os.path.join(a.work, "clip"), os.path.join(work, "cfg.yaml")
    os.makedirs(work, exist_ok=True)
    rec = {"arch": a.arch, "gpus": world, "iters_w": a.iters, "precision": a.precision, "hadamard": a.hadamard, "batch_size": a.batch}
    os.chdir(work)
    if not a.skip_train and rank == 0 and world > 1:
        # the regression script is single-GPU (as in the reference): run it in a child process outside the process group
        import subprocess
        env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "LOCAL_WORLD_SIZE", "GROUP_RANK",
                                                                  "MASTER_ADDR", "MASTER_PORT", "TORCHELASTIC_RUN_ID")}
        t0 = time.time()
        subprocess.run([sys.executable, os.path.abspath(__file__), "--arch", a.arch, "--work", a.work, "--fp-epochs", str(a.fp_epochs),
                        "--train-only"], check=True, env=env, cwd=ROOT)
        rec["fp_training_seconds"] = time.time() - t0
    elif not a.skip_train and rank == 0:
        t0 = time.time()
        if not os.path.isdir(clip) or len(os.listdir(clip)) != 132:
            make_clip(clip)
        rec["clip_seconds"] = time.time() - t0
        cfg = dict(CFGS[a.arch], epoch=a.fp_epochs, eval_freq=max(1, a.fp_epochs // 2))
        yaml.safe_dump(cfg, open(cfg_path, "w"))
        t0 = time.time()
        regress.main(["--config", cfg_path, "--arch", a.arch, "--data_path", clip, "--vid", "Synth", "--outf", "fp", "-p", "1000"])
        rec["fp_training_seconds"] = time.time() - t0
    if a.train_only:
        return
    if world > 1:
        from neuroquant_b200.methods.common import init_distributed
        init_distributed()
        torch.distributed.barrier()
    ckpt = sorted(glob.glob(os.path.join(work, "results", "fp", "**", "epoch*.pth"), recursive=True))[-1]
    # the calibration command of readme.md:103-108 (the YAML's own epoch / lr only name the output directory)
    argv = ["--config", cfg_path, "--arch", a.arch, "--data_path", clip, "--vid", "Synth", "--batch_size", str(a.batch),
            "--precision", *[str(b) for b in a.precision], "--channel_wise", "--iters_w", str(a.iters), "--weight", "0.01",
            "--b_start", "20", "--b_end", "2", "--warmup", "0.2", "--lr", "0.003", "--ckpt", ckpt, "--outf", f"calib{world}", "-p", "1000"]
    if a.hadamard:
        argv.append("--hadamard")
    torch.cuda.synchronize()
    t0 = time.time()
    calibrate_network.main(argv)
    torch.cuda.synchronize()
    rec["calibrate_network_wall_seconds"] = time.time() - t0
    if rank != 0:
        return
    files = sorted(glob.glob(os.path.join(work, "results", f"calib{world}", "**", f"{a.arch}_W*.pth"), recursive=True), key=os.path.getmtime)
    logs = sorted(glob.glob(os.path.join(os.path.dirname(files[-1]), "*.log")), key=os.path.getmtime)
    rec.update(parse_log(logs[-1]))
    rec["checkpoint"] = os.path.relpath(files[-1], work)
    rec["checkpoint_bytes"] = os.path.getsize(files[-1])
    took = rec.get("training_complete_in")
    if took:
        hh, mm, ss = took.split(":")
        sec = int(hh) * 3600 + int(mm) * 60 + float(ss)
        n_iter = 0
        n_b = 132 // a.batch
        ep1 = int(0.05 * a.iters / n_b)
        n_iter = int(a.iters / n_b) * n_b
        rec.update(model_reconstruction_seconds=sec, iterations_executed=n_iter, step_size_iterations=ep1 * n_b,
                   iters_per_s=n_iter / sec, frames_per_s=n_iter * a.batch / sec)
    # ---- the checkpoint: reload (whole-object pickle under the reference's module paths), decode, re-derive the codes
    from neuroquant_b200.quantization import QuantModule
    from oracle import nq_oracle as O
    qnn = torch.load(files[-1], weights_only=False).cuda()
    rec["pickle_module"] = type(qnn).__module__
    mods = [m for m in qnn.model.modules() if isinstance(m, QuantModule)]
    arch_cfg = CFGS[a.arch]
    if a.arch == "hnerv":
        embed = torch.randn(2, 16, 2, 4, device="cuda", generator=torch.Generator("cuda").manual_seed(1))
    else:
        embed = qnn.model.encode(torch.tensor([0.0, 0.5], device="cuda"))
    out1, _, _ = qnn(embed)
    out2, _, _ = qnn(embed)
    rec["reloaded_decode_deterministic"] = bool(torch.equal(out1, out2))
    exact, undecided, n_codes = True, 0, 0
    for m in mods:
        wq = m.weight_quantizer
        src = m.hadamard_weight if m.hadamard else m.org_weight
        want, _ = O.adaround_quant(src.cpu(), wq.alpha.detach().cpu(), wq.delta.detach().cpu(), wq.zero_point.cpu(), wq.n_bits, False)
        got = wq.x_quant.cpu()
        exact = exact and bool(torch.equal(got, want)) and bool(torch.equal(got, got.round()))
        h = O.soft_targets(wq.alpha.detach().cpu())
        undecided += int(((h > 0) & (h < 1)).sum())
        n_codes += got.numel()
        assert not wq.soft_targets and m.bias_quantizer.soft_targets   # calib_model.py:231-240
    rec.update(codes_bit_exact_with_oracle_quantiser=exact, weight_codes=n_codes, soft_targets_still_undecided=undecided)
    os.chdir(ROOT)
    os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
    json.dump(rec, open(a.out, "w"), indent=1)
    print(json.dumps({k: v for k, v in rec.items() if k != "count_lines"}))
    print("count lines:", [(c["count"], c["b"], round(c["rec"], 6)) for c in rec["count_lines"][:3]], "...",
          [(c["count"], c["b"]) for c in rec["count_lines"] if c["count"] in (4500, 19500)])


if __name__ == "__main__":
    main()
