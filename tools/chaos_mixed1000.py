"""How well defined is the end point of the two-phase calibration the reference fixture `fullsize_mixed1000.npz` records
(HNeRV-Bunny-3M, W 6 5 4 5 5 6 6, 50 step-size + 950 AdaRound iterations, tests/golden/make_fullsize_golden.py)?

Runs the product's model_reconstruction on the fixture's seeded inputs four times -- tensor-core engine, exact-fp32 FFMA
engine, and each again with the embeddings perturbed by one part in 1e7 -- and prints the PSNR of the hard-rounded decode
after calibration next to the unmodified reference's.  The spread between the four runs is the yardstick for the
run-to-reference difference: the step-size gradient is a difference of two large sums and Adam's first steps are
sign-like, so the trajectories separate from the fourth iteration on (losses 2.2941e-07, 4.775e-04, 3.2314e-05 agree with
the reference's to four digits, then 6.26e-06 against 6.21e-06, 9.7e-07 against 5.9e-07).  GPU only.

    python tools/chaos_mixed1000.py > profiles/rNN_chaos_mixed1000.json
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(conv, perturb):
    os.environ["NQ_CONV"] = conv
    import numpy as np
    import torch
    from neuroquant_b200.models import HNeRV
    from neuroquant_b200.quantization import QuantModel, model_reconstruction
    from neuroquant_b200.utils import psnr_fn_single
    from tests.test_gpu_fullsize_oracle import ListLoader, _fixture_inputs
    g = np.load(os.path.join(ROOT, "tests", "golden", "fullsize_mixed1000.npz"))
    cfg, params, embeds, frames = _fixture_inputs(g)
    if perturb:
        embeds = embeds * (1.0 + perturb * torch.randn(embeds.shape, generator=torch.Generator().manual_seed(5)))
    torch.manual_seed(1)
    model = HNeRV(dict(cfg))
    convs = [model.decoder[0]] + [blk.conv[0] for blk in list(model.decoder)[1:]] + [model.head_layer]
    with torch.no_grad():
        for c, (w, b) in zip(convs, params):
            c.weight.copy_(w)
            c.bias.copy_(b)
    model = model.cuda()
    embeds_d, frames_d = embeds.cuda(), frames.cuda()
    qnn = QuantModel(model, hadamard=False, weight_quant_params={"n_bits": 8, "channel_wise": True, "scale_method": "max"}).cuda()
    qnn.set_bitwidth(g["bits"].tolist())
    qnn.eval()
    qnn.set_quant_state(True)
    qnn(embeds_d[:2])
    loader = ListLoader([{"img": frames[torch.tensor(ix)], "norm_idx": torch.tensor(ix).float() / 20, "idx": torch.tensor(ix)}
                         for ix in g["order"].tolist()])
    losses = []
    model_reconstruction(qnn, cali_data=embeds_d, gt=loader, arch="hnerv", batch_size=2, iters=int(g["iters"]), weight=0.01, opt_mode="mse",
                         hadamard=False, b_range=(20, 2), warmup=0.2, p=2.0, lr=0.003,
                         on_iteration=lambda phase, count, loss: losses.append(loss.clone()))
    with torch.no_grad():
        ps = torch.cat([psnr_fn_single(qnn(embeds_d[i:i + 2])[0], frames_d[i:i + 2]) for i in range(0, embeds.shape[0], 2)])
    rec = torch.stack(losses).view(-1).cpu().double().numpy()
    print(json.dumps({"engine": conv, "perturb": perturb, "psnr_calibrated": float(ps.double().mean()), "loss_it50": float(rec[49]),
                      "loss_last100_mean": float(rec[-100:].mean()), "psnr_reference": float(g["psnr_calibrated"].mean()),
                      "psnr_nearest_reference": float(g["psnr_nearest"].mean()), "loss_it50_reference": float(g["traj"][49, 2]),
                      "loss_last100_mean_reference": float(g["traj"][-100:, 2].mean())}))


if __name__ == "__main__":
    if len(sys.argv) == 3:
        one(sys.argv[1], float(sys.argv[2]))
    else:
        for conv in ("tc", "simt"):
            for pt in (0.0, 1e-7):
                subprocess.run([sys.executable, os.path.abspath(__file__), conv, str(pt)], check=True, cwd=ROOT)
