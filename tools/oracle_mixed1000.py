"""CPU: the ORACLE (oracle/nq_oracle.py, the pinned restatement of the reference) on the inputs of
tests/golden/fullsize_mixed1000.npz -- 50 step-size + 950 AdaRound iterations at full size -- optionally with the embeddings
perturbed by one part in 1e7: how far does the reference's own arithmetic move under a rounding-level change of its input?
Prints one JSON line (PSNR of the hard-rounded decode, loss at iteration 50, mean loss of the last 100 iterations) next to
the unmodified reference's values from the fixture.  About 50 minutes on 4 threads.

    python tools/oracle_mixed1000.py <perturb> <threads> > profiles/rNN_oracle_mixed1000_<perturb>.json
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from neuroquant_b200.workloads import WORKLOADS, random_decoder  # noqa: E402  (host-side helpers only)
from oracle import nq_oracle as O  # noqa: E402

perturb = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0
torch.set_num_threads(int(sys.argv[2]) if len(sys.argv) > 2 else 4)
g = np.load(os.path.join(ROOT, "tests", "golden", "fullsize_mixed1000.npz"))
arch, cfg = WORKLOADS["hnerv-bunny-3m"]
geoms, params = random_decoder(cfg, arch, 903)
gen = torch.Generator().manual_seed(29)
n = int(g["order"].size)
embeds = torch.randn(n, 16, 2, 4, generator=gen)
stages = [O.Stage(w.clone(), b.clone(), gm.rh, gm.rw, gm.act) for gm, (w, b) in zip(geoms, params)]
with torch.no_grad():
    frames = torch.cat([O.decode(stages, embeds[i:i + 2]) for i in range(0, n, 2)])
frames = (frames + float(g["noise"]) * torch.randn(frames.shape, generator=gen)).contiguous()
if perturb:
    embeds = embeds * (1.0 + perturb * torch.randn(embeds.shape, generator=torch.Generator().manual_seed(5)))
qd = O.QuantDecoder(stages, g["bits"].tolist(), False)
log = []
O.model_reconstruction(qd, embeds, frames, g["order"].tolist(), int(g["iters"]), weight=0.01, b_range=(20, 2), warmup=0.2, p=2.0, lr=0.003, log=log)
with torch.no_grad():
    out = torch.cat([qd.forward(embeds[i:i + 2]) for i in range(0, n, 2)])
rec = np.array([r[2] for r in log])
print(json.dumps({"engine": "oracle (CPU)", "perturb": perturb, "psnr_calibrated": float(O.psnr(out, frames).double().mean()),
                  "loss_first6": [float(f"{v:.4e}") for v in rec[:6]], "loss_it50": float(rec[49]), "loss_last100_mean": float(rec[-100:].mean()),
                  "psnr_reference": float(g["psnr_calibrated"].mean()), "loss_it50_reference": float(g["traj"][49, 2]),
                  "loss_last100_mean_reference": float(g["traj"][-100:, 2].mean())}))
