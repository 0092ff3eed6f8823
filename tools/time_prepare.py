"""Time DecoderEngine.prepare_weights (fake-quant of every tensor + every operand pack) and one whole eager iteration with
CUDA events; NQ_LIB_PATH selects the library build (A/B on one box)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import neuroquant_b200 as nq
from neuroquant_b200.workloads import WORKLOADS, embed_shape, random_decoder

arch, cfg = WORKLOADS["hnerv-bunny-3m"]
geoms, params = random_decoder(cfg, arch, 903)
bits = [6, 5, 4, 5, 5, 6, 6]
eng = nq.DecoderEngine([nq.QuantStage(g, w.cuda(), b.cuda(), nb, False) for g, (w, b), nb in zip(geoms, params, bits)])
eng.init_scales(); eng.start_adaround()
c, h0, w0 = embed_shape(cfg, arch)
embed = torch.randn(2, c, h0, w0).cuda(); frames = torch.rand(2, 3, cfg["crop_h"], cfg["crop_w"]).cuda()
eng.forward(embed, train=True, target=frames); eng.backward()
p = eng._last_plan
def timed(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
print(json.dumps({"lib": os.environ.get("NQ_LIB_PATH", "in-tree"), "prepare_weights_us": timed(lambda: eng.prepare_weights(p, need_wt=True, reg_b=10.0)),
                  "iteration_eager_us": timed(lambda: (eng.forward(embed, train=True, target=frames, reg_b=10.0), eng.backward()), 20)}))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(20):
        eng.prepare_weights(p, need_wt=True, reg_b=10.0)
    torch.cuda.synchronize()
for ev in prof.key_averages():
    if ev.device_time_total > 0:
        print(f"  {ev.key[:60]:60s} n={ev.count:4d} avg {ev.device_time_total / ev.count:8.1f} us")
