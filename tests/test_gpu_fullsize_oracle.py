"""Full-size parity on the GPU against the CPU ORACLE (oracle/nq_oracle.py, pinned to the unmodified reference by
tests/golden): the tensor-core engine at the sizes the benchmark is quoted on.

BASELINE.json configs[0] is the reference's own CPU-runnable case (HNeRV-Bunny-3M, channel-wise W6); the metric's
configuration is W-mixed 6 5 4 5 5 6 6; configs[1] is NeRV-Bunny-3M and configs[2] adds --hadamard.  For each the CPU
oracle runs forward + lp_loss + autograd in fp32 on the box's host cores (about 0.6 s per pass at 2 x 640 x 1280) and
the tensor-core engine must reproduce, for IDENTICAL V (alpha) and scales:

  * step sizes / zero points bit-exact, hard AdaRound codes bit-exact,
  * soft-rounded and hard-rounded frames within 1e-3 max-abs (north star), loss within 1e-5 relative,
  * every d_alpha (phase 2) within 2e-4 and every d_delta (phase 1) within DELTA_TOL of that tensor's largest entry,
  * after ITERS AdaRound iterations in the same mini-batch order: PSNR of the hard-rounded decode within 0.01 dB.

Reference lines: quantization/calib_model.py:206-226 (the iteration), quant_layer.py:67-81 (the quantised forward),
quantizer.py:111-125,278-300 (the two quantisers)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

MIXED = [6, 5, 4, 5, 5, 6, 6]
# (workload, bits, hadamard)
RUNS = [("hnerv-bunny-3m", [6] * 7, False),   # BASELINE configs[0]
        ("hnerv-bunny-3m", MIXED, False),     # the metric's configuration
        ("nerv-bunny-3m", MIXED, False),      # configs[1]
        ("hnerv-bunny-3m", MIXED, True)]      # configs[2] (--hadamard)
# d_delta is a difference of two large sums (x_q - zp and x / delta summed over a channel): the reference's own value is
# only defined to about 1e-3 of its scale in fp32, so the bar is looser than for d_alpha
DELTA_TOL = 2e-3
ALPHA_TOL = 2e-4


def _build(workload, bits, hadamard):
    import neuroquant_b200 as nq
    from neuroquant_b200.workloads import WORKLOADS, embed_shape, random_decoder
    from oracle import nq_oracle as O
    arch, cfg = WORKLOADS[workload]
    geoms, params = random_decoder(cfg, arch, 903)
    stages = [nq.QuantStage(g, w.cuda(), b.cuda(), nb, hadamard) for g, (w, b), nb in zip(geoms, params, bits)]
    eng = nq.DecoderEngine(stages)
    assert eng.use_tc, "this file checks the tensor-core engine"
    eng.init_scales()
    qd = O.QuantDecoder([O.Stage(w.clone(), b.clone(), g.rh, g.rw, g.act) for g, (w, b) in zip(geoms, params)], bits, hadamard)
    c, h0, w0 = embed_shape(cfg, arch)
    return eng, qd, cfg, (c, h0, w0)


def _adopt_alpha(eng, qd):
    """Identical V on both sides: the engine takes the oracle's alpha (GPU logf differs from the CPU's in the last ulp)."""
    for s, q in zip(eng.stages, qd.q):
        assert float((s.alpha_w.cpu() - q.alpha_w).abs().max()) < 1e-4
        s.alpha_w.copy_(q.alpha_w)
        s.alpha_b.copy_(q.alpha_b)
    eng.invalidate()


@pytest.mark.parametrize("workload,bits,hadamard", RUNS)
def test_tensor_core_engine_matches_cpu_oracle_at_full_size(workload, bits, hadamard):
    from oracle import nq_oracle as O
    torch.set_num_threads(max(1, (torch.get_num_threads() or 1)))
    eng, qd, cfg, (c, h0, w0) = _build(workload, bits, hadamard)
    gen = torch.Generator().manual_seed(17)
    batch = 2
    embed = torch.randn(batch, c, h0, w0, generator=gen)
    frames = torch.rand(batch, 3, cfg["crop_h"], cfg["crop_w"], generator=gen)
    embed_d, frames_d = embed.cuda(), frames.cuda()
    report = {}

    # ---- scales: bit-exact (quantizer.py:127-168)
    for s, q in zip(eng.stages, qd.q):
        assert torch.equal(s.delta_w.cpu().view(-1), q.delta_w.view(-1)) and torch.equal(s.zp_w.cpu().view(-1), q.zp_w.view(-1))
        assert torch.equal(s.delta_b.cpu().view(-1), q.delta_b.view(-1)) and torch.equal(s.zp_b.cpu().view(-1), q.zp_b.view(-1))

    # ---- phase 1 (UAQ, straight-through rounding): frames, loss, d_delta (calib_model.py:145-165)
    for q in qd.q:
        q.delta_w = q.delta_w.clone().requires_grad_(True)
        q.delta_b = q.delta_b.clone().requires_grad_(True)
    out = qd.forward(embed)
    loss = O.lp_loss(out, frames, 2.0)
    loss.backward()
    img = eng.forward(embed_d, train=True, target=frames_d, p_norm=2.0).cpu()
    report["uaq_frames"] = float((img - out.detach()).abs().max())
    assert report["uaq_frames"] < 1e-3, report
    assert float(eng.last_loss()) == pytest.approx(float(loss), rel=1e-5)
    eng.backward()
    worst = 0.0
    for i, ((dw, db), q) in enumerate(zip(eng.param_grads(), qd.q)):
        for name, a, b in (("d_delta_w", dw.cpu().view(-1), q.delta_w.grad.view(-1)), ("d_delta_b", db.cpu().view(-1), q.delta_b.grad.view(-1))):
            rel = float((a - b).abs().max()) / (float(b.abs().max()) + 1e-20)
            worst = max(worst, rel)
            assert rel <= DELTA_TOL, (workload, i, name, rel)
    report["d_delta_rel"] = worst
    for q in qd.q:
        q.delta_w, q.delta_b = q.delta_w.detach(), q.delta_b.detach()

    # ---- phase 2 (AdaRound, soft targets): frames, loss, d_alpha (calib_model.py:206-226)
    qd.start_adaround()
    eng.start_adaround()
    for s, q in zip(eng.stages, qd.q):  # fp16-rounded scales (quantizer.py:264-265): bit-exact
        assert torch.equal(s.delta_w.cpu().view(-1), q.delta_w.view(-1)) and torch.equal(s.zp_w.cpu().view(-1), q.zp_w.view(-1))
    _adopt_alpha(eng, qd)
    for q in qd.q:
        q.alpha_w.requires_grad_(True)
        q.alpha_b.requires_grad_(True)
    out = qd.forward(embed)
    loss = O.lp_loss(out, frames, 2.0)
    loss.backward()
    img = eng.forward(embed_d, train=True, target=frames_d, p_norm=2.0).cpu()
    report["soft_frames"] = float((img - out.detach()).abs().max())
    assert report["soft_frames"] < 1e-3, report
    assert float(eng.last_loss()) == pytest.approx(float(loss), rel=1e-5)
    eng.backward()
    worst = 0.0
    for i, ((da_w, da_b), q) in enumerate(zip(eng.param_grads(), qd.q)):
        for name, a, b in (("d_alpha_w", da_w.cpu(), q.alpha_w.grad), ("d_alpha_b", da_b.cpu(), q.alpha_b.grad)):
            rel = float((a - b).abs().max()) / (float(b.abs().max()) + 1e-20)
            worst = max(worst, rel)
            assert rel <= ALPHA_TOL, (workload, i, name, rel)
    report["d_alpha_rel"] = worst

    # ---- the deliverable: hard-rounded codes bit-exact, hard-rounded decode within 1e-3
    for q in qd.q:
        q.alpha_w = q.alpha_w.detach()
        q.alpha_b = q.alpha_b.detach()
    qd.soft_w = False
    eng.soft_w = False
    eng.invalidate()
    with torch.no_grad():
        out = qd.forward(embed)
    img = eng.forward(embed_d).cpu()
    for s, q in zip(eng.stages, qd.q):
        assert torch.equal(s.codes_w.cpu(), q.codes_w), "hard codes differ from the reference quantiser's"
        assert torch.equal(s.codes_w, s.codes_w.round())
    report["hard_frames"] = float((img - out).abs().max())
    report["hard_psnr_diff_db"] = float((O.psnr(img, frames) - O.psnr(out, frames)).abs().max())
    print(workload, bits, hadamard, report)
    assert report["hard_frames"] < 1e-3, report
    assert report["hard_psnr_diff_db"] < 0.01, report


ITERS = 24


@pytest.mark.parametrize("workload,bits,hadamard", [RUNS[1], RUNS[2], RUNS[3]])
def test_adaround_iterations_match_cpu_oracle_at_full_size(workload, bits, hadamard):
    """ITERS AdaRound iterations (calib_model.py:206-226, regulariser on from count 5) on 8 frames, 4 mini-batches of 2 in
    a fixed order, on the CPU oracle and on the tensor-core engine through CalibrationLoop:
    loss trajectories agree, the hard-rounded decodes end within 0.01 dB PSNR of each other (mean; 0.03 dB per frame)."""
    import neuroquant_b200 as nq
    from oracle import nq_oracle as O
    eng, qd, cfg, (c, h0, w0) = _build(workload, bits, hadamard)
    gen = torch.Generator().manual_seed(29)
    embeds = torch.randn(8, c, h0, w0, generator=gen)
    # targets = the full-precision decoder's frames + noise of the quantisation error's own power (2e-4: nearest rounding
    # costs ~74 dB on a random-init decoder), the reference's situation -- fit error and rounding error comparable -- and
    # the one in which a PSNR bar discriminates (tests/golden/make_fullsize_golden.py)
    with torch.no_grad():
        frames = O.decode(qd.stages, embeds)
    frames = (frames + 2.0e-4 * torch.randn(frames.shape, generator=gen)).contiguous()
    order = [[0, 5], [3, 6], [1, 4], [7, 2]]
    hyper = dict(weight=0.01, b_range=(20, 2), warmup=0.2, p=2.0, lr=0.003)
    log_o = []
    O.model_reconstruction(qd, embeds, frames, order, ITERS, log=log_o, **hyper)
    with torch.no_grad():
        out_o = torch.cat([qd.forward(embeds[i:i + 2]) for i in range(0, 8, 2)])
    psnr_o = O.psnr(out_o, frames)

    embeds_d, frames_d = embeds.cuda(), frames.cuda()

    def fetch(idx):
        idx = torch.as_tensor(idx, device="cuda")
        return embeds_d[idx], frames_d[idx]

    log_g = []
    loop = nq.CalibrationLoop(eng, fetch, len(order), iters=ITERS, log=log_g, **hyper)
    loop.run(lambda: order)
    out_g = torch.cat([eng.forward(embeds_d[i:i + 2]).clone() for i in range(0, 8, 2)]).cpu()
    psnr_g = O.psnr(out_g, frames)
    assert len(log_g) == len(log_o) == ITERS
    rec_o = torch.tensor([r[2] for r in log_o])
    rec_g = torch.tensor([r[2] for r in log_g])
    rnd_o = torch.tensor([r[3] for r in log_o])
    rnd_g = torch.tensor([r[3] for r in log_g])
    report = dict(first_rel=float(((rec_g[:4] - rec_o[:4]).abs() / rec_o[:4]).max()), all_rel=float(((rec_g - rec_o).abs() / rec_o).max()),
                  round_rel=float(((rnd_g - rnd_o).abs() / rnd_o.clamp_min(1e-12)).max()),
                  psnr_oracle=float(psnr_o.mean()), psnr_gpu=float(psnr_g.mean()), psnr_frame=float((psnr_g - psnr_o).abs().max()))
    print(workload, bits, hadamard, report)
    assert report["first_rel"] < 1e-4, report     # the first passes over each batch: same function, same inputs
    assert report["all_rel"] < 2e-2, report       # later: Adam's sign-like first steps amplify rounding-level gradients
    assert report["round_rel"] < 1e-3, report
    assert abs(report["psnr_gpu"] - report["psnr_oracle"]) < 0.01, report
    assert report["psnr_frame"] < 0.03, report
    # final integer codes: the reference quantiser's for the engine's own V and scales
    for s in eng.stages:
        want, _ = O.adaround_quant(s.w_src.cpu(), s.alpha_w.cpu(), s.delta_w.cpu(), s.zp_w.cpu(), s.n_bits, False)
        assert torch.equal(s.codes_w.cpu(), want)


# ---------------------------------------------------------------------------------------------------------------------
# Against the UNMODIFIED reference itself, run at full size on CPU by tests/golden/make_fullsize_golden.py
# ---------------------------------------------------------------------------------------------------------------------
class ListLoader(list):
    """Stand-in for the DataLoader `gt` (len + iteration of sample dicts), as the fixture generator uses."""


def _fixture_inputs(g):
    """Weights, embeddings, frames of the fixture, regenerated from its seeds (and checked against its checksums)."""
    from neuroquant_b200.workloads import WORKLOADS, random_decoder
    from oracle import nq_oracle as O
    arch, cfg = WORKLOADS["hnerv-bunny-3m"]
    geoms, params = random_decoder(cfg, arch, 903)
    for i, (w, _) in enumerate(params):
        assert float(w.double().sum()) == pytest.approx(float(g["w_sum"][i]), rel=1e-12, abs=1e-12)
        assert float(w.double().abs().sum()) == pytest.approx(float(g["w_abs"][i]), rel=1e-12)
    gen = torch.Generator().manual_seed(29)
    n = int(g["order"].size)
    embeds = torch.randn(n, 16, 2, 4, generator=gen)
    assert float(embeds.double().sum()) == pytest.approx(float(g["embeds_sum"]), rel=1e-12)
    stages = [O.Stage(w, b, gm.rh, gm.rw, gm.act) for gm, (w, b) in zip(geoms, params)]
    with torch.no_grad():
        frames = torch.cat([O.decode(stages, embeds[i:i + 2]) for i in range(0, n, 2)])
    frames = (frames + float(g["noise"]) * torch.randn(frames.shape, generator=gen)).contiguous()
    assert float(frames.double().mean()) == pytest.approx(float(g["frames_mean"]), rel=1e-6)
    return cfg, params, embeds, frames


@pytest.mark.parametrize("name", ["config1", "mixed1000"])
def test_model_reconstruction_matches_reference_run_at_full_size(name):
    """BASELINE.json configs[0] (HNeRV-Bunny-3M, channel-wise W6, 100 AdaRound iterations, batch 2) and a 1000-iteration
    two-phase run at the metric's W 6 5 4 5 5 6 6, through the product API exactly as calibrate_network.py drives it
    (QuantModel -> set_bitwidth -> quantised forward -> model_reconstruction), against what the UNMODIFIED reference
    produced on CPU from the same seeded inputs in the same mini-batch order (tests/golden/fullsize_*.npz): average
    bit-width, nearest-rounding PSNR, the loss trajectory, and the PSNR of the hard-rounded decode after calibration --
    within the north star's 0.01 dB for configs[0] (73.30165 against 73.30166 dB); for the two-phase run, whose step-size
    phase is chaotic in the reference's own arithmetic, within the spread measured between equivalent evaluations (see the
    comment at the assertions).  Final codes: bit-exact with the reference quantiser for the run's own V."""
    import os
    import numpy as np
    from neuroquant_b200.models import HNeRV
    from neuroquant_b200.quantization import QuantModel, QuantModule, model_reconstruction
    from neuroquant_b200.utils import psnr_fn_single
    from oracle import nq_oracle as O
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"fullsize_{name}.npz")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated")
    g = np.load(path)
    cfg, params, embeds, frames = _fixture_inputs(g)
    bits, iters = g["bits"].tolist(), int(g["iters"])
    torch.manual_seed(1)
    model = HNeRV(dict(cfg))
    convs = [model.decoder[0]] + [blk.conv[0] for blk in list(model.decoder)[1:]] + [model.head_layer]
    with torch.no_grad():
        for c, (w, b) in zip(convs, params):
            c.weight.copy_(w)
            c.bias.copy_(b)
    model = model.cuda()
    embeds_d, frames_d = embeds.cuda(), frames.cuda()

    def psnr_all(net):
        with torch.no_grad():
            return torch.cat([psnr_fn_single((net.decode(embeds_d[i:i + 2]) if isinstance(net, HNeRV) else net(embeds_d[i:i + 2]))[0],
                                             frames_d[i:i + 2]) for i in range(0, embeds.shape[0], 2)]).double().numpy()

    report = {"psnr_fp": float(np.abs(psnr_all(model) - g["psnr_fp"]).max())}
    qnn = QuantModel(model, hadamard=False, weight_quant_params={"n_bits": 8, "channel_wise": True, "scale_method": "max"}).cuda()
    assert qnn.set_bitwidth(bits) == float(g["avg_bits"])
    qnn.eval()
    qnn.set_quant_state(True)
    qnn(embeds_d[:2])
    mods = [m for m in qnn.model.modules() if isinstance(m, QuantModule)]
    for i, m in enumerate(mods):  # 'max' step sizes: bit-exact with the reference
        assert np.array_equal(m.weight_quantizer.delta.detach().cpu().numpy(), g[f"init/{i}/delta_w"])
    report["psnr_nearest"] = float(np.abs(psnr_all(qnn) - g["psnr_nearest"]).max())
    loader = ListLoader([{"img": frames[torch.tensor(ix)], "norm_idx": torch.tensor(ix).float() / 20, "idx": torch.tensor(ix)}
                         for ix in g["order"].tolist()])
    losses = []
    model_reconstruction(qnn, cali_data=embeds_d, gt=loader, arch="hnerv", batch_size=2, iters=iters, weight=0.01, opt_mode="mse",
                         hadamard=False, b_range=(20, 2), warmup=0.2, p=2.0, lr=0.003,
                         on_iteration=lambda phase, count, loss: losses.append(loss.clone()))
    rec = torch.stack(losses).view(-1).cpu().double().numpy()
    traj = g["traj"]
    assert len(rec) == len(traj)
    rel = np.abs(rec - traj[:, 2]) / traj[:, 2]
    report.update(traj_first=float(rel[:10].max()), traj_all=float(rel.max()), traj_last100=float(rel[-100:].mean()))
    report.update(first6_gpu=[float(f'{v:.4e}') for v in rec[:6]], first6_ref=[float(f'{v:.4e}') for v in traj[:6, 2]])
    got = psnr_all(qnn)
    want = g["psnr_calibrated"]
    report.update(psnr_gpu=float(got.mean()), psnr_ref=float(want.mean()), psnr_nearest_ref=float(g["psnr_nearest"].mean()),
                  psnr_fp_ref=float(g["psnr_fp"].mean()), psnr_frame=float(np.abs(got - want).max()))
    print(name, report)
    assert report["psnr_fp"] < 1e-3 and report["psnr_nearest"] < 2e-3, report
    if name == "config1":  # AdaRound phase only (int(0.05 * 100 / 10) = 0 step-size epochs): the run tracks the reference's
        assert report["traj_first"] < 1e-4, report
        assert report["traj_all"] < 5e-2 and report["traj_last100"] < 1e-2, report
        assert abs(report["psnr_gpu"] - report["psnr_ref"]) < 0.01, report     # dB: the north-star bar
        assert report["psnr_frame"] < 0.03, report
    else:
        # 50 step-size iterations first.  Their gradient is a difference of two large sums and Adam's first steps are
        # sign-like: the loss goes 2.29e-07 -> 4.775e-04 -> 3.23e-05 in three iterations, identically here and in the
        # reference, and from the fourth iteration on ANY two evaluations separate -- the reference's own arithmetic (the
        # CPU oracle) with its input perturbed by 1e-7, the exact-fp32 engine, the tensor-core engine land 0.07-0.3 dB
        # apart after 1000 iterations (tools/chaos_mixed1000.py, tools/oracle_mixed1000.py, profiles/r02z_chaos_*.json).
        # What is well defined is asserted: the first three iterations, the level the loss settles at, and that the
        # calibration wins back what the reference's does (71.10 dB nearest -> 72.47 dB calibrated).
        assert float(rel[:3].max()) < 1e-3, report
        assert report["traj_last100"] < 0.1, report
        gain_ref = report["psnr_ref"] - report["psnr_nearest_ref"]
        assert abs(report["psnr_gpu"] - report["psnr_ref"]) < 0.3, report
        assert report["psnr_gpu"] - report["psnr_nearest_ref"] > 0.8 * gain_ref, report
    for m in mods:  # final integer codes: the reference quantiser's for this run's own V and (fp16-rounded) scales
        wq = m.weight_quantizer
        c = wq.x_quant
        want_c, _ = O.adaround_quant(m.org_weight.cpu(), wq.alpha.detach().cpu(), wq.delta.detach().cpu(), wq.zero_point.cpu(), wq.n_bits, False)
        assert torch.equal(c.cpu(), want_c)
