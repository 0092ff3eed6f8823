"""GPU tests of the reference-shaped API (QuantModel / QuantModule / quantisers / model_reconstruction /
checkpoint layout) against the golden outputs of the unmodified reference."""
import io

import numpy as np
import pytest
import torch

from oracle import nq_oracle as O
from tests.helpers import CASES, LW_CASES, cw, load, t

pytestmark = pytest.mark.gpu


def build_model(tag):
    from neuroquant_b200.models import HNeRV, NeRV
    arch, cfg = CASES[tag] if tag in CASES else LW_CASES[tag]
    g = load(tag)
    model = (HNeRV if arch == "hnerv" else NeRV)(cfg)
    sd = {k[3:]: t(g[k]) for k in g.files if k.startswith("sd/")}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not [k for k in missing if "encoder" not in k], missing  # decoder / head keys are the reference's
    assert not unexpected, unexpected
    return g, arch, cfg, model.cuda()


class ListLoader(list):
    """Stand-in for the DataLoader `gt` (len + iteration of sample dicts), as tests/golden/make_golden.py uses."""


def loader_for(g):
    frames = t(g["frames"])
    return ListLoader([{"img": frames[idx], "idx": torch.as_tensor(idx), "norm_idx": torch.as_tensor(idx).float()}
                       for idx in g["order"].tolist()])


@pytest.mark.parametrize("tag", list(CASES) + list(LW_CASES))
def test_quantmodel_matches_reference(tag):
    from neuroquant_b200.quantization import QuantModel, QuantModule
    g, arch, cfg, model = build_model(tag)
    cali = t(g["cali"]).cuda()
    out, embed_list, dec_time = model.decode(cali[:2])
    assert np.abs(out.cpu().numpy() - g["fp_out"]).max() < 2e-5 and dec_time > 0
    # embed_list as the reference returns it: the input embedding, then the stem's and every block's output
    assert torch.equal(embed_list[0], cali[:2])
    n_feat = sum(1 for k in g.files if k.startswith("fp_embed"))
    assert len(embed_list) == n_feat
    for i in range(n_feat):
        want = g[f"fp_embed{i}"]
        assert tuple(embed_list[i].shape) == want.shape and np.abs(embed_list[i].cpu().numpy() - want).max() < 2e-5, i
    qnn = QuantModel(model, hadamard=bool(g["hadamard"]), weight_quant_params={"n_bits": 8, "channel_wise": cw(g),
                                                                                "scale_method": "max"}).cuda()
    assert qnn.set_bitwidth(g["bits"].tolist()) == float(g["avg_bits"])
    qnn.set_quant_state(True)
    out, _, _ = qnn(cali[:2])
    assert np.abs(out.cpu().numpy() - g["uaq_out"]).max() < 2e-5
    mods = [m for m in qnn.model.modules() if isinstance(m, QuantModule)]
    for i, m in enumerate(mods):
        assert np.array_equal(m.weight_quantizer.delta.detach().cpu().numpy(), g[f"init/{i}/delta_w"])
        assert np.array_equal(m.weight_quantizer.zero_point.cpu().numpy(), g[f"init/{i}/zp_w"])
        assert np.array_equal(m.bias_quantizer.delta.detach().cpu().numpy(), g[f"init/{i}/delta_b"])
    codes = qnn.get_quantized_param()
    assert len(codes) == 2 * len(mods) and all(torch.equal(c, c.round()) for c in codes)
    for i, v in enumerate(qnn.get_perturbation()):
        assert np.array_equal(v.cpu().numpy(), g[f"pert/{i}"])
    qnn.set_quant_state(False)
    out, _, _ = qnn(cali[:2])
    assert np.abs(out.cpu().numpy() - g["fp_out"]).max() < 2e-5


def test_standalone_module_and_quantizer_autograd():
    from neuroquant_b200.quantization.quant_layer import QuantModule
    from neuroquant_b200.quantization.quantizer import AdaRoundQuantizer, UniformAffineQuantizer, lp_loss
    torch.manual_seed(0)
    conv = torch.nn.Conv2d(5, 7, 3, 1, 1).cuda()
    x = torch.randn(2, 5, 9, 11, device="cuda")
    qm = QuantModule(conv, hadamard=False, weight_quant_params={"n_bits": 6, "channel_wise": True, "scale_method": "max"})
    assert (qm(x) - conv(x)).abs().max() < 1e-5  # quantisation off: org weights
    qm.set_quant_state(True)
    y = qm(x)
    d, z = O.uaq_init_max(conv.weight.detach().cpu(), 6, True)
    _, wq = O.uaq_quant(conv.weight.detach().cpu(), d, z, 6)
    db, zb = O.uaq_init_max(conv.bias.detach().cpu(), 6, True)
    _, bq = O.uaq_quant(conv.bias.detach().cpu(), db, zb, 6)
    want = torch.nn.functional.conv2d(x.cpu(), wq, bq, padding=1)
    assert (y.cpu() - want).abs().max() < 1e-5
    # quantiser autograd: d_delta (UAQ, straight-through) and d_alpha (AdaRound soft) match torch autograd
    w = conv.weight.detach()
    uaq = UniformAffineQuantizer(n_bits=4, channel_wise=True, scale_method="max")
    r = torch.randn_like(w)
    (uaq(w) * r).sum().backward()
    dref = d4 = None
    d4, z4 = O.uaq_init_max(w.cpu(), 4, True)
    d4 = d4.clone().requires_grad_(True)
    (O.uaq_quant(w.cpu(), d4, z4, 4)[1] * r.cpu()).sum().backward()
    assert torch.allclose(uaq.delta.grad.cpu(), d4.grad, rtol=1e-3, atol=1e-4)
    ada = AdaRoundQuantizer(uaq, weight_tensor=w, round_mode="learned_hard_sigmoid")
    ada.soft_targets = True
    (ada(w) * r).sum().backward()
    a = ada.alpha.detach().cpu().clone().requires_grad_(True)
    (O.adaround_quant(w.cpu(), a, ada.delta.detach().cpu(), ada.zero_point.cpu(), 4, True)[1] * r.cpu()).sum().backward()
    assert torch.allclose(ada.alpha.grad.cpu(), a.grad, rtol=1e-4, atol=1e-7)
    # lp_loss with gradient
    p_, t_ = torch.rand(2, 3, 8, 8, device="cuda", requires_grad=True), torch.rand(2, 3, 8, 8, device="cuda")
    lp_loss(p_, t_, 2.0).backward()
    pc = p_.detach().cpu().requires_grad_(True)
    O.lp_loss(pc, t_.cpu(), 2.0).backward()
    assert torch.allclose(p_.grad.cpu(), pc.grad, rtol=1e-5, atol=1e-8)
    with pytest.raises(AssertionError):
        UniformAffineQuantizer(n_bits=9)
    with pytest.raises(ValueError):
        QuantModule(torch.nn.Linear(3, 3))


@pytest.mark.parametrize("tag", ["tiny_hnerv", "tiny_nerv_had", "tiny_hnerv_lw"])
def test_model_reconstruction_and_checkpoint(tag, tmp_path):
    """calibrate_network.py flow on the golden tiny net: quantised forward initialises scales, then
    model_reconstruction (80 iterations in the injected batch order), then the whole-object checkpoint."""
    from neuroquant_b200.quantization import QuantModel, QuantModule, model_reconstruction
    from neuroquant_b200.quantization.quantizer import AdaRoundQuantizer
    g, arch, cfg, model = build_model(tag)
    cali, frames = t(g["cali"]).cuda(), t(g["frames"])
    qnn = QuantModel(model, hadamard=bool(g["hadamard"]), weight_quant_params={"n_bits": 8, "channel_wise": cw(g),
                                                                                "scale_method": "max"}).cuda()
    qnn.set_bitwidth(g["bits"].tolist())
    qnn.set_quant_state(True)
    qnn(cali[:2])
    model_reconstruction(qnn, cali_data=cali, gt=loader_for(g), arch=arch, batch_size=2, iters=80, weight=0.01,
                         opt_mode="mse", hadamard=bool(g["hadamard"]), b_range=(20, 2), warmup=0.2, p=2.0, lr=0.003)
    out, _, _ = qnn(cali[:2])
    assert np.abs(out.cpu().numpy() - g["calib_out"]).max() < 5e-3
    from neuroquant_b200.utils import psnr_fn_single
    assert np.abs(psnr_fn_single(out, frames[:2].cuda()).numpy() - g["calib_psnr"]).max() < 0.01
    mods = [m for m in qnn.model.modules() if isinstance(m, QuantModule)]
    for m in mods:
        assert isinstance(m.weight_quantizer, AdaRoundQuantizer) and not m.weight_quantizer.soft_targets
        assert m.bias_quantizer.soft_targets
        c = m.weight_quantizer.x_quant
        assert torch.equal(c, c.round())
        src = m.hadamard_weight if m.hadamard else m.org_weight
        want, _ = O.adaround_quant(src.cpu(), m.weight_quantizer.alpha.detach().cpu(), m.weight_quantizer.delta.detach().cpu(),
                                   m.weight_quantizer.zero_point.cpu(), m.weight_quantizer.n_bits, soft=False)
        assert torch.equal(c.cpu(), want)  # bit-exact codes for identical V and scales
    # checkpoint: whole-object pickle with the reference's state_dict keys, round trip
    path = tmp_path / "q.pth"
    torch.save(qnn, str(path))
    q2 = torch.load(str(path), weights_only=False)
    keys = set(q2.state_dict().keys())
    assert "model.decoder.0.weight_quantizer.alpha" in keys and "model.decoder.1.conv.weight_quantizer.delta" in keys
    assert "model.head_layer.bias_quantizer.alpha" in keys
    out2, _, _ = q2(cali[:2])
    assert torch.equal(out2, out)


@pytest.mark.parametrize("tag", ["tiny_hnerv", "tiny_hnerv_had", "tiny_nerv"])
def test_omega_golden(tag):
    """Omega of the reference (double-backward HVP over 4 batches of 2 frames) vs the forward-jet kernels."""
    from neuroquant_b200.quantization import QuantModel
    from neuroquant_b200.runner import DecoderRunner
    from neuroquant_b200.sensitivity import OmegaEvaluator, fisher_diag
    import copy
    g, arch, cfg, model = build_model(tag)
    cali, frames = t(g["cali"]).cuda(), t(g["frames"]).cuda()
    if arch == "hnerv":
        cali = cali / 3.0  # make_golden stored 3x the embeddings bit_assign re-encodes (see tests/test_oracle_golden.py)
    qnn = QuantModel(copy.deepcopy(model), hadamard=bool(g["hadamard"]),
                     weight_quant_params={"n_bits": 8, "channel_wise": True, "scale_method": "max"}).cuda()
    qnn.set_bitwidth(g["bits"].tolist())
    qnn.set_quant_state(True)
    qnn(t(g["cali"]).cuda()[:2])
    vec = qnn.get_perturbation()
    runner = DecoderRunner.of(model)
    runner.sync()
    ev = OmegaEvaluator(runner.engine)
    ev.set_direction(vec, 2, cali.shape[2], cali.shape[3])
    for i in range(0, 8, 2):
        ev.add_batch(cali[i:i + 2], frames[i:i + 2])
    assert ev.value() == pytest.approx(float(g["omega"]), rel=5e-3, abs=1e-12)
    # the per-layer terms the reference logs (bit_assign.py:194-200), here by polarisation of the same jet
    from neuroquant_b200.sensitivity import omega, omega_layers
    gl = load(tag + "_sens_layers")
    batches = [(cali[i:i + 2], frames[i:i + 2]) for i in range(0, 8, 2)]
    assert omega(runner.engine, vec, batches) == pytest.approx(float(g["omega"]), rel=5e-3, abs=1e-12)
    per = np.array(omega_layers(runner.engine, vec, batches))
    assert np.abs(per - gl["omega_layers"]).max() <= 5e-3 * np.abs(gl["omega_layers"]).max(), (per, gl["omega_layers"])
    assert per.sum() == pytest.approx(float(g["omega"]), rel=2e-2)  # the terms cancel: several are negative
    # fisher_diag against autograd of the oracle
    _, _, _, stages = __import__("tests.helpers", fromlist=["case_stages"]).case_stages(tag)
    ws = [s.weight.clone().requires_grad_(True) for s in stages]
    tot = [torch.zeros_like(w) for w in ws]
    for i in range(0, 8, 2):
        out = O.decode(stages, cali[i:i + 2].cpu(), ws, [s.bias for s in stages])
        gr = torch.autograd.grad(torch.nn.functional.mse_loss(out, frames[i:i + 2].cpu()), ws)
        tot = [a + b for a, b in zip(tot, gr)]
    want = sum(float((v.cpu() ** 2 * gg ** 2).sum()) for v, gg in zip(vec, tot))
    got = fisher_diag(runner.engine, vec, batches)
    assert got == pytest.approx(want, rel=2e-3)
    assert got == pytest.approx(float(g["fisher_diag"]), rel=5e-3)
    assert np.allclose(fisher_diag(runner.engine, vec, batches, per_layer=True), gl["fisher_layers"], rtol=5e-3, atol=1e-4 * gl["fisher_layers"].max())


def test_packed_weight_cache_follows_load_state_dict():
    """the runner's packed-weight cache is keyed on the owning Parameters' version counters, so
    decode -> load_state_dict(other weights) -> decode uses the NEW weights (a `.data` alias never changes version)."""
    from neuroquant_b200.quantization import QuantModel
    g, arch, cfg, model = build_model("tiny_hnerv")
    cali = t(g["cali"]).cuda()
    out0, _, _ = model.decode(cali[:2])
    other = {k: v + 0.01 * torch.randn_like(v) for k, v in model.state_dict().items() if "encoder" not in k}
    model.load_state_dict(other, strict=False)
    out1, _, _ = model.decode(cali[:2])
    assert (out1 - out0).abs().max() > 1e-4, "decode after load_state_dict still used the stale packed weights"
    g2, _, _, fresh = build_model("tiny_hnerv")
    fresh.load_state_dict(other, strict=False)
    want, _, _ = fresh.decode(cali[:2])
    assert torch.equal(out1, want)
    # the quantised model too: in-place change of a weight through torch invalidates the cache
    qnn = QuantModel(model, hadamard=False, weight_quant_params={"n_bits": 8, "channel_wise": True, "scale_method": "max"}).cuda()
    qnn.set_bitwidth([6] * 7)
    qnn.set_quant_state(True)
    a, _, _ = qnn(cali[:2])
    with torch.no_grad():
        qnn.model.head_layer.weight.mul_(1.05)
    b, _, _ = qnn(cali[:2])
    assert (a - b).abs().max() > 1e-5


def test_data_utils_surface():
    """quantization.data_utils under the reference's import path and signatures (data_utils.py:45-272)."""
    from neuroquant_b200.compat import install_reference_aliases
    install_reference_aliases()
    from quantization.data_utils import GetLayerGrad, GetLayerInpOut, quantize_model_till, save_grad_data, save_inp_oup_data
    from neuroquant_b200.quantization import QuantModel
    from tests.helpers import block_case
    g, arch, cfg, model = build_model("tiny_hnerv")
    cali = t(g["cali"]).cuda()
    qnn = QuantModel(model, hadamard=False, weight_quant_params={"n_bits": 8, "channel_wise": True, "scale_method": "max"}).cuda()
    qnn.set_bitwidth(g["bits"].tolist())
    qnn.set_quant_state(True)
    qnn(cali[:2])
    block = qnn.model.decoder[3]
    (inps, syms), outs = save_inp_oup_data(qnn, block, cali, asym=True, batch_size=4, keep_gpu=True, input_prob=True)
    assert inps.shape == syms.shape and inps.shape[0] == outs.shape[0] == 8
    (inps1,), outs1 = save_inp_oup_data(qnn, block, cali, asym=False, batch_size=4)
    assert torch.equal(inps1, syms) and torch.equal(outs1, outs)
    # the same numbers through the per-batch callable
    inp_b, out_b, sym_b = GetLayerInpOut(qnn, block, cali.device, asym=True, input_prob=True)(cali[:4])
    assert torch.equal(inp_b, inps[:4]) and torch.equal(out_b, outs[:4]) and torch.equal(sym_b, syms[:4])
    feats = qnn.model.decode(cali[:4])[1]
    qnn.set_quant_state(False)
    feats = qnn.model.decode(cali[:4])[1]
    assert (feats[3] - syms[:4]).abs().max() < 1e-6 and (feats[4] - outs[:4]).abs().max() < 1e-6
    grads = save_grad_data(qnn, block, cali, batch_size=4)
    raw = GetLayerGrad(qnn, block, cali.device)(cali[:2])
    assert grads.shape == outs.shape and torch.allclose(grads[:2], raw.abs() + 1.0)
    quantize_model_till(qnn, block)
    states = [m.use_weight_quant for m in qnn.quant_modules()]
    assert states == [True, True, True, True, False, False, False]


@pytest.mark.parametrize("tag", ["tiny_hnerv", "tiny_nerv"])
def test_omega_table_reproduces_reference_scores_of_arbitrary_configurations(tag):
    """bit_assign as a search (BASELINE configs[3]): the Gram table of Omega, measured once with forward jets, reproduces
    the UNMODIFIED reference's sensitivity_criterion (double-backward HVP, bit_assign.py:171-203) on nine configurations
    it was not built around (tests/golden/make_omega_golden.py) -- negative scores included -- and the device search over
    all 7^7 configurations returns the true minimum of the table under the average-bit budget."""
    from neuroquant_b200.methods.bit_assign import search_bit_assignment
    g, arch, cfg, model = build_model(tag)
    gc = load(tag + "_omega_configs")
    cali, frames = t(g["cali"]).cuda(), t(g["frames"]).cuda()
    # the loader's frames must map to the fixture's embeddings (bit_assign re-encodes them with the encoder, whose random
    # weights the fixture does not store)
    embeds = cali / 3.0 if arch == "hnerv" else cali
    loader = [{"img": frames[i:i + 2], "norm_idx": torch.arange(i, i + 2).float() / 8, "idx": torch.arange(i, i + 2)} for i in range(0, 8, 2)]
    table_embeds = {int(s["idx"][0]): embeds[int(s["idx"][0]):int(s["idx"][0]) + 2] for s in loader}
    model.encode = lambda x, _m=model: (table_embeds[min(table_embeds, key=lambda k: float((frames[k:k + 2] - x).abs().sum()))]
                                        if arch == "hnerv" else type(model).encode(_m, x))
    options = [2, 3, 4, 5, 6, 7, 8]
    bits, score, avg_bits, table = search_bit_assignment(arch, model, loader, cali, options, 4.5, batch_size=2)
    assert len(table.directions()) == 49 + 21 * 49 and table.jets == len(table.directions())
    want = gc["omega"]
    got = np.array([table.score(c) for c in gc["configs"].tolist()])
    scale = np.abs(want).max()
    assert np.abs(got - want).max() <= 5e-3 * scale, (got, want)
    rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-3 * scale)
    assert rel.max() < 2e-2, (got, want)
    # the search: exhaustive on the host over the same table
    bw = table.bits_weight().numpy().reshape(7, 7)
    gm = table.gram.numpy()
    idx = np.arange(7 ** 7)
    c = np.stack([(idx // 7 ** l) % 7 for l in range(7)], 1)            # c[i, l] = option index of layer l (index = sum c_l 7^l)
    a = c + 7 * np.arange(7)[None, :]
    sc = sum(gm[a[:, l], a[:, l]] for l in range(7)) + 2 * sum(gm[a[:, l], a[:, m]] for l in range(7) for m in range(l + 1, 7))
    ok = sum(bw[l, c[:, l]] for l in range(7)) <= 4.5
    sc = np.where(ok, sc, np.inf)
    best_i = int(np.argmin(sc))
    best, best_c = float(sc[best_i]), c[best_i].tolist()
    assert bits == [options[k] for k in best_c] and score == pytest.approx(best, rel=1e-12)
    assert avg_bits <= 4.5 and score == pytest.approx(table.score(bits), rel=1e-12)
    none, sc, _, _ = table.search(1.9)                     # nothing fits below 2 bits on average
    assert none is None and sc == float("inf")


def test_model_reconstruction_with_a_ragged_last_batch():
    """A loader with drop_last=False ends every epoch on a short batch (here 2 + 2 + 1 frames).  lp_loss is a mean over the
    frames actually in the batch (quantizer.py:66-73), so loss and gradients of that iteration are normalised by 1, not by
    the nominal batch size of 2; the loop alternates between the captured graphs of the two batch sizes.  Checked against
    the CPU oracle in the same batch order: per-iteration losses through `on_iteration`, then frames and PSNR."""
    from neuroquant_b200.quantization import QuantModel, model_reconstruction
    from neuroquant_b200.utils import psnr_fn_single
    tag = "tiny_hnerv"
    g, arch, cfg, model = build_model(tag)
    cali, frames = t(g["cali"]), t(g["frames"])
    batches = [[0, 5], [3, 6], [1]]
    iters = 60  # 1 epoch of step sizes (3 iterations) + 19 epochs of AdaRound (57)
    # oracle
    stages = O.stages_from_state_dict({k[3:]: t(g[k]) for k in g.files if k.startswith("sd/")}, cfg, arch)
    qd = O.QuantDecoder(stages, g["bits"].tolist(), bool(g["hadamard"]), channel_wise=cw(g))
    qd.init_scales()
    want_log = []
    O.model_reconstruction(qd, cali, frames, batches, iters, weight=0.01, b_range=(20, 2), warmup=0.2, p=2.0, lr=0.003, log=want_log)
    with torch.no_grad():
        want_out = qd.forward(cali[:2])
    # product
    qnn = QuantModel(model, hadamard=bool(g["hadamard"]), weight_quant_params={"n_bits": 8, "channel_wise": cw(g),
                                                                                "scale_method": "max"}).cuda()
    qnn.set_bitwidth(g["bits"].tolist())
    qnn.set_quant_state(True)
    qnn(cali[:2].cuda())
    loader = ListLoader([{"img": frames[idx], "idx": torch.as_tensor(idx), "norm_idx": torch.as_tensor(idx).float()}
                         for idx in batches])
    got = []
    model_reconstruction(qnn, cali_data=cali.cuda(), gt=loader, arch=arch, batch_size=2, iters=iters, weight=0.01,
                         opt_mode="mse", hadamard=bool(g["hadamard"]), b_range=(20, 2), warmup=0.2, p=2.0, lr=0.003,
                         on_iteration=lambda phase, count, loss: got.append((phase, count, loss.clone())))
    assert [(ph, c) for ph, c, _ in got] == [(r[0], r[1]) for r in want_log]
    got_rec = np.array([float(l) for _, _, l in got])
    want_rec = np.array([r[2] for r in want_log])
    # three epochs, the short batch three times; a loss normalised by the nominal batch size would be off by a factor of 2
    # (measured on B200: every one of the 60 losses within 4.2e-7 relative, frames within 1.6e-6)
    assert np.allclose(got_rec[:9], want_rec[:9], rtol=1e-4, atol=1e-7)
    assert np.allclose(got_rec, want_rec, rtol=5e-3, atol=1e-6)  # the bars of the 80-iteration golden run
    out, _, _ = qnn(cali[:2].cuda())
    assert (out.cpu() - want_out).abs().max() < 5e-3
    assert (psnr_fn_single(out, frames[:2].cuda()).cpu() - O.psnr(want_out, frames[:2])).abs().max() < 0.01
