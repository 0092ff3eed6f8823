"""GPU parity of block-wise reconstruction (neuroquant_b200.quantization.block_reconstruction)
against the unmodified reference's calib_block.block_reconstruction: fixtures from tests/golden/make_block_golden.py, the
reference's unseeded torch.randperm / torch.rand_like draws replayed."""
import numpy as np
import pytest
import torch

from oracle import nq_oracle as O
from tests.helpers import BLOCK_CASES, LAYER_CASES, load, t

pytestmark = pytest.mark.gpu


def build(tag):
    from neuroquant_b200.models import HNeRV, NeRV
    from neuroquant_b200.quantization import QuantModel
    arch, cfg = BLOCK_CASES[tag] if tag in BLOCK_CASES else LAYER_CASES[tag]
    g = load(tag)
    model = (HNeRV if arch == "hnerv" else NeRV)(cfg)
    sd = {k[3:]: t(g[k]) for k in g.files if k.startswith("sd/")}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not [k for k in missing if "encoder" not in k] and not unexpected
    qnn = QuantModel(model.cuda(), hadamard=bool(g["hadamard"]),
                     weight_quant_params={"n_bits": 8, "channel_wise": True, "scale_method": "max"}).cuda()
    qnn.set_bitwidth(g["bits"].tolist())
    qnn.eval()
    qnn.set_quant_state(True)
    qnn(t(g["cali"])[:2].cuda())
    return g, qnn


@pytest.mark.parametrize("tag", list(BLOCK_CASES) + list(LAYER_CASES))
def test_block_reconstruction_matches_reference(tag, monkeypatch):
    """layer_* cases: layer_reconstruction against the reference's own function with its missing `opt_params = []`
    inserted in memory (tests/golden/make_block_golden.py)."""
    from neuroquant_b200.quantization import QuantModule, block_reconstruction, layer_reconstruction
    import neuroquant_b200.quantization.calib_block as cb
    g, qnn = build(tag)
    which = str(g["layer"]) if "layer" in g.files else ""
    block = {"": lambda: qnn.model.decoder[int(g["block_idx"])], "conv": lambda: qnn.model.decoder[int(g["block_idx"])].conv,
             "head": lambda: qnn.model.head_layer, "stem": lambda: qnn.model.decoder[0]}[which]()
    reconstruct = layer_reconstruction if which else block_reconstruction
    conv = [m for m in block.modules() if isinstance(m, QuantModule)][0]
    idx_seq = [torch.as_tensor(r) for r in g["idx"]]
    masks = [t(m) for m in g["masks"]] if "masks" in g.files else []
    calls = {"perm": 0, "rand": 0}
    n_cached = g["cache_inp"].shape[0]

    def randperm(n, *a, **k):
        assert n == n_cached
        r = idx_seq[calls["perm"]]
        calls["perm"] += 1
        rest = torch.tensor([v for v in range(n) if v not in r.tolist()])
        return torch.cat([r, rest])

    def rand_like(x, *a, **k):
        # the product draws its QDrop mask in the engine's NHWC layout (n, h, w, cin_p); the recorded draw is NCHW
        r = masks[calls["rand"]].to(x.device).permute(0, 2, 3, 1)
        calls["rand"] += 1
        full = torch.ones(x.shape, device=x.device)
        full[..., :r.shape[-1]] = r
        return full

    caches = {}
    _save = cb.save_inp_oup_data
    _grads = cb.block_output_grads

    def grads(*a, **k):
        caches["raw_grad"] = _grads(*a, **k)
        return caches["raw_grad"]

    def save(*a, **k):
        r = _save(*a, **k)
        caches["inp"], caches["sym"], caches["out"] = r[0][0], r[0][1], r[1]
        return r

    traj = []
    _run = cb.BlockStep.run_cached

    def run(self, *a, **k):
        _run(self, *a, **k)
        traj.append(self.rec_loss())

    monkeypatch.setattr(torch, "randperm", randperm)
    monkeypatch.setattr(torch, "rand_like", rand_like)
    monkeypatch.setattr(cb, "save_inp_oup_data", save)
    monkeypatch.setattr(cb, "block_output_grads", grads)
    monkeypatch.setattr(cb.BlockStep, "run_cached", run)
    reconstruct(qnn, block, t(g["cali"]).cuda(), batch_size=int(g["bsz"]), iters=int(g["iters"]), weight=0.01,
                         opt_mode=str(g["opt_mode"]) if "opt_mode" in g.files else "mse", asym=bool(g["asym"]), b_range=(20, 2), warmup=0.2, input_prob=float(g["input_prob"]),
                         p=2.0, lr=0.003)
    monkeypatch.undo()
    assert calls["perm"] == int(g["iters"]) and calls["rand"] == len(masks)
    # cached block inputs / outputs: full-precision and (asym) quantised-predecessor activations
    for name in ("inp", "sym", "out"):
        ref = g["cache_" + name]
        assert np.abs(caches[name].cpu().numpy() - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max()), name
    if "raw_grad" in g.files:
        # Fisher modes: the cached output gradients (GetLayerGrad), raw -- |g| + 1 is 1.0 in fp32 at these magnitudes
        ref = g["raw_grad"]
        got = caches["raw_grad"].cpu().numpy()
        assert got.shape == ref.shape
        # g is the back-propagated DIFFERENCE of two nearly equal frames (out_q - out_fp ~ 1e-4 at 5-6 bits): fp32
        # cancellation leaves about three digits in either implementation
        assert np.abs(got - ref).max() <= 1e-2 * np.abs(ref).max(), (np.abs(got - ref).max(), np.abs(ref).max())
        # the cache itself: |g| + 1 in fp32 -- 1.0 or 1.0000001 (the head's gradients straddle half an ulp of 1)
        cache = np.abs(got) + np.float32(1.0)
        assert np.abs(cache - g["cache_grad"]).max() <= 1.2e-7 and (cache != g["cache_grad"]).mean() < 0.01
    # loss trajectory of the block output
    assert np.allclose(np.array(traj), g["traj"][:, 2], rtol=5e-3, atol=1e-9)
    wq, bq = conv.weight_quantizer, conv.bias_quantizer
    assert wq.soft_targets is False and bq.soft_targets is False
    assert np.array_equal(wq.delta.detach().cpu().numpy(), g["final/delta_w"]) and np.array_equal(wq.zero_point.cpu().numpy(), g["final/zp_w"])
    assert np.array_equal(bq.delta.detach().cpu().numpy(), g["final/delta_b"]) and np.array_equal(bq.zero_point.cpu().numpy(), g["final/zp_b"])
    far = (np.abs(wq.alpha.detach().cpu().numpy() - g["final/alpha_w"]) > 2e-3).mean()
    assert far < 0.02, far
    # identical alpha / scales -> bit-exact hard codes: re-derive on the oracle from OUR alpha
    want, _ = O.adaround_quant(conv.org_weight.cpu(), wq.alpha.detach().cpu(), wq.delta.detach().cpu(), wq.zero_point.cpu(),
                               wq.n_bits, soft=False)
    with torch.no_grad():
        y = block(caches["inp"][:2])
    assert torch.equal(wq.x_quant.cpu(), want)
    assert (wq.x_quant.cpu().numpy() != g["final/codes_w"]).mean() < 1e-2
    ref = g["final/block_out"]
    assert np.abs(y.cpu().numpy() - ref).max() <= 0.02 * np.abs(ref).max()
    # the whole decoder still runs with this block AdaRound-hard and every other layer on plain rounding
    qnn.eval()
    qnn.set_quant_state(True)
    out, _, _ = qnn(t(g["cali"])[:2].cuda())
    assert torch.isfinite(out).all()


@pytest.mark.parametrize("mode", ["fisher_diag", "fisher_full"])
@pytest.mark.parametrize("conv", ["tc"])
def test_fisher_block_step_against_autograd(mode, conv):
    """One block iteration with a NON-trivial Fisher cache (the reference's own is 1.0 everywhere): loss and weight /
    bias gradients of calib_block.py:66-72 against PyTorch autograd of the oracle's soft fake-quant."""
    import torch.nn.functional as F
    import neuroquant_b200 as nq
    from neuroquant_b200.quantization.calib_block import BlockStep, nhwc_cache
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    gen = torch.Generator().manual_seed(5)
    g = nq.StageGeom(14, 12 * 4, 3, 2, 2, "gelu")
    wt = (torch.randn(g.cout, g.cin, 3, 3, generator=gen) * 0.1).cuda()
    bs = (torch.randn(g.cout, generator=gen) * 0.05).cuda()
    st = nq.QuantStage(g, wt, bs, 5, False)
    st.init_scales()
    st.start_adaround()
    n, h, w = 3, 9, 20
    x = torch.randn(n, g.cin, h, w, generator=gen).cuda()
    n_cache = 5
    tgt = (torch.randn(n_cache, g.c_grp, h * 2, w * 2, generator=gen) * 0.3).cuda()
    fis = (1.0 + torch.rand(n_cache, g.c_grp, h * 2, w * 2, generator=gen)).cuda()
    idx = torch.tensor([4, 0, 2], dtype=torch.int32).cuda()
    _, wq = O.adaround_quant(wt, st.alpha_w, st.delta_w, st.zp_w, st.n_bits, True)
    _, bq = O.adaround_quant(bs, st.alpha_b, st.delta_b, st.zp_b, st.n_bits, True)
    wq, bq = wq.detach().requires_grad_(True), bq.detach().requires_grad_(True)
    y = F.gelu(F.pixel_shuffle(F.conv2d(x, wq, bq, padding=1), 2))
    cur_t, cur_f = tgt[idx.long()], fis[idx.long()]
    if mode == "fisher_diag":
        loss = ((y - cur_t).pow(2) * cur_f.pow(2)).sum(1).mean()
    else:
        a = (y - cur_t).abs()
        loss = (torch.sum(a * cur_f, (1, 2, 3)).view(-1, 1, 1, 1) * a * cur_f).mean() / 100
    loss.backward()
    step = BlockStep(st, n, h, w, lr=1e-3)
    step.run(x, tgt[:n], 0.0, 0.0, 2.0)  # fills step.x with the split-bf16 input (and runs one plain iteration)
    st.start_adaround()
    step2 = BlockStep(st, n, h, w, lr=1e-3)
    step2.run_cached(step.x, nhwc_cache(step2, tgt), idx, 0.0, 0.0, 2.0, opt_mode=mode, fisher_cache=nhwc_cache(step2, fis))
    assert step2.rec_loss() == pytest.approx(float(loss), rel=3e-5)
    for name, a_, b_ in (("dW", step2.gw, wq.grad), ("db", step2.gb, bq.grad)):
        tol = 2e-4 * float(b_.abs().max()) + 1e-12
        assert float((a_ - b_).abs().max()) <= tol, (name, float((a_ - b_).abs().max()), tol)


@pytest.mark.parametrize("conv", ["tc", "simt"])
def test_output_gradient_cache_every_block_and_layer(conv, monkeypatch):
    """GetLayerGrad (data_utils.py:222-258) for EVERY block (the last one's consumer is the head) and every lone layer
    (stem and head included) against the oracle's block_grad_cache, which is pinned to the reference's raw hook
    gradients on the fixtures."""
    monkeypatch.setenv("NQ_CONV", conv)
    import neuroquant_b200.quantization.calib_block as cb
    from neuroquant_b200.quantization import QuantModule
    from neuroquant_b200.quantization.quantizer import AdaRoundQuantizer
    from neuroquant_b200.runner import DecoderRunner
    from tests.helpers import block_case
    tag = "block_tiny_hnerv"
    g, arch, cfg, stages = block_case(tag)
    cali = t(g["cali"])[:3]
    n_stage = len(stages)
    for k, layer in [(k, False) for k in range(1, n_stage - 1)] + [(k, True) for k in (0, 2, n_stage - 1)]:
        _, qnn = build(tag)
        runner = DecoderRunner.of(qnn.model)
        convs = [m for m in qnn.model.modules() if isinstance(m, QuantModule)]
        conv_k = convs[k]
        block = conv_k if layer else qnn.model.decoder[k]
        qnn.set_quant_state(False)
        block.set_quant_state(True)
        conv_k.weight_quantizer = AdaRoundQuantizer(uaq=conv_k.weight_quantizer, round_mode="learned_hard_sigmoid",
                                                    weight_tensor=conv_k.org_weight.data)
        conv_k.bias_quantizer = AdaRoundQuantizer(uaq=conv_k.bias_quantizer, round_mode="learned_hard_sigmoid",
                                                  weight_tensor=conv_k.bias.data)
        conv_k.weight_quantizer.soft_targets = conv_k.bias_quantizer.soft_targets = True
        runner._key = None
        got = cb.block_output_grads(qnn, runner, block, k, cali.cuda(), layer).cpu()
        qd = O.QuantDecoder(stages, g["bits"].tolist(), False)
        q = qd.q[k]
        q.delta_w, q.zp_w = O.fp16_round(q.delta_w), O.fp16_round(q.zp_w)
        q.delta_b, q.zp_b = O.fp16_round(q.delta_b), O.fp16_round(q.zp_b)
        q.alpha_w, q.alpha_b = O.adaround_init_alpha(q.stage.weight, q.delta_w), O.adaround_init_alpha(q.stage.bias, q.delta_b)
        want = O.block_grad_cache(qd, k, cali, raw=True, layer=layer)
        assert got.shape == want.shape, (k, layer)
        assert float((got - want).abs().max()) <= 1e-2 * float(want.abs().max()), (k, layer, float((got - want).abs().max()), float(want.abs().max()))


def test_batch_assembly_kernel_equals_torch_where():
    """nq_qdrop_gather (calib_block.py:160-164) against advanced indexing + torch.where on the split-bf16 planes, with and
    without QDrop, and its argument checks."""
    import neuroquant_b200._lib as L
    from neuroquant_b200.quantization.calib_block import assemble_batch
    gen = torch.Generator().manual_seed(9)
    N, n, h, w, c = 7, 3, 5, 6, 16
    inp = torch.randn(2, N, h, w, c, generator=gen).cuda().bfloat16()
    sym = torch.randn(2, N, h, w, c, generator=gen).cuda().bfloat16()
    idx = torch.tensor([6, 0, 3], dtype=torch.int32).cuda()
    out = torch.empty(2, n, h, w, c, device="cuda", dtype=torch.bfloat16)
    assemble_batch(inp, sym, idx, 1.0, out)
    assert torch.equal(out, inp[:, idx.long()])
    torch.manual_seed(4)
    assemble_batch(inp, sym, idx, 0.3, out)
    torch.manual_seed(4)
    keep = torch.rand_like(out[0], dtype=torch.float32) < 0.3
    want = torch.where(keep.unsqueeze(0), inp[:, idx.long()], sym[:, idx.long()])
    assert torch.equal(out, want) and 0.2 < float(keep.float().mean()) < 0.4
    st = L.stream()
    assert L.lib.nq_qdrop_gather(inp.data_ptr(), sym.data_ptr(), idx.data_ptr(), None, 0.5, n, N, h * w * c, out.data_ptr(), st) != 0
    assert L.lib.nq_qdrop_gather(inp.data_ptr(), None, idx.data_ptr(), None, 1.0, n, N, h * w * c + 4, out.data_ptr(), st) != 0
