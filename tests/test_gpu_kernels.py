"""GPU parity tests (run with -m gpu on a B200): every kernel behind the C ABI against the CPU oracle
(oracle/nq_oracle.py) and the golden fixtures produced by the unmodified reference.

Bars: bit-exact for integer work (hard / nearest codes, zero points, de-quantised values that are
products of exact integers and scales); fp32 tolerances stated per test for everything that goes
through transcendental functions (GPU expf/logf/erff differ from the CPU's by <= 2 ulp) or through
differently-ordered fp32 sums.
"""
import math

import numpy as np
import pytest
import torch

from oracle import nq_oracle as O
from tests.helpers import CASES, case_stages, load, t

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nq():
    import neuroquant_b200 as pkg
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return pkg


def dev(x):
    return x.cuda().contiguous()


# ------------------------------------------------------------------------------------------ quantisers
@pytest.mark.parametrize("bits", [2, 3, 4, 6, 8])
@pytest.mark.parametrize("name", ["w", "b"])
def test_uaq_kernels_golden(nq, bits, name):
    L = nq._lib
    g = load("quantizer_kats")
    x = dev(t(g[name]))
    delta, zp = L.uaq_init_max(x, bits, True)
    assert np.array_equal(delta.cpu().numpy(), g[f"uaq{bits}_{name}_delta"])  # bit-exact scales
    assert np.array_equal(zp.cpu().numpy(), g[f"uaq{bits}_{name}_zp"])
    codes, deq = L.fakequant_fwd(x, None, delta, zp, bits, 0)
    assert np.array_equal(deq.cpu().numpy(), g[f"uaq{bits}_{name}_deq"])  # bit-exact
    want_codes, _ = O.uaq_quant(t(g[name]), t(g[f"uaq{bits}_{name}_delta"]), t(g[f"uaq{bits}_{name}_zp"]), bits)
    assert torch.equal(codes.cpu(), want_codes)
    r = dev(t(g[f"uaq{bits}_{name}_r"]))
    dd = L.fakequant_bwd(r, x, None, delta, zp, bits, 0)
    want = g[f"uaq{bits}_{name}_ddelta"]
    # each term is (code - zp) - x/delta: a difference of numbers up to 2^bits, so the fp32 rounding
    # of a row sum is ~ eps * 2^bits * sqrt(row_len); the reference sums the two halves separately
    row_len = x.numel() // max(1, delta.numel())
    tol = 1e-6 * 2 ** bits * math.sqrt(row_len) + 1e-6
    assert np.abs(dd.cpu().numpy().reshape(want.shape) - want).max() <= tol


@pytest.mark.parametrize("method", ["mse", "l1", "gaussian"])
@pytest.mark.parametrize("bits", [2, 4, 6, 8])
def test_scale_search_initialisers_golden(nq, method, bits):
    """'mse' / 'l1' / 'gaussian' scale initialisers (quantizer.py:170-222) against the reference's own
    UniformAffineQuantizer(scale_method=...) (tests/golden/make_init_golden.py): the search picks the reference's range
    and the step sizes are bit-exact; the gaussian one within an ulp or two (its mean / variance are reductions)."""
    L = nq._lib
    g = load("scale_inits")
    for name in ("w", "b"):
        x = dev(t(g[name]))
        delta, zp = L.uaq_init_search(x, bits, True, method)
        ref_d, ref_z = g[f"{method}{bits}_{name}_delta"], g[f"{method}{bits}_{name}_zp"]
        got_d, got_z = delta.cpu().numpy().reshape(ref_d.shape), zp.cpu().numpy().reshape(ref_z.shape)
        if method == "gaussian":
            assert np.allclose(got_d, ref_d, rtol=1e-6, atol=0) and np.abs(got_z - ref_z).max() <= 1
        else:
            assert np.array_equal(got_d, ref_d), (name, np.abs(got_d - ref_d).max())
            assert np.array_equal(got_z, ref_z)
            _, deq = L.fakequant_fwd(x, None, delta, zp, bits, 0)
            assert np.array_equal(deq.cpu().numpy(), g[f"{method}{bits}_{name}_deq"])


def test_quantmodel_with_mse_init(nq):
    """--init mse end to end: the module binding initialises every layer through the quantiser's own scale_method."""
    from neuroquant_b200.models import HNeRV
    from neuroquant_b200.quantization import QuantModel, QuantModule
    from tests.helpers import TINY_HNERV
    torch.manual_seed(3)
    model = HNeRV(TINY_HNERV).cuda()
    qnn = QuantModel(model, hadamard=False, weight_quant_params={"n_bits": 8, "channel_wise": True, "scale_method": "mse"}).cuda()
    qnn.set_bitwidth([4] * 7)
    qnn.eval()
    qnn.set_quant_state(True)
    embed = torch.randn(2, 4, 2, 4).cuda()
    out, _, _ = qnn(embed)
    assert torch.isfinite(out).all()
    for m in qnn.model.modules():
        if isinstance(m, QuantModule):
            w = m.weight.detach().cpu()
            d, z = O.uaq_init(w, 4, True, "mse")
            assert torch.equal(m.weight_quantizer.delta.detach().cpu(), d) and torch.equal(m.weight_quantizer.zero_point.cpu(), z)


@pytest.mark.parametrize("bits", [2, 4, 6, 8])
@pytest.mark.parametrize("name", ["w", "b"])
def test_adaround_kernels_golden(nq, bits, name):
    L = nq._lib
    g = load("quantizer_kats")
    x = dev(t(g[f"ada{bits}_{name}_x"]))
    delta, zp = dev(t(g[f"ada{bits}_{name}_delta"])), dev(t(g[f"ada{bits}_{name}_zp"]))
    a0 = L.adaround_init_alpha(x, delta)
    assert np.allclose(a0.cpu().numpy(), g[f"ada{bits}_{name}_alpha0"], rtol=2e-6, atol=2e-6)  # logf ulp
    alpha = dev(t(g[f"ada{bits}_{name}_alpha"]))
    # hard rounding: integer codes, bit-exact (the deliverable)
    codes, deq = L.fakequant_fwd(x, alpha, delta, zp, bits, 2)
    assert np.array_equal(codes.cpu().numpy(), g[f"ada{bits}_{name}_hard_codes"])
    assert np.array_equal(deq.cpu().numpy(), g[f"ada{bits}_{name}_hard_deq"])
    # soft rounding: sigmoid through expf -> 1e-6
    reg = torch.zeros(1, device="cuda")
    codes, deq = L.fakequant_fwd(x, alpha, delta, zp, bits, 1, reg_sum=reg, reg_b=7.5)
    assert np.allclose(codes.cpu().numpy(), g[f"ada{bits}_{name}_soft_codes"], rtol=0, atol=2e-6 * 2 ** bits)
    assert np.allclose(deq.cpu().numpy(), g[f"ada{bits}_{name}_soft_deq"], rtol=1e-6, atol=1e-6)
    assert float(reg) == pytest.approx(float(g[f"ada{bits}_{name}_reg_b7.5"]), rel=1e-5)
    r = dev(t(g[f"uaq{bits}_{name}_r"]))
    da = L.fakequant_bwd(r, x, alpha, delta, zp, bits, 1, 1.0, 0.01, 7.5)
    want = g[f"ada{bits}_{name}_dalpha"]
    assert np.allclose(da.cpu().numpy(), want, rtol=2e-5, atol=1e-7)


def test_fakequant_big_random_bit_exact(nq):
    """A weight-sized tensor (HNeRV stage 3: 848 x 64 x 5 x 5): nearest and hard codes bit-exact."""
    L = nq._lib
    g = torch.Generator().manual_seed(5)
    w = torch.randn(848, 64, 5, 5, generator=g) * 0.05
    for bits in (4, 6):
        d, z = O.uaq_init_max(w, bits, True)
        dg, zg = L.uaq_init_max(dev(w), bits, True)
        assert torch.equal(dg.cpu(), d) and torch.equal(zg.cpu(), z)
        codes, deq = L.fakequant_fwd(dev(w), None, dg, zg, bits, 0)
        wc, wd = O.uaq_quant(w, d, z, bits)
        assert torch.equal(codes.cpu(), wc) and torch.equal(deq.cpu(), wd)
        d16, z16 = O.fp16_round(d), O.fp16_round(z)
        alpha = torch.randn(w.shape, generator=g) * 3
        codes, deq = L.fakequant_fwd(dev(w), dev(alpha), dev(d16), dev(z16), bits, 2)
        wc, wd = O.adaround_quant(w, alpha, d16, z16, bits, soft=False)
        assert torch.equal(codes.cpu(), wc) and torch.equal(deq.cpu(), wd)


def test_fakequant_rejects_bad_args(nq):
    L = nq._lib
    x = torch.zeros(4, 4, 1, 1, device="cuda")
    d = torch.ones(4, 1, 1, 1, device="cuda")
    with pytest.raises(L.NqError):
        L.fakequant_fwd(x, None, d, d, 9, 0)  # assert 2 <= n_bits <= 8 (quantizer.py:96)
    with pytest.raises(L.NqError):
        L.fakequant_fwd(x, None, d, d, 4, 1)  # AdaRound needs alpha
    with pytest.raises(L.NqError):
        L.fakequant_fwd(x.cpu(), None, d, d, 4, 0)  # no CPU fallback


@pytest.mark.parametrize("n", [1, 2, 16, 32, 64, 128, 256])
def test_fwht_rows(nq, n):
    L = nq._lib
    x = torch.randn(77, n, generator=torch.Generator().manual_seed(n))
    got = L.fwht_rows(dev(x)).cpu()
    want = O.fwht_last(x)
    assert torch.allclose(got, want, rtol=1e-6, atol=1e-6)
    from scipy.linalg import hadamard
    ref = x.double() @ t(hadamard(n).astype(np.float64)) / math.sqrt(n)
    assert torch.allclose(got.double(), ref, atol=1e-5)


def test_fwht_channel_golden_and_involution(nq):
    L = nq._lib
    g = load("quantizer_kats")
    got = L.fwht_channel(dev(t(g["had_in"])))
    assert np.allclose(got.cpu().numpy(), g["had_out"], rtol=1e-6, atol=1e-6)
    x = torch.randn(2, 8, 4, 4)
    assert (L.fwht_channel(L.fwht_channel(dev(x))).cpu() - x).abs().max() < 1e-6  # quant_layer.py:94-100
    with pytest.raises(L.NqError):
        L.fwht_rows(torch.zeros(3, 24, device="cuda"))  # not a power of two
    with pytest.raises(L.NqError):
        L.fwht_rows(torch.zeros(3, 512, device="cuda"))  # > 256


def test_adam_matches_torch(nq):
    L = nq._lib
    g = torch.Generator().manual_seed(1)
    p0 = torch.randn(1000, generator=g)
    p_ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([p_ref], lr=0.003)
    p, m, v = dev(p0.clone()), torch.zeros(1000, device="cuda"), torch.zeros(1000, device="cuda")
    for step in range(1, 6):
        gr = torch.randn(1000, generator=g)
        p_ref.grad = gr.clone()
        opt.step()
        L.adam_step(p, dev(gr), m, v, 0.003, step)
    assert torch.allclose(p.cpu(), p_ref.detach(), rtol=1e-6, atol=1e-7)


def test_loss_psnr_dot(nq):
    L = nq._lib
    g = load("quantizer_kats")
    pr, tg = t(g["lp_pred"]), t(g["lp_tgt"])
    n_mean = pr.shape[0] * pr.shape[2] * pr.shape[3]
    for p, key in ((2.0, "lp_p2"), (2.4, "lp_p24")):
        s, grad = L.lp_loss_sum(dev(pr), dev(tg), p, 1.0 / n_mean, want_grad=True)
        assert float(s) / n_mean == pytest.approx(float(g[key]), rel=2e-6)
        pa = pr.clone().requires_grad_(True)
        O.lp_loss(pa, tg, p).backward()
        assert torch.allclose(grad.cpu(), pa.grad, rtol=1e-5, atol=1e-8)
    a, b = torch.rand(3, 3, 20, 30), torch.rand(3, 3, 20, 30)
    assert torch.allclose(L.psnr(dev(a), dev(b)).cpu(), O.psnr(a, b), rtol=1e-6)
    xs = [torch.randn(n) for n in (5, 1000, 70001)]
    ys = [torch.randn(n) for n in (5, 1000, 70001)]
    got = L.multi_dot([dev(x) for x in xs], [dev(y) for y in ys], 0).cpu()
    want = torch.stack([(x.double() * y.double()).sum() for x, y in zip(xs, ys)]).float()
    assert torch.allclose(got, want, rtol=1e-4, atol=1e-3)
    got = L.multi_dot([dev(x) for x in xs], [dev(y) for y in ys], 1).cpu()
    want = torch.stack([((x.double() * y.double()) ** 2).sum() for x, y in zip(xs, ys)]).float()
    assert torch.allclose(got, want, rtol=1e-4)


# ------------------------------------------------------------------------------------------ decoder engine
@pytest.fixture(params=["tc", "simt"])
def conv_path(request, monkeypatch):
    """Both convolution paths of the library: tcgen05 tensor cores (default) and exact-fp32 FFMA."""
    monkeypatch.setenv("NQ_CONV", request.param)
    return request.param


def make_engine(nq, tag, mode="uaq"):
    g, arch, cfg, stages = case_stages(tag)
    geoms = nq.geometry_from_cfg(cfg, arch)
    had = bool(g["hadamard"])
    qs = [nq.QuantStage(gm, dev(st.weight), dev(st.bias), int(b), had) for gm, st, b in zip(geoms, stages, g["bits"].tolist())]
    eng = nq.DecoderEngine(qs)
    eng.mode = mode
    return g, arch, cfg, stages, eng


@pytest.mark.parametrize("tag", list(CASES))
def test_decode_golden(nq, tag, conv_path):
    """FP decode and nearest-rounded quantised decode against the reference's outputs; scales and
    perturbation-free codes bit-exact."""
    g, arch, cfg, stages, eng = make_engine(nq, tag, "off")
    assert eng.use_tc == (conv_path == "tc")
    cali = dev(t(g["cali"]))
    out = eng.forward(cali[:2]).cpu()
    assert np.abs(out.numpy() - g["fp_out"]).max() < 1e-5
    eng.mode = "uaq"
    eng.init_scales()
    assert eng.avg_bits() == float(g["avg_bits"])
    for i, s in enumerate(eng.stages):
        assert np.array_equal(s.delta_w.cpu().numpy(), g[f"init/{i}/delta_w"])
        assert np.array_equal(s.zp_w.cpu().numpy(), g[f"init/{i}/zp_w"])
        assert np.array_equal(s.delta_b.cpu().numpy(), g[f"init/{i}/delta_b"])
        assert np.array_equal(s.zp_b.cpu().numpy(), g[f"init/{i}/zp_b"])
    out = eng.forward(cali[:2]).cpu()
    assert np.abs(out.numpy() - g["uaq_out"]).max() < 1e-5
    # codes of the last forward (quant_model.py:74-80) against the oracle, bit-exact
    qd = O.QuantDecoder(stages, g["bits"].tolist(), bool(g["hadamard"]))
    with torch.no_grad():
        qd.forward(t(g["cali"])[:2])
    for s, q in zip(eng.stages, qd.q):
        if not bool(g["hadamard"]):  # rotated inputs differ by fp32 rounding of the WHT -> ties may flip
            assert torch.equal(s.codes_w.cpu(), q.codes_w)
        else:
            assert (s.codes_w.cpu() != q.codes_w).float().mean() < 1e-3
        assert torch.equal(s.codes_b.cpu(), q.codes_b)


def oracle_grads(stages, g, soft, batch=slice(0, 2)):
    """Oracle: loss gradients w.r.t. alpha (soft AdaRound) or delta (UAQ) on the first two frames."""
    qd = O.QuantDecoder(stages, g["bits"].tolist(), bool(g["hadamard"]))
    cali, frames = t(g["cali"]), t(g["frames"])
    if soft:
        qd.start_adaround()
        leaves = []
        for q in qd.q:
            q.alpha_w.requires_grad_(True)
            q.alpha_b.requires_grad_(True)
            leaves += [q.alpha_w, q.alpha_b]
    else:
        leaves = []
        for q in qd.q:
            q.delta_w = q.delta_w.clone().requires_grad_(True)
            q.delta_b = q.delta_b.clone().requires_grad_(True)
            leaves += [q.delta_w, q.delta_b]
    out = qd.forward(cali[batch])
    rec = O.lp_loss(out, frames[batch], 2.0)
    reg = sum(0.01 * O.round_reg(q.alpha_w, 7.5) for q in qd.q) if soft else 0.0
    (rec + reg).backward()
    return qd, out.detach(), float(rec), [x.grad for x in leaves]


@pytest.mark.parametrize("tag", list(CASES))
@pytest.mark.parametrize("soft", [True, False])
def test_backward_matches_oracle_autograd(nq, tag, soft, conv_path):
    """forward + loss + backward + quantiser Jacobian vs autograd of the oracle.  Tolerance: 2e-4 of
    each tensor's gradient scale (fp32 sums in a different order through 7 convolutions)."""
    g, arch, cfg, stages, eng = make_engine(nq, tag, "uaq")
    eng.init_scales()
    qd, want_out, want_rec, want = oracle_grads(stages, g, soft)
    if soft:
        eng.start_adaround()
    cali, frames = dev(t(g["cali"])), dev(t(g["frames"]))
    out = eng.forward(cali[:2], train=True, target=frames[:2], p_norm=2.0, reg_b=7.5 if soft else None)
    assert (out.cpu() - want_out).abs().max() < 2e-5
    assert float(eng.last_loss()) == pytest.approx(want_rec, rel=1e-5)
    eng.backward()
    got = eng.param_grads(1.0, 0.01 if soft else 0.0, 7.5 if soft else 0.0)
    got = [x for pair in got for x in pair]
    assert len(got) == len(want)
    for i, (a, b) in enumerate(zip(got, want)):
        a = a.cpu().reshape(b.shape)
        scale = b.abs().max().item() + 1e-12
        # d_delta terms are (code - zp) - x/delta, differences of numbers up to 2^bits: looser bar
        rtol = 2e-4 if soft else 2e-3
        assert (a - b).abs().max().item() <= rtol * scale + 1e-9, (i, (a - b).abs().max().item(), scale)


@pytest.mark.parametrize("tag", list(CASES))
def test_calibration_golden(nq, tag, conv_path):
    """80 iterations (4 step-size + 76 AdaRound) in the reference's injected batch order: same
    acceptance as the oracle's own pin (tests/test_oracle_golden.py::test_calibration_golden)."""
    g, arch, cfg, stages, eng = make_engine(nq, tag, "uaq")
    eng.init_scales()
    cali, frames = dev(t(g["cali"])), dev(t(g["frames"]))
    order = g["order"].tolist()
    log = []

    def fetch(idx):
        idx = torch.as_tensor(idx, device="cuda")
        return cali[idx], frames[idx]

    loop = nq.CalibrationLoop(eng, fetch, len(order), iters=80, weight=0.01, b_range=(20, 2), warmup=0.2, p=2.0,
                              lr=0.003, log=log)
    loop.run(lambda: order)
    traj = g["traj"]
    assert len(log) == len(traj)
    got_total = np.array([r[2] + r[3] for r in log])
    assert np.allclose(got_total[:10], traj[:10, 1], rtol=1e-4, atol=1e-7)
    assert np.allclose(got_total, traj[:, 1], rtol=5e-3, atol=1e-6)
    out = eng.forward(cali[:2]).cpu()
    assert np.abs(out.numpy() - g["calib_out"]).max() < 5e-3
    assert np.abs(O.psnr(out, t(g["frames"])[:2]).numpy() - g["calib_psnr"]).max() < 0.01
    # The trajectory is chaotic (one floor()/clamp flip or the sign of a ~0 step-size gradient is
    # amplified by Adam), so element-wise agreement is asserted on aggregate fractions.
    n_diff = n_tot = a_far = a_tot = d_far = d_tot = 0
    for i, s in enumerate(eng.stages):
        a_far += int((np.abs(s.alpha_w.cpu().numpy() - g[f"final/{i}/alpha_w"]) > 1e-3).sum())
        a_tot += s.alpha_w.numel()
        want_d = g[f"final/{i}/delta_w"]
        d_far += int((np.abs(s.delta_w.cpu().numpy() - want_d) > 2e-3 * np.abs(want_d)).sum())
        d_tot += want_d.size
        cw = s.codes_w.cpu()
        n_diff += int((cw.numpy() != g[f"final/{i}/codes_w"]).sum())
        n_tot += cw.numel()
        assert torch.equal(cw, cw.round())  # hard codes are integers
        assert np.allclose(s.codes_b.cpu().numpy(), g[f"final/{i}/codes_b"], atol=0.15)  # biases stay soft (Q3)
        # identical V and scales -> bit-exact codes: re-derive on the oracle from OUR alpha/delta
        wc, _ = O.adaround_quant(s.w_src.cpu(), s.alpha_w.cpu(), s.delta_w.cpu(), s.zp_w.cpu(), s.n_bits, soft=False)
        assert torch.equal(cw, wc)
    # a step size that differs by 1e-4 relative (the reference's own d_delta is two large sums that
    # cancel, so it is only defined to ~1e-3) moves frac(x/delta), hence every alpha of that channel
    assert a_far / a_tot < 0.5, (a_far, a_tot, n_diff, n_tot)
    assert d_far / d_tot < 0.08, (d_far, d_tot)
    assert n_diff / n_tot < 1e-2, (n_diff, n_tot)


def test_graph_replay_equals_eager(nq, monkeypatch):
    """The CUDA-graph replay of the AdaRound iteration (calibration.GraphedStep) is the same kernel sequence as the
    eager one: identical alpha after 30 iterations, bit for bit."""
    res = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("NQ_GRAPH", flag)
        g, arch, cfg, stages, eng = make_engine(nq, "tiny_hnerv", "uaq")
        eng.init_scales()
        cali, frames = dev(t(g["cali"])), dev(t(g["frames"]))
        order = g["order"].tolist()

        def fetch(idx):
            idx = torch.as_tensor(idx, device="cuda")
            return cali[idx], frames[idx]

        loop = nq.CalibrationLoop(eng, fetch, len(order), iters=40, weight=0.01, b_range=(20, 2), warmup=0.2, p=2.0, lr=0.003)
        assert loop.use_graph == (flag == "1")
        loop.run(lambda: order)
        res[flag] = [s.alpha_w.clone() for s in eng.stages] + [s.alpha_b.clone() for s in eng.stages]
        if flag == "1":
            assert loop._graphed, "graph path was not taken"
    for a, b in zip(res["1"], res["0"]):
        assert torch.equal(a, b)


def test_tensor_core_head_variant(nq, monkeypatch):
    """The three head forward kernels (FFMA strip kernel, generic tcgen05 kernel with the head epilogue, tap-expanded
    tcgen05 kernel = default) give the same frames, loss and gradients."""
    outs = {}
    for mode in ("simt", "tc", "tapexp"):
        monkeypatch.setenv("NQ_HEAD", mode)
        g, arch, cfg, stages, eng = make_engine(nq, "tiny_hnerv", "uaq")
        assert eng.head_tc == (mode == "tc") and eng.head_tapexp == (mode == "tapexp")
        eng.init_scales()
        eng.start_adaround()
        cali, frames = dev(t(g["cali"])), dev(t(g["frames"]))
        img = eng.forward(cali[:2], train=True, target=frames[:2]).clone()
        loss = float(eng.last_loss())
        eng.backward()
        grads = [x.clone() for pair in eng.param_grads() for x in pair]
        outs[mode] = (img, loss, grads)
    for mode in ("tc", "tapexp"):
        assert (outs[mode][0] - outs["simt"][0]).abs().max() < 2e-6
        assert outs[mode][1] == pytest.approx(outs["simt"][1], rel=1e-5)
        for a, b in zip(outs[mode][2], outs["simt"][2]):
            assert (a - b).abs().max() <= 2e-4 * b.abs().max() + 1e-10


def test_saved_activation_derivative_equals_recomputed(nq, monkeypatch, conv_path):
    """nq_conv_desc.act = 2 (forward keeps GELU'(z) in the z buffer, dgrad multiplies) against act = 1 (forward keeps
    z, dgrad re-evaluates GELU'): same frames, and the same gradients up to fp32 rounding of one product."""
    outs = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("NQ_SAVE_ACT_GRAD", flag)
        g, arch, cfg, stages, eng = make_engine(nq, "tiny_hnerv", "uaq")
        eng.init_scales()
        eng.start_adaround()
        cali, frames = dev(t(g["cali"])), dev(t(g["frames"]))
        img = eng.forward(cali[:2], train=True, target=frames[:2]).clone()
        acts = [d.act for d in eng._last_plan.desc]
        assert (2 in acts) == (flag == "1") and (1 in acts) == (flag == "0")
        eng.backward()
        outs[flag] = (img, [x.clone() for pair in eng.param_grads() for x in pair])
    assert torch.equal(outs["1"][0], outs["0"][0])
    for a, b in zip(outs["1"][1], outs["0"][1]):
        assert (a - b).abs().max() <= 1e-5 * b.abs().max() + 1e-12


def test_multi_tensor_quantiser_launches_equal_single_tensor(nq):
    """nq_fakequant_fwd_multi and nq_adaround_step_multi against the single-tensor entry points they batch:
    bit-identical outputs on ragged sizes (a 3-element bias, a tensor that is not a multiple of the block chunk,
    per-tensor and channel-wise scales, 20 tensors = two launches)."""
    import ctypes as C
    L = nq._lib
    torch.manual_seed(5)
    shapes = [(3,), (16, 3, 3, 3), (7, 300), (64, 48, 5, 5), (1, 2049), (5,), (33, 7)] * 3
    shapes = shapes[:20]
    hyper = dev(torch.tensor([0.01, 7.5, 0.003 / (1 - 0.9 ** 3), (1 - 0.999 ** 3) ** 0.5]))
    single, multi, fq_tasks, ada_tasks, keep = [], [], [], [], []
    for k, shp in enumerate(shapes):
        x = dev(torch.randn(shp))
        cw = len(shp) > 1 and k % 2 == 0
        bits = 2 + k % 7
        rows, row_len = (shp[0], x.numel() // shp[0]) if cw else (1, x.numel())
        delta = dev(torch.rand(rows) * 0.05 + 0.01)
        zp = dev(torch.randint(0, 2 ** bits, (rows,)).float())
        alpha = dev(torch.randn(shp))
        g = dev(torch.randn(shp))
        mode = k % 3  # nq_round_mode: nearest, soft, hard
        # single-tensor reference
        c1, d1 = torch.empty_like(x), torch.empty_like(x)
        L.check(L.lib.nq_fakequant_fwd(L.ptr(x), L.ptr(alpha), L.ptr(delta), L.ptr(zp), rows, row_len, int(cw), bits, mode,
                                       L.ptr(c1), L.ptr(d1), None, 0.0, L.stream()))
        a1, m1, v1 = alpha.clone(), dev(torch.rand(shp) * 1e-3), dev(torch.rand(shp) * 1e-6)
        a2, m2, v2 = a1.clone(), m1.clone(), v1.clone()
        da = torch.empty_like(x)
        L.check(L.lib.nq_fakequant_bwd_soft_dev(L.ptr(g), L.ptr(x), L.ptr(a1), L.ptr(delta), L.ptr(zp), rows, row_len, int(cw), bits,
                                                1.0, k % 2, L.ptr(hyper), L.ptr(da), L.stream()))
        L.check(L.lib.nq_adam_step_dev(L.ptr(a1), L.ptr(da), L.ptr(m1), L.ptr(v1), x.numel(), 0.9, 0.999, 1e-8, L.ptr(hyper), L.stream()))
        single.append((c1, d1, a1, m1, v1))
        c2, d2 = torch.empty_like(x), torch.empty_like(x)
        fq_tasks.append(L.FqTask(L.ptr(x), L.ptr(alpha), L.ptr(delta), L.ptr(zp), L.ptr(c2), L.ptr(d2), rows, row_len, int(cw), bits, mode, 0))
        ada_tasks.append(L.AdaTask(L.ptr(g), L.ptr(x), L.ptr(a2), L.ptr(delta), L.ptr(zp), L.ptr(m2), L.ptr(v2), rows, row_len, int(cw),
                                   bits, k % 2, 0))
        multi.append((c2, d2, a2, m2, v2))
        keep.append((x, delta, zp, alpha, g))
    L.check(L.lib.nq_fakequant_fwd_multi((L.FqTask * len(fq_tasks))(*fq_tasks), len(fq_tasks), None, 0.0, L.stream()))
    L.check(L.lib.nq_adaround_step_multi((L.AdaTask * len(ada_tasks))(*ada_tasks), len(ada_tasks), 1.0, 0.9, 0.999, 1e-8,
                                         L.ptr(hyper), L.stream()))
    torch.cuda.synchronize()
    for s_, m_ in zip(single, multi):
        for a, b in zip(s_, m_):
            assert torch.equal(a, b)
    assert L.lib.nq_fakequant_fwd_multi(None, 3, None, 0.0, L.stream()) != 0


def test_host_batch_pipe_orders_and_reuses_slots(nq):
    """HostBatchPipe hands batches out in the order they were put, from pinned host memory, and may be refilled
    while the previous batch is still being consumed (two slots)."""
    from neuroquant_b200.calibration import HostBatchPipe
    pipe = HostBatchPipe([((2, 3, 4, 5), torch.float32), ((2, 3, 8, 8), torch.float32)])
    # pinned sources are copied directly, unpinned ones (odd k) through the pipe's pinned staging slots
    host = [(torch.full((2, 3, 4, 5), float(k)), torch.full((2, 3, 8, 8), float(-k))) for k in range(7)]
    host = [tuple(t_.pin_memory() for t_ in pair) if k % 2 == 0 else pair for k, pair in enumerate(host)]
    pipe.put(*host[0])
    acc = torch.zeros((), device="cuda")
    for k in range(7):
        e, f = pipe.get()
        if k + 1 < 7:
            pipe.put(*host[k + 1])
        acc = acc + e.sum() * 1000 + f.sum()      # consumer work enqueued on the compute stream
        assert torch.equal(e.cpu(), host[k][0]) and torch.equal(f.cpu(), host[k][1])
    pipe.release()
    want = sum(k * 120 * 1000 - k * 384 for k in range(7))
    assert float(acc) == want
    with pytest.raises(RuntimeError):
        pipe.get()


def test_head_weight_gradient_variants_agree(nq, monkeypatch):
    """Tap-expanded head weight gradient (default, nq_head_wgrad_tapexp) against the generic tensor-core wgrad kernel: same
    gradient for every tensor (only the head's path differs; the tolerance is fp32 summation order)."""
    outs = {}
    for mode in ("generic", "tapexp"):
        monkeypatch.setenv("NQ_HEAD_WG", mode)
        g, arch, cfg, stages, eng = make_engine(nq, "tiny_hnerv", "uaq")
        assert eng.head_wg_tapexp == (mode == "tapexp")
        eng.init_scales()
        eng.start_adaround()
        cali, frames = dev(t(g["cali"])), dev(t(g["frames"]))
        eng.forward(cali[:2], train=True, target=frames[:2])
        eng.backward()
        outs[mode] = [(gw.clone(), gb.clone()) for gw, gb in eng._grad_buffers()[1]]
    for (gw_a, gb_a), (gw_b, gb_b) in zip(outs["tapexp"], outs["generic"]):
        assert (gw_a - gw_b).abs().max() <= 1e-5 * gw_b.abs().max() + 1e-12
        assert (gb_a - gb_b).abs().max() <= 1e-5 * gb_b.abs().max() + 1e-12


def test_uint8_targets_equal_float_targets(nq, conv_path):
    """Frames as the data set stores them (uint8; videosets/datasets.py:8-54 divides by 255 on the host): the ingest
    kernel reproduces `u8 / 255.0` bit for bit, and a training forward + backward fed with uint8 targets gives the same
    loss and the same gradients (bit for bit) as one fed with the float frames."""
    from neuroquant_b200 import _lib as L
    gen = torch.Generator().manual_seed(5)
    u8 = torch.randint(0, 256, (2, 3, 32, 64), generator=gen, dtype=torch.uint8)
    want = u8 / 255.0                                  # the reference's host-side conversion
    u8_d = u8.cuda()
    got = torch.empty(u8.shape, device="cuda")
    L.check(L.lib.nq_u8_to_f32(u8_d.data_ptr(), L.ptr(got), u8_d.numel(), L.stream()), "nq_u8_to_f32")
    assert torch.equal(got.cpu(), want)
    odd = torch.arange(0, 256, dtype=torch.uint8).repeat(3)[:613].cuda()          # all 256 values, ragged tail
    got = torch.empty(613, device="cuda")
    L.check(L.lib.nq_u8_to_f32(odd.data_ptr(), L.ptr(got), 613, L.stream()), "nq_u8_to_f32")
    assert torch.equal(got.cpu(), odd.cpu() / 255.0)
    res = []
    for tgt in (want.cuda(), u8_d):
        g, arch, cfg, stages, eng = make_engine(nq, "tiny_hnerv", "uaq")
        eng.init_scales()
        eng.start_adaround()
        cali = dev(t(g["cali"]))
        eng.forward(cali[:2], train=True, target=tgt, p_norm=2.0)
        loss = float(eng.last_loss())
        flat = eng.backward().clone()
        res.append((loss, flat))
    assert res[0][0] == pytest.approx(res[1][0], rel=1e-6)   # the loss is an atomic sum over CTAs: order-dependent last bits
    assert torch.equal(res[0][1], res[1][1])


def test_step_size_phase_graph_replay_equals_eager(nq, monkeypatch):
    """Phase 1 (step sizes, straight-through rounding) replayed as a CUDA graph is the eager kernel sequence: identical
    delta after 2 epochs, bit for bit, and phase 2 continues from it identically."""
    res = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("NQ_GRAPH", flag)
        g, arch, cfg, stages, eng = make_engine(nq, "tiny_hnerv", "uaq")
        eng.init_scales()
        cali, frames = dev(t(g["cali"])), dev(t(g["frames"]))
        order = g["order"].tolist()

        def fetch(idx):
            idx = torch.as_tensor(idx, device="cuda")
            return cali[idx], frames[idx]

        loop = nq.CalibrationLoop(eng, fetch, len(order), iters=160, weight=0.01, b_range=(20, 2), warmup=0.2, p=2.0, lr=0.003)
        assert loop.ep1 == 2
        n1 = loop.run_phase1(lambda: order)
        assert n1 == 8
        res[flag] = [s.delta_w.clone() for s in eng.stages] + [s.delta_b.clone() for s in eng.stages]
        if flag == "1":
            assert any(gs.phase == "delta" and gs.graph is not None for gs in loop._graphed.values()), "graph path was not taken"
        loop.ep2 = 3
        loop.run_phase2(lambda: order)
        res[flag] += [s.alpha_w.clone() for s in eng.stages]
    for a, b in zip(res["1"], res["0"]):
        assert torch.equal(a, b)
