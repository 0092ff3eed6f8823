"""CPU oracle for the NeuroQuant post-training-quantisation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in `neuroquant_b200/` may import this module; only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may.  It is a plain PyTorch-CPU (fp32) restatement of the reference's algorithm, written
as pure functions over explicit tensors (no nn.Module surgery), each citing the reference
file:line it follows (paths relative to /root/reference).

Pinning: `tests/golden/make_golden.py` (and make_block_golden.py, make_init_golden.py,
make_regress_golden.py next to it) import the UNMODIFIED reference from /root/reference (with the
three shim modules under `oracle/ref_shims/`) in the build container and store its outputs in
`tests/golden/*.npz`; `tests/test_oracle_golden.py` checks every function below against those
fixtures.  Two fixtures are not plain calls of a reference entry point and say so in their
generators: layer_reconstruction is the reference's own source compiled in memory with its missing
`opt_params = []` inserted (it cannot run as shipped), and the FP32 training loop replays
regress.py:249-271 with the reference's model, loss_fn and adjust_lr (its train() needs a PNG data
set, worker processes and TensorBoard).  The only un-pinned piece is the third-party
`hadamard_transform` package (PyPI `hadamard-transform`, version not pinned by the reference,
absent from /root/reference): restated as the orthonormal Sylvester-ordered Walsh-Hadamard
transform and checked against `scipy.linalg.hadamard(n)/sqrt(n)`; the ordering of codes along
the rotated channel axis is therefore "parity unpinned" (weight-space results do not depend on
it).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

ZETA, GAMMA = 1.1, -0.1  # quantizer.py:274


# --------------------------------------------------------------------------------------------
# Walsh-Hadamard rotation (quant_layer.py:13-22, third-party hadamard_transform at :7,:19)
# --------------------------------------------------------------------------------------------
def next_pow2(n: int) -> int:
    """quant_layer.py:13-14."""
    return 1 if n == 0 else 2 ** math.ceil(math.log2(n))


def fwht_last(x: torch.Tensor) -> torch.Tensor:
    """Orthonormal Sylvester WHT along the last dim (what quant_layer.py:19 relies on)."""
    n = x.shape[-1]
    assert n & (n - 1) == 0
    y = x.reshape(-1, n).clone()
    h = 1
    while h < n:
        y = y.view(-1, n // (2 * h), 2, h)
        y = torch.stack((y[:, :, 0] + y[:, :, 1], y[:, :, 0] - y[:, :, 1]), dim=2).reshape(-1, n)
        h *= 2
    return (y / math.sqrt(n)).view(x.shape)


def hadamard_along_channel(w: torch.Tensor) -> torch.Tensor:
    """quant_layer.py:16-22: WHT over C_in of a (C_out, C_in, KH, KW) weight."""
    co, ci, kh, kw = w.shape
    rows = w.permute(0, 2, 3, 1).reshape(-1, ci)
    rows = fwht_last(rows)
    return rows.view(co, kh, kw, ci).permute(0, 3, 1, 2).contiguous()


def rotate_weight(w: torch.Tensor) -> torch.Tensor:
    """quant_layer.py:43-49: zero-pad C_in to a power of two, then rotate."""
    ci = w.shape[1]
    pad = next_pow2(ci) - ci
    return hadamard_along_channel(F.pad(w, (0, 0, 0, 0, 0, pad)))


# --------------------------------------------------------------------------------------------
# Uniform affine quantiser (quantizer.py:76-243)
# --------------------------------------------------------------------------------------------
def _scale_max_1(x: torch.Tensor, n_levels: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """quantizer.py:153-168 ('max', asymmetric): python-float min/max -> fp32 tensors."""
    x_min = min(x.min().item(), 0)
    x_max = max(x.max().item(), 0)
    delta = torch.tensor((x_max - x_min) / (n_levels - 1))  # python double -> fp32 tensor
    delta = torch.max(delta, torch.tensor(1e-8, dtype=torch.float32))
    zp = (-x_min / delta).round()
    return delta.to(x.dtype), zp.to(x.dtype)


def uaq_init_max(x: torch.Tensor, n_bits: int, channel_wise: bool = True):
    """quantizer.py:127-152: per-out-channel for 4-D weights, per-tensor for 1-D biases."""
    n_levels = 2 ** n_bits
    if channel_wise and x.dim() == 4:
        pairs = [_scale_max_1(x[c], n_levels) for c in range(x.shape[0])]
        delta = torch.stack([p[0] for p in pairs]).view(-1, 1, 1, 1)
        zp = torch.stack([p[1] for p in pairs]).view(-1, 1, 1, 1)
        return delta, zp
    d, z = _scale_max_1(x, n_levels)
    if channel_wise:
        return d.view(-1), z.view(-1)
    return d, z


def _scale_search_1(x: torch.Tensor, n_bits: int, method: str):
    """quantizer.py:170-187 ('mse': L_3.5) and :204-220 ('l1') for one channel / tensor: ten shrinking ranges, the first
    strictly better score wins; all arithmetic in fp32 tensors (no clamp of the range to zero, unlike 'max')."""
    eps = torch.tensor(1e-8)
    n_levels = 2 ** n_bits
    x_max, x_min = x.max(), x.min()
    best, delta, zp = 1e+10, None, None
    for i in range(10):
        new_max = x_max * (1.0 - (i * 0.05))
        new_min = x_min * (1.0 - (i * 0.05))
        d = torch.max((new_max - new_min) / (2 ** n_bits - 1), eps)      # quantize(), quantizer.py:224-231
        z = (-new_min / d).round()
        x_q = (torch.clamp(torch.round(x / d) + z, 0, n_levels - 1) - z) * d
        score = (x - x_q).abs().pow(3.5).mean() if method == "mse" else (x - x_q).abs().mean()
        if score < best:
            best, delta, zp = score, d, z
    return delta, zp


def _scale_gaussian_1(x: torch.Tensor, n_bits: int):
    """quantizer.py:189-202: range mu +- 6 * VARIANCE (sic), clamped to contain zero."""
    n_levels = 2 ** n_bits
    mu, sigma = torch.mean(x), torch.var(x)
    x_min = min(mu - 6 * sigma, 0)
    x_max = max(mu + 6 * sigma, 0)
    delta = torch.max(torch.as_tensor((x_max - x_min) / (n_levels - 1), dtype=torch.float32), torch.tensor(1e-8))
    zp = torch.as_tensor(-x_min / delta, dtype=torch.float32).round()
    return delta, zp


def uaq_init(x: torch.Tensor, n_bits: int, channel_wise: bool, method: str):
    """init_quantization_scale for scale_method 'max' | 'mse' | 'l1' | 'gaussian', asymmetric (quantizer.py:127-222)."""
    if method == "max":
        return uaq_init_max(x, n_bits, channel_wise)
    one = (lambda t: _scale_gaussian_1(t, n_bits)) if method == "gaussian" else (lambda t: _scale_search_1(t, n_bits, method))
    if channel_wise and x.dim() == 4:
        pairs = [one(x[c]) for c in range(x.shape[0])]
        return (torch.stack([p[0] for p in pairs]).view(-1, 1, 1, 1), torch.stack([p[1] for p in pairs]).view(-1, 1, 1, 1))
    d, z = one(x)
    return (d.view(-1), z.view(-1)) if channel_wise else (d, z)


def round_ste(x: torch.Tensor) -> torch.Tensor:
    """quantizer.py:53-57."""
    return (x.round() - x).detach() + x


def uaq_quant(x, delta, zp, n_bits: int):
    """quantizer.py:117-119.  Returns (codes, dequantised)."""
    n_levels = 2 ** n_bits
    codes = torch.clamp(round_ste(x / delta) + zp, 0, n_levels - 1)
    return codes, (codes - zp) * delta


# --------------------------------------------------------------------------------------------
# AdaRound quantiser (quantizer.py:247-323)
# --------------------------------------------------------------------------------------------
def fp16_round(t: torch.Tensor) -> torch.Tensor:
    """quantizer.py:264-265: delta / zero_point pass through fp16 when AdaRound starts."""
    return t.detach().half().float()


def adaround_init_alpha(x: torch.Tensor, delta: torch.Tensor) -> torch.Tensor:
    """quantizer.py:305-313: alpha s.t. the rectified sigmoid equals the fractional part."""
    q = x / delta
    rest = q - torch.floor(q)
    return -torch.log((ZETA - GAMMA) / (rest - GAMMA) - 1)


def soft_targets(alpha: torch.Tensor) -> torch.Tensor:
    """quantizer.py:302-303."""
    return torch.clamp(torch.sigmoid(alpha) * (ZETA - GAMMA) + GAMMA, 0, 1)


def adaround_quant(x, alpha, delta, zp, n_bits: int, soft: bool):
    """quantizer.py:288-300 ('learned_hard_sigmoid').  Returns (codes, dequantised)."""
    n_levels = 2 ** n_bits
    x_floor = torch.floor(x / delta)
    x_int = x_floor + (soft_targets(alpha) if soft else (alpha >= 0).float())
    codes = torch.clamp(x_int + zp, 0, n_levels - 1)
    return codes, (codes - zp) * delta


# --------------------------------------------------------------------------------------------
# Losses and schedules (quantizer.py:66-73, calib_model.py:16-89, data_utils.py:24-41)
# --------------------------------------------------------------------------------------------
def lp_loss(pred, tgt, p: float = 2.0, reduction: str = "none"):
    if reduction == "none":
        return (pred - tgt).abs().pow(p).sum(1).mean()
    return (pred - tgt).abs().pow(p).mean()


def round_reg(alpha: torch.Tensor, b: float) -> torch.Tensor:
    """calib_model.py:44-45 (one tensor, before the `weight` factor)."""
    return (1 - ((soft_targets(alpha) - 0.5).abs() * 2).pow(b)).sum()


class LinearTempDecay:
    """data_utils.py:24-41."""

    def __init__(self, t_max, rel_start_decay=0.2, start_b=10, end_b=2):
        self.t_max = t_max
        self.start_decay = rel_start_decay * t_max
        self.start_b = start_b
        self.end_b = end_b

    def __call__(self, t):
        if t < self.start_decay:
            return self.start_b
        rel_t = (t - self.start_decay) / (self.t_max - self.start_decay)
        return self.end_b + (self.start_b - self.end_b) * max(0.0, 1 - rel_t)


def psnr(out: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    """utils.py:148-151, per frame."""
    mse = ((out - gt) ** 2).flatten(1).mean(1)
    return -10 * torch.log10(mse + 1e-9)


# --------------------------------------------------------------------------------------------
# Decoder description: every decoder stage is conv(k, same) -> shuffle(rh, rw) -> activation
# (HNeRV.py:29-42,49-71; NeRV.py:23-38,44-65; _layers.py:10-36)
# --------------------------------------------------------------------------------------------
@dataclass
class Stage:
    weight: torch.Tensor  # (C_out, C_in, k, k) reference layout
    bias: torch.Tensor  # (C_out,)
    rh: int = 1  # up-shuffle factors (PixelShuffle r, or the stem's fc_h/fc_w fold)
    rw: int = 1
    act: str = "none"  # 'none' | 'gelu' | 'tanh' (OutImg: 0.5*tanh+0.5)

    @property
    def k(self) -> int:
        return self.weight.shape[-1]


def decoder_geometry(cfg: dict, arch: str):
    """Per-stage (c_in, c_out, k, rh, rw, act) from a reference YAML config."""
    arch = arch.lower()
    strides = list(cfg["dec_strides"])
    geo = []
    if arch == "hnerv":
        import numpy as np

        fc = int(np.prod(cfg["enc_strides"]) // np.prod(cfg["dec_strides"]))
        c = cfg["dec_in_channel"]
        geo.append((cfg["enc_channel"][-1], c, 1, fc, fc, "none"))  # HNeRV.py:32,57
    elif arch == "nerv":
        import numpy as np

        fch = cfg["crop_h"] // int(np.prod(strides))
        fcw = cfg["crop_w"] // int(np.prod(strides))
        c = cfg["dec_in_channel"]
        geo.append((int(cfg["level"] * 2), c * fch * fcw, 1, fch, fcw, "none"))  # NeRV.py:26,51
    else:
        raise ValueError(arch)
    for ks, s in zip(cfg["dec_kernels"], strides):
        co = int(max(round(c / cfg["channel_reduce"]), cfg["channel_lbound"]))
        geo.append((c, co * s * s, ks, s, s, cfg["dec_acts"]))
        c = co
    geo.append((c, 3, 3, 1, 1, cfg["out_bias"]))
    return geo


def stages_from_state_dict(sd: dict, cfg: dict, arch: str) -> List[Stage]:
    """Pick decoder/head tensors out of a reference state_dict (keys per SURVEY section 5)."""
    geo = decoder_geometry(cfg, arch)
    n_blocks = len(cfg["dec_kernels"])
    names = ["decoder.0"] + [f"decoder.{i}.conv.0" for i in range(1, n_blocks + 1)] + ["head_layer"]
    out = []
    for (ci, co, k, rh, rw, act), nm in zip(geo, names):
        w, b = sd[nm + ".weight"].detach().float().cpu(), sd[nm + ".bias"].detach().float().cpu()
        assert tuple(w.shape) == (co, ci, k, k), (nm, tuple(w.shape), (co, ci, k, k))
        out.append(Stage(w.clone(), b.clone(), rh, rw, act))
    return out


def up_shuffle(x: torch.Tensor, rh: int, rw: int) -> torch.Tensor:
    """out[n,c,h*rh+i,w*rw+j] = in[n,c*rh*rw+i*rw+j,h,w]  (nn.PixelShuffle for rh==rw;
    the stem fold of HNeRV.py:57 / NeRV.py:51 in general)."""
    if rh == 1 and rw == 1:
        return x
    n, c, h, w = x.shape
    return x.view(n, -1, rh, rw, h, w).permute(0, 1, 4, 2, 5, 3).reshape(n, -1, rh * h, rw * w)


def apply_act(x: torch.Tensor, act: str) -> torch.Tensor:
    if act == "none":
        return x
    if act == "gelu":
        return F.gelu(x)  # exact erf, _layers.py:105
    if act == "tanh":
        return torch.tanh(x) * 0.5 + 0.5  # _layers.py:13-14
    if act == "sigmoid":
        return torch.sigmoid(x)
    raise ValueError(act)


def decode(stages: Sequence[Stage], embed: torch.Tensor,
           weights: Optional[Sequence[torch.Tensor]] = None,
           biases: Optional[Sequence[torch.Tensor]] = None,
           keep: bool = False):
    """HNeRV.decode / NeRV.decode with (optionally substituted) weights."""
    x = embed
    feats = []
    for i, st in enumerate(stages):
        w = st.weight if weights is None else weights[i]
        b = st.bias if biases is None else biases[i]
        x = F.conv2d(x, w, b, stride=1, padding=st.k // 2)  # quant_layer.py:80
        x = apply_act(up_shuffle(x, st.rh, st.rw), st.act)
        if keep:
            feats.append(x)
    return (x, feats) if keep else x


# --------------------------------------------------------------------------------------------
# Quantised decoder state (QuantModel / QuantModule bookkeeping, quant_model.py, quant_layer.py)
# --------------------------------------------------------------------------------------------
@dataclass
class QStage:
    stage: Stage
    n_bits: int
    hadamard: bool
    w_src: torch.Tensor = None  # tensor that is quantised: weight, or rotated+padded weight
    delta_w: torch.Tensor = None
    zp_w: torch.Tensor = None
    delta_b: torch.Tensor = None
    zp_b: torch.Tensor = None
    alpha_w: Optional[torch.Tensor] = None
    alpha_b: Optional[torch.Tensor] = None
    codes_w: Optional[torch.Tensor] = None  # cache of the last forward (quantizer.py:297)
    codes_b: Optional[torch.Tensor] = None


class QuantDecoder:
    """Functional stand-in for QuantModel(model, hadamard, {'channel_wise': channel_wise, 'max'}).  channel_wise=False
    (the command line without --channel_wise): one scalar step size per tensor, shape () as quantizer.py:160-168 makes it."""

    def __init__(self, stages: Sequence[Stage], bits: Sequence[int], hadamard: bool, channel_wise: bool = True):
        assert len(bits) == len(stages)
        self.channel_wise = channel_wise
        self.q: List[QStage] = []
        for st, nb in zip(stages, bits):
            assert 2 <= nb <= 8  # quantizer.py:96,237
            qs = QStage(st, nb, hadamard)
            qs.w_src = rotate_weight(st.weight) if hadamard else st.weight.clone()
            self.q.append(qs)
        self.mode = "uaq"  # 'off' | 'uaq' | 'ada'
        self.soft_w = False
        self.soft_b = False
        self.init_scales()

    @property
    def stages(self):
        return [q.stage for q in self.q]

    def avg_bits(self) -> float:
        """quant_model.py:58-72."""
        bits = sum(q.n_bits * (q.stage.weight.numel() + q.stage.bias.numel()) for q in self.q)
        return bits / sum(q.stage.weight.numel() + q.stage.bias.numel() for q in self.q)

    def init_scales(self):
        """First quantised forward: quantizer.py:112-115 on what quant_layer.py:70-74 feeds it."""
        for q in self.q:
            q.delta_w, q.zp_w = uaq_init_max(q.w_src, q.n_bits, self.channel_wise)
            q.delta_b, q.zp_b = uaq_init_max(q.stage.bias, q.n_bits, self.channel_wise)

    def start_adaround(self):
        """calib_model.py:169-184 + quantizer.py:259-319."""
        for q in self.q:
            q.delta_w, q.zp_w = fp16_round(q.delta_w), fp16_round(q.zp_w)
            q.delta_b, q.zp_b = fp16_round(q.delta_b), fp16_round(q.zp_b)
            src = q.w_src if q.hadamard else q.stage.weight  # hadamard_weight | org_weight
            q.alpha_w = adaround_init_alpha(src, q.delta_w)
            q.alpha_b = adaround_init_alpha(q.stage.bias, q.delta_b)
        self.mode, self.soft_w, self.soft_b = "ada", True, True

    def quantised_params(self, q: QStage):
        """quant_layer.py:67-77: returns the (weight, bias) the conv sees."""
        if self.mode == "off":
            return q.stage.weight, q.stage.bias
        if self.mode == "uaq":
            q.codes_w, wq = uaq_quant(q.w_src, q.delta_w, q.zp_w, q.n_bits)
            q.codes_b, bq = uaq_quant(q.stage.bias, q.delta_b, q.zp_b, q.n_bits)
        else:
            q.codes_w, wq = adaround_quant(q.w_src, q.alpha_w, q.delta_w, q.zp_w, q.n_bits, self.soft_w)
            q.codes_b, bq = adaround_quant(q.stage.bias, q.alpha_b, q.delta_b, q.zp_b, q.n_bits, self.soft_b)
        if q.hadamard:
            ci = q.stage.weight.shape[1]
            wq = hadamard_along_channel(wq)[:, :ci]
        return wq, bq

    def forward(self, embed: torch.Tensor) -> torch.Tensor:
        ws, bs = zip(*[self.quantised_params(q) for q in self.q])
        return decode(self.stages, embed, ws, bs)

    def perturbation(self):
        """quant_layer.py:86-89: org_weight - UAQ(weight) -- never rotated."""
        out = []
        for q in self.q:
            # with --hadamard the scales were fitted on the rotated tensor but are applied to the
            # plain weight here (reference quirk, quant_layer.py:70 vs :88)
            _, wq = uaq_quant(q.stage.weight, q.delta_w, q.zp_w, q.n_bits)
            out.append(q.stage.weight - wq)
        return out


def model_reconstruction(qd: QuantDecoder, cali: torch.Tensor, frames: torch.Tensor,
                         batches: Sequence[Sequence[int]], iters: int, weight: float = 0.01,
                         b_range=(20, 2), warmup: float = 0.0, p: float = 2.0, lr: float = 0.0015,
                         log: Optional[list] = None):
    """calib_model.py:92-240 with the mini-batch order injected (`batches` = one epoch of index
    lists; the reference shuffles un-seeded, calibrate_network.py:161)."""
    n_b = len(batches)
    # ---- phase 1: step sizes, Adam lr 1e-3 (calib_model.py:120-165)
    ep1 = int(0.05 * iters / n_b)
    deltas = []
    for q in qd.q:
        q.delta_w = q.delta_w.clone().requires_grad_(True)
        q.delta_b = q.delta_b.clone().requires_grad_(True)
        deltas += [q.delta_w, q.delta_b]
    opt = torch.optim.Adam(deltas, lr=0.001)
    count = 0
    for _ in range(ep1):
        for idx in batches:
            idx = torch.as_tensor(idx)
            out = qd.forward(cali[idx])
            opt.zero_grad()
            count += 1
            loss = lp_loss(out, frames[idx], p=p)
            loss.backward()
            opt.step()
            if log is not None:
                log.append(("delta", count, float(loss), 0.0, 0.0))
    for q in qd.q:
        q.delta_w, q.delta_b = q.delta_w.detach(), q.delta_b.detach()
    # ---- phase 2: rounding variables (calib_model.py:169-226)
    qd.start_adaround()
    alphas = []
    for q in qd.q:
        q.alpha_w = q.alpha_w.clone().requires_grad_(True)
        q.alpha_b = q.alpha_b.clone().requires_grad_(True)
        alphas += [q.alpha_w, q.alpha_b]
    opt = torch.optim.Adam(alphas, lr=lr)
    decay = LinearTempDecay(iters, rel_start_decay=warmup, start_b=b_range[0], end_b=b_range[1])
    loss_start = iters * warmup
    count = 0
    for _ in range(int(iters / n_b) - ep1):
        for idx in batches:
            idx = torch.as_tensor(idx)
            out = qd.forward(cali[idx])
            opt.zero_grad()
            count += 1
            rec = lp_loss(out, frames[idx], p=p)
            b = decay(count)
            if count < loss_start:
                b, rnd = 0, torch.zeros(())
            else:
                rnd = sum(weight * round_reg(q.alpha_w, b) for q in qd.q)  # bias alpha excluded
            (rec + rnd).backward()
            opt.step()
            if log is not None:
                log.append(("alpha", count, float(rec), float(rnd), float(b)))
    for q in qd.q:
        q.alpha_w, q.alpha_b = q.alpha_w.detach(), q.alpha_b.detach()
    qd.soft_w = False  # calib_model.py:231-240: only the weight quantiser goes hard
    return qd


# --------------------------------------------------------------------------------------------
# Block-wise reconstruction (calib_block.py:91-183, data_utils.py:45-86,146-196)
# --------------------------------------------------------------------------------------------
def block_cache(qd: QuantDecoder, k: int, cali: torch.Tensor, asym: bool, cache_bs: int = 10, layer: bool = False):
    """save_inp_oup_data(model, block, cali_data, asym, batch_size=10, input_prob=True) for the block that is stage k:
    (input the optimisation sees, full-precision input, full-precision output) over the calibration set; the last
    cali.size(0) % 10 samples are dropped as in data_utils.py:67.  With asym the input comes from a pass with EVERY
    layer quantised in its current state (GetLayerInpOut.__call__, data_utils.py:172-180)."""
    n = int(cali.size(0) / cache_bs) * cache_bs
    inps, syms, outs = [], [], []
    mode = qd.mode
    for i in range(0, n, cache_bs):
        e = cali[i:i + cache_bs]
        qd.mode = "off"
        with torch.no_grad():
            _, feats = decode(qd.stages, e, keep=True)
        x_fp = e if k == 0 else feats[k - 1]
        syms.append(x_fp)
        if layer:  # hook on the QuantModule itself (layer_reconstruction): the convolution's own output
            st = qd.stages[k]
            with torch.no_grad():
                outs.append(F.conv2d(x_fp, st.weight, st.bias, stride=1, padding=st.k // 2))
        else:
            outs.append(feats[k])
        if asym:
            qd.mode = mode
            with torch.no_grad():
                ws, bs = zip(*[qd.quantised_params(q) for q in qd.q])
                _, fq = decode(qd.stages, e, ws, bs, keep=True)
            inps.append(e if k == 0 else fq[k - 1])
        else:
            inps.append(x_fp)
    qd.mode = mode
    return torch.cat(inps), torch.cat(syms), torch.cat(outs)


def block_grad_cache(qd: QuantDecoder, k: int, cali: torch.Tensor, branch: str = "q", raw: bool = False, layer: bool = False):
    """save_grad_data(model, block, cali_data, batch_size=1) (data_utils.py:91-119) with GetLayerGrad (:222-258): per
    sample, loss = mean((out_fp - out_q)^2) where out_q has stages 0..k quantised in their current state
    (quantize_model_till, :261-272) and the rest full precision; the hook on the block keeps the gradient w.r.t. the
    block's output; the cache is |g| + 1.  The block is traversed by BOTH passes; the (non-full) backward hook keeps
    the gradient of the QUANTISED pass (`branch` 'q'; pinned bit-exactly against the reference's raw gradients in
    tests/golden/block_tiny_hnerv_f*.npz -- 'fp' is off by a factor of -1 and the activations' difference).  For any
    frame of more than a few thousand values |g| < 6e-8 and the cache is exactly 1.0 in fp32."""
    outs = []
    for i in range(cali.size(0)):
        e = cali[i:i + 1]
        ws, bs = [], []
        for j, q in enumerate(qd.q):
            if j < k:
                w, b = qd.quantised_params(q)  # predecessors: whatever state they are in
            elif j == k:  # the block itself: AdaRound, soft (calib_block.py:123-130 ran before save_grad_data)
                _, w = adaround_quant(q.stage.weight, q.alpha_w, q.delta_w, q.zp_w, q.n_bits, True)
                _, b = adaround_quant(q.stage.bias, q.alpha_b, q.delta_b, q.zp_b, q.n_bits, True)
            else:
                w, b = q.stage.weight, q.stage.bias
            ws.append(w.detach()); bs.append(b.detach())
        x_fp, x_q = e, e
        y_fp = y_q = None
        for j, st in enumerate(qd.stages):
            x_fp = F.conv2d(x_fp, st.weight, st.bias, stride=1, padding=st.k // 2)
            x_q = F.conv2d(x_q, ws[j], bs[j], stride=1, padding=st.k // 2)
            if j == k and layer:  # `layer`: the hooked module is the convolution alone
                x_fp, x_q = x_fp.detach().requires_grad_(True), x_q.detach().requires_grad_(True)
                y_fp, y_q = x_fp, x_q
            x_fp = apply_act(up_shuffle(x_fp, st.rh, st.rw), st.act)
            x_q = apply_act(up_shuffle(x_q, st.rh, st.rw), st.act)
            if j == k and not layer:
                x_fp = x_fp.detach().requires_grad_(True)
                x_q = x_q.detach().requires_grad_(True)
                y_fp, y_q = x_fp, x_q
        loss = F.mse_loss(x_fp, x_q, reduction="none").flatten(1).mean(1)
        g_fp, g_q = torch.autograd.grad(loss.sum(), [y_fp, y_q])
        outs.append(g_fp if branch == "fp" else g_q)
    g = torch.cat(outs)
    return g if raw else g.abs() + 1.0


def block_reconstruction(qd: QuantDecoder, k: int, cali: torch.Tensor, idx_seq: Sequence[Sequence[int]], iters: int,
                         weight: float = 0.01, asym: bool = False, b_range=(20, 2), warmup: float = 0.0,
                         input_prob: float = 1.0, p: float = 2.0, lr: float = 0.0015,
                         masks: Optional[Sequence[torch.Tensor]] = None, log: Optional[list] = None,
                         opt_mode: str = "mse", grads_out: Optional[list] = None, layer: bool = False):
    """calib_block.py:91-183 for the block that is decoder stage k (opt_mode 'mse' | 'fisher_diag' | 'fisher_full',
    calib_block.py:62-72 with the output-gradient cache of data_utils.py:91-119), with the reference's random draws
    injected: idx_seq[i] = torch.randperm(N)[:batch_size] of iteration i, masks[i] = its torch.rand_like (QDrop).
    Only this stage's quantisers become AdaRound (fp16-rounded scales, quantizer.py:264-265); both its weight and bias
    quantisers end hard-rounded (calib_block.py:180-183, unlike the network-wise variant).

    layer=True: layer_reconstruction (calib_layer.py:89-179) REPAIRED -- as shipped it stops at :130 (`opt_params +=`
    before any assignment); with `opt_params = []` it is this function on the convolution alone: the cached / compared
    output is the conv's own (before up-shuffle and activation), and the rounding regulariser is never applied, because
    LossFunction.collect_round_loss (calib_layer.py:38-46) walks the CHILDREN of the given module and a QuantModule's
    children are its two quantisers, not a QuantModule."""
    q = qd.q[k]
    if q.hadamard:
        raise NotImplementedError("the reference's block_reconstruction cannot run with hadamard=True (calib_block.py:125)")
    st = q.stage
    # AdaRoundQuantizer(uaq, weight_tensor=org_weight / bias): calib_block.py:123-130
    q.delta_w, q.zp_w = fp16_round(q.delta_w), fp16_round(q.zp_w)
    q.delta_b, q.zp_b = fp16_round(q.delta_b), fp16_round(q.zp_b)
    alpha_w = adaround_init_alpha(st.weight, q.delta_w).requires_grad_(True)
    alpha_b = adaround_init_alpha(st.bias, q.delta_b).requires_grad_(True)
    opt = torch.optim.Adam([alpha_w, alpha_b], lr=lr)
    decay = LinearTempDecay(iters, rel_start_decay=warmup, start_b=b_range[0], end_b=b_range[1])
    loss_start = iters * warmup
    # the cache is taken AFTER the block's quantisers were swapped (calib_block.py:151), but predecessors only matter
    q.alpha_w, q.alpha_b = alpha_w.detach(), alpha_b.detach()
    inp, sym, out_fp = block_cache(qd, k, cali, asym, layer=layer)
    grads = block_grad_cache(qd, k, cali, layer=layer) if opt_mode != "mse" else None
    if grads_out is not None:
        grads_out.append(grads)
    for it in range(iters):
        idx = torch.as_tensor(idx_seq[it])
        cur_inp, cur_sym, cur_out = inp[idx], sym[idx], out_fp[idx]
        if input_prob < 1.0:
            cur_inp = torch.where(masks[it] < input_prob, cur_inp, cur_sym)
        opt.zero_grad()
        _, wq = adaround_quant(st.weight, alpha_w, q.delta_w, q.zp_w, q.n_bits, True)
        _, bq = adaround_quant(st.bias, alpha_b, q.delta_b, q.zp_b, q.n_bits, True)
        y = F.conv2d(cur_inp, wq, bq, stride=1, padding=st.k // 2)
        if not layer:
            y = apply_act(up_shuffle(y, st.rh, st.rw), st.act)
        count = it + 1
        if opt_mode == "mse":
            rec = lp_loss(y, cur_out, p=p)
        elif opt_mode == "fisher_diag":  # calib_block.py:66-67
            rec = ((y - cur_out).pow(2) * grads[idx].pow(2)).sum(1).mean()
        elif opt_mode == "fisher_full":  # calib_block.py:68-72
            a, gr = (y - cur_out).abs(), grads[idx].abs()
            rec = (torch.sum(a * gr, (1, 2, 3)).view(-1, 1, 1, 1) * a * gr).mean() / 100
        else:
            raise ValueError(opt_mode)
        b = decay(count)
        if count < loss_start or layer:
            b, rnd = 0, torch.zeros(())
        else:
            rnd = weight * round_reg(alpha_w, b)
        (rec + rnd).backward()
        opt.step()
        if log is not None:
            log.append((count, float(rec + rnd), float(rec), float(rnd)))
    q.alpha_w, q.alpha_b = alpha_w.detach(), alpha_b.detach()
    return inp, sym, out_fp


# --------------------------------------------------------------------------------------------
# FP32 regression of the decoder (regress.py:239-271, utils.py:79-99,112-116): SURVEY 8(f) rank 4
# --------------------------------------------------------------------------------------------
def adjust_lr(base_lr: float, cur_epoch: float, lr_type: str, eta_min: float = 0.05) -> float:
    """utils.py:79-99."""
    if "hybrid" in lr_type:
        up_ratio, up_pow, down_pow, min_lr, final_lr = [float(x) for x in lr_type.split("_")[1:]]
        if cur_epoch < up_ratio:
            mult = min_lr + (1. - min_lr) * (cur_epoch / up_ratio) ** up_pow
        else:
            mult = 1 - (1 - final_lr) * ((cur_epoch - up_ratio) / (1. - up_ratio)) ** down_pow
    elif "cosine" in lr_type:
        up_ratio, up_pow, min_lr = [float(x) for x in lr_type.split("_")[1:]]
        if cur_epoch < up_ratio:
            mult = min_lr + (1. - min_lr) * (cur_epoch / up_ratio) ** up_pow
        else:
            mult = max(0.5 * (math.cos(math.pi * (cur_epoch - up_ratio) / (1 - up_ratio)) + 1.0), eta_min)
    else:
        raise NotImplementedError(lr_type)
    return base_lr * mult


def regress_decoder(stages: Sequence[Stage], embeds: torch.Tensor, frames: torch.Tensor, order: Sequence[Sequence[int]],
                    epochs: int, lr: float, lr_type: str = "cosine_0.1_1_0.1", log: Optional[list] = None, loss_type: str = "l2"):
    """The training loop of regress.py:249-271 for a decoder fed with FIXED embeddings (NeRV: positional encoding), loss
    'l2' (utils.py:115-116: per-frame mean over C*H*W, then the batch mean), Adam defaults, lr set per step by adjust_lr.
    order: the mini-batches of all epochs, in sequence (len(order) / epochs steps per epoch).  Updates `stages` in place."""
    ws = [s.weight.clone().requires_grad_(True) for s in stages]
    bs = [s.bias.clone().requires_grad_(True) for s in stages]
    opt = torch.optim.Adam([t for pair in zip(ws, bs) for t in pair], weight_decay=0.)
    per_epoch = len(order) // epochs
    for it, idx in enumerate(order):
        epoch, i = divmod(it, per_epoch)
        cur_lr = adjust_lr(lr, (epoch + float(i) / per_epoch) / epochs, lr_type)
        for gr in opt.param_groups:
            gr["lr"] = cur_lr
        idx = torch.as_tensor(idx)
        out = decode(stages, embeds[idx], ws, bs)
        per = F.mse_loss if loss_type == "l2" else F.l1_loss     # utils.py:115-118
        loss = per(out, frames[idx], reduction="none").flatten(1).mean(1).mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
        if log is not None:
            log.append((float(loss), cur_lr))
    for s, w, b in zip(stages, ws, bs):
        s.weight, s.bias = w.detach(), b.detach()


# --------------------------------------------------------------------------------------------
# Omega = dw^T H dw sensitivity (bit_assign.py:57-118,171-203)
# --------------------------------------------------------------------------------------------
def omega(stages: Sequence[Stage], vec: Sequence[torch.Tensor], embeds: Sequence[torch.Tensor],
          frames: Sequence[torch.Tensor]):
    """Sum over the given batches of v^T H v, H the Hessian of MSE-mean wrt the conv weights.
    Returns (omega_total, per_layer list)."""
    ws = [s.weight.clone().requires_grad_(True) for s in stages]
    bs = [s.bias for s in stages]
    hv = [torch.zeros_like(w) for w in ws]
    for e, f in zip(embeds, frames):
        out = decode(stages, e, ws, bs)
        loss = F.mse_loss(out, f)
        g = torch.autograd.grad(loss, ws, create_graph=True)
        prod = sum((gi * vi).sum() for gi, vi in zip(g, vec))
        h = torch.autograd.grad(prod, ws)
        hv = [a + b for a, b in zip(hv, h)]
    per = [float((h * v).sum()) for h, v in zip(hv, vec)]
    return sum(per), per


# --------------------------------------------------------------------------------------------
# Packed codes of the quantised artefact (include/neuroquant_b200.h: nq_pack_codes): numpy statement of the bit stream
# --------------------------------------------------------------------------------------------
def pack_codes_np(codes, n_bits: int):
    """Integer codes -> little-endian bit stream, element i in bits [i*b, (i+1)*b), padded to groups of eight."""
    import numpy as np
    c = np.asarray(codes).reshape(-1).astype(np.uint64)
    pad = (-len(c)) % 8
    c = np.concatenate([c, np.zeros(pad, dtype=np.uint64)])
    bits = ((c[:, None] >> np.arange(n_bits, dtype=np.uint64)[None, :]) & 1).astype(np.uint8).reshape(-1)
    return np.packbits(bits, bitorder="little")


def unpack_codes_np(packed, numel: int, n_bits: int):
    import numpy as np
    bits = np.unpackbits(np.asarray(packed, dtype=np.uint8), bitorder="little")[: ((numel + 7) // 8) * 8 * n_bits]
    vals = (bits.reshape(-1, n_bits).astype(np.uint32) << np.arange(n_bits, dtype=np.uint32)[None, :]).sum(1)
    return vals[:numel].astype(np.float32)
