"""Quantisers with the reference's names, constructor signatures and attributes
(quantization/quantizer.py), backed by the fused libnq_sm100 kernels.

Scope: the uniform-affine quantiser with the 'max' scale initialisation the
documented commands use (`--init max`) and the 'mse' / 'l1' / 'gaussian' alternatives (quantizer.py:170-222,
nq_uaq_init_search), its straight-through forward/backward, and the AdaRound 'learned_hard_sigmoid'
quantiser.  Symmetric quantisation, QATQuantizer, qfn and round_noise_ste are not on the path named by the
north star and raise NotImplementedError here.
"""
import logging
import time

import torch
import torch.nn as nn

from .. import _lib as L

ROUND_NEAREST, ROUND_SOFT, ROUND_HARD = 0, 1, 2


class StraightThrough(nn.Module):
    def __init__(self, channel_num: int = 1):
        super().__init__()

    def forward(self, input):
        return input


def round_ste(x: torch.Tensor):
    """quantizer.py:53-57 (plain tensor algebra; the fused kernels implement the same estimator)."""
    return (x.round() - x).detach() + x


class _LpLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, tgt, p, denom):
        loss, grad = L.lp_loss_sum(pred.detach().contiguous().float(), tgt.detach().contiguous().float(), p,
                                   grad_scale=1.0 / denom, want_grad=True)
        ctx.save_for_backward(grad)
        return (loss / denom).reshape(())

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None, None


def lp_loss(pred, tgt, p=2.0, reduction="none"):
    """quantizer.py:66-73: sum over channels of |pred - tgt|^p, mean over the rest ('none'), or the plain
    mean; one fused reduction kernel (nq_lp_loss) with the gradient produced in the same pass."""
    denom = pred.numel() / pred.shape[1] if reduction == "none" else pred.numel()
    return _LpLoss.apply(pred, tgt, float(p), float(denom))


class _FakeQuant(torch.autograd.Function):
    """Fused fake-quantisation with the closed-form backward of the phase's learnable
    (step size for UAQ-STE, alpha for AdaRound-soft); see include/neuroquant_b200.h."""

    @staticmethod
    def forward(ctx, x, learnable, quantizer, mode):
        xc = x.detach().contiguous().float()
        alpha = quantizer.alpha.detach() if mode != ROUND_NEAREST else None
        delta, zp = quantizer.delta.detach().contiguous(), quantizer.zero_point.detach().contiguous()
        codes, deq = L.fakequant_fwd(xc, alpha, delta, zp, quantizer.n_bits, mode)
        ctx.q, ctx.mode, ctx.xc = quantizer, mode, xc
        return deq, codes

    @staticmethod
    def backward(ctx, g, _g_codes):
        q, mode = ctx.q, ctx.mode
        if mode == ROUND_HARD:
            return None, None, None, None
        delta, zp = q.delta.detach().contiguous(), q.zero_point.detach().contiguous()
        alpha = q.alpha.detach() if mode == ROUND_SOFT else None
        d = L.fakequant_bwd(g.contiguous().float(), ctx.xc, alpha, delta, zp, q.n_bits, mode)
        # the gradient w.r.t. x is not produced: nothing in the reference's calibration consumes it
        # (the weight is not in either optimiser)
        return None, d.view_as(q.alpha if mode == ROUND_SOFT else q.delta), None, None


class UniformAffineQuantizer(nn.Module):
    """quantizer.py:76-243.  Asymmetric uniform quantisation, per output channel for 4-D weights and per
    tensor for biases when channel_wise, straight-through rounding."""

    def __init__(self, n_bits: int = 8, symmetric: bool = False, channel_wise: bool = False, scale_method: str = "max",
                 prob: float = 1.0):
        super().__init__()
        self.sym = symmetric
        assert 2 <= n_bits <= 8, "bitwidth not supported"
        self.n_bits = n_bits
        self.n_levels = 2 ** self.n_bits
        self.delta = None
        self.zero_point = None
        self.eps = torch.tensor(1e-8, dtype=torch.float32)
        self.inited = False
        self.channel_wise = channel_wise
        self.scale_method = scale_method
        self.prob = prob
        self.is_training = False
        self.x_quant = None

    def forward(self, x: torch.Tensor):
        if self.inited is False:
            self.delta, self.zero_point = self.init_quantization_scale(x, self.channel_wise)
            self.delta = nn.Parameter(self.delta)
            self.inited = True
        if self.is_training and self.prob < 1.0:
            raise NotImplementedError("QDrop (prob < 1) belongs to the block-wise variant (calib_block.py), SURVEY 8(f)")
        deq, codes = _FakeQuant.apply(x, self.delta, self, ROUND_NEAREST)
        self.x_quant = codes
        return deq

    def init_quantization_scale(self, x: torch.Tensor, channel_wise: bool = False):
        if self.sym or self.scale_method not in ("max", "mse", "l1", "gaussian"):
            raise NotImplementedError(f"scale_method={self.scale_method!r} symmetric={self.sym}: the asymmetric 'max', "
                                      "'mse', 'l1' and 'gaussian' initialisers are provided (quantizer.py:160-222)")
        if self.scale_method == "max":
            init = lambda t_, cw: L.uaq_init_max(t_, self.n_bits, cw)  # noqa: E731
        else:
            init = lambda t_, cw: L.uaq_init_search(t_, self.n_bits, cw, self.scale_method)  # noqa: E731
        if not channel_wise:
            d, z = init(x.detach().contiguous().float().view(-1), False)
            return d.view(()), z.view(())
        return init(x.detach().contiguous().float(), True)

    def bitwidth_refactor(self, refactored_bit: int):
        assert 2 <= refactored_bit <= 8, "bitwidth not supported"
        self.n_bits = refactored_bit
        self.n_levels = 2 ** self.n_bits

    def extra_repr(self):
        return f"bit={self.n_bits}, scale_method={self.scale_method}, symmetric={self.sym}, channel_wise={self.channel_wise},"


class AdaRoundQuantizer(nn.Module):
    """quantizer.py:247-323.  Learned rounding: floor(x / delta) + rectified-sigmoid(alpha) while
    soft_targets, + [alpha >= 0] afterwards."""

    def __init__(self, uaq: UniformAffineQuantizer, weight_tensor: torch.Tensor, round_mode="learned_round_sigmoid"):
        super().__init__()
        self.n_bits = uaq.n_bits
        self.sym = uaq.sym
        self.delta = uaq.delta.detach().half().float()  # quantizer.py:264-265
        self.zero_point = uaq.zero_point.detach().half().float()
        self.n_levels = uaq.n_levels
        self.round_mode = round_mode
        self.alpha = None
        self.soft_targets = False
        self.x_quant = None
        self.gamma, self.zeta = -0.1, 1.1
        self.beta = 2 / 3
        self.init_alpha(x=weight_tensor.clone())

    def forward(self, x):
        if self.round_mode in ("nearest", "nearest_ste"):
            mode = ROUND_NEAREST
        elif self.round_mode == "learned_hard_sigmoid":
            mode = ROUND_SOFT if self.soft_targets else ROUND_HARD
        elif self.round_mode == "stochastic":
            raise NotImplementedError("stochastic rounding is never selected by the reference's calibration")
        else:
            raise ValueError("Wrong rounding mode")
        deq, codes = _FakeQuant.apply(x, self.alpha if mode == ROUND_SOFT else self.delta, self, mode)
        self.x_quant = codes
        return deq

    def get_soft_targets(self):
        """quantizer.py:302-303.  Accessor only: on the calibration path the soft targets and the
        rounding regulariser are computed inside nq_fakequant_fwd / nq_fakequant_bwd."""
        return torch.clamp(torch.sigmoid(self.alpha) * (self.zeta - self.gamma) + self.gamma, 0, 1)

    def init_alpha(self, x: torch.Tensor):
        if self.round_mode != "learned_hard_sigmoid":
            raise NotImplementedError
        logging.info("Init alpha to be FP32")
        t0 = time.time()
        alpha = L.adaround_init_alpha(x.detach().contiguous().float(), self.delta.contiguous())
        self.alpha = nn.Parameter(alpha)
        self.delta = nn.Parameter(self.delta)
        logging.info("init time: {}".format(time.time() - t0))

    def extra_repr(self):
        return "bit={}".format(self.n_bits)
