"""Building blocks of the NeRV / HNeRV models, named and parameterised as in the reference
(models/_layers.py) so that its checkpoints (`state_dict` keys) load unchanged.

Only NeRVBlock, OutImg and the GELU activation are on the quantised hot path; there they run as
epilogues of the libnq_sm100 convolution kernels (neuroquant_b200/engine.py).  The modules below are
the *containers* of that path (weights, geometry) plus the never-quantised encoders (ConvNeXt,
PositionEncoding: quant_model.py:28-29 skips them), which stay stock PyTorch -- SURVEY section 8(f).
"""
from math import ceil, pi

import torch
import torch.nn as nn
import torch.nn.functional as F


def OutImg(x, out_bias="tanh"):
    """models/_layers.py:10-16."""
    if out_bias == "sigmoid":
        return torch.sigmoid(x)
    if out_bias == "tanh":
        return torch.tanh(x) * 0.5 + 0.5
    return x + float(out_bias)


class Sin(nn.Module):
    def forward(self, x):
        return torch.sin(x)


def ActivationLayer(act_type):
    """models/_layers.py:95-117 (the same table of names)."""
    table = {
        "relu": lambda: nn.ReLU(True), "leaky": lambda: nn.LeakyReLU(inplace=True),
        "leaky01": lambda: nn.LeakyReLU(negative_slope=0.1, inplace=True), "relu6": lambda: nn.ReLU6(inplace=True),
        "gelu": nn.GELU, "sin": Sin, "swish": lambda: nn.SiLU(inplace=True), "softplus": nn.Softplus,
        "hardswish": lambda: nn.Hardswish(inplace=True),
    }
    if act_type not in table:
        raise KeyError(f"Unknown activation function {act_type}.")
    return table[act_type]()


def NormLayer(norm_type, ch_width):
    """models/_layers.py:120-130."""
    if norm_type == "none":
        return nn.Identity()
    if norm_type == "batch":
        return nn.BatchNorm2d(num_features=ch_width, track_running_stats=False)
    if norm_type == "instance":
        return nn.InstanceNorm2d(num_features=ch_width)
    raise NotImplementedError


class NeRVBlock(nn.Module):
    """conv(k, stride 1, same) -> PixelShuffle(stride) -> norm -> act   (models/_layers.py:20-36)."""

    def __init__(self, in_channel, out_channel, kernel_size, stride, bias, norm, act):
        super().__init__()
        self.conv = nn.Sequential(
            nn.Conv2d(in_channel, out_channel * stride * stride, kernel_size, stride=1,
                      padding=ceil((kernel_size - 1) // 2), bias=bias),
            nn.PixelShuffle(stride) if stride != 1 else nn.Identity(),
        )
        self.norm = NormLayer(norm, out_channel)
        self.act = ActivationLayer(act)

    def forward(self, x):
        return self.act(self.norm(self.conv(x)))


class PositionEncoding(nn.Module):
    """models/_layers.py:77-85."""

    def __init__(self, base, level):
        super().__init__()
        self.pe_bases = base ** torch.arange(int(level)) * pi

    def forward(self, pos):
        value_list = pos * self.pe_bases.to(pos.device)
        pe_embed = torch.cat([torch.sin(value_list), torch.cos(value_list)], dim=-1)
        return pe_embed.view(pos.size(0), -1, 1, 1)


class LayerNorm(nn.Module):
    """channels_last / channels_first LayerNorm (models/_layers.py:235-259)."""

    def __init__(self, normalized_shape, eps=1e-6, data_format="channels_last"):
        super().__init__()
        if data_format not in ("channels_last", "channels_first"):
            raise NotImplementedError
        self.weight = nn.Parameter(torch.ones(normalized_shape))
        self.bias = nn.Parameter(torch.zeros(normalized_shape))
        self.eps = eps
        self.data_format = data_format
        self.normalized_shape = (normalized_shape,)

    def forward(self, x):
        if self.data_format == "channels_last":
            return F.layer_norm(x, self.normalized_shape, self.weight, self.bias, self.eps)
        mean = x.mean(1, keepdim=True)
        var = (x - mean).pow(2).mean(1, keepdim=True)
        x = (x - mean) / torch.sqrt(var + self.eps)
        return self.weight[:, None, None] * x + self.bias[:, None, None]


class Block(nn.Module):
    """ConvNeXt block: 7x7 depthwise conv -> LN -> 1x1 (4x) -> GELU -> 1x1 -> layer scale -> residual
    (models/_layers.py:197-232)."""

    def __init__(self, dim, drop_path=0.0, layer_scale_init_value=1e-6):
        super().__init__()
        self.dwconv = nn.Conv2d(dim, dim, kernel_size=7, padding=3, groups=dim)
        self.norm = LayerNorm(dim, eps=1e-6)
        self.pwconv1 = nn.Linear(dim, 4 * dim)
        self.act = nn.GELU()
        self.pwconv2 = nn.Linear(4 * dim, dim)
        self.gamma = nn.Parameter(layer_scale_init_value * torch.ones(dim)) if layer_scale_init_value > 0 else None
        self.drop_path = nn.Identity()  # the reference builds the encoder with drop_path_rate=0 (HNeRV.py:26)

    def forward(self, x):
        y = self.dwconv(x).permute(0, 2, 3, 1)
        y = self.pwconv2(self.act(self.pwconv1(self.norm(y))))
        if self.gamma is not None:
            y = self.gamma * y
        return x + self.drop_path(y.permute(0, 3, 1, 2))


class ConvNeXt(nn.Module):
    """The HNeRV frame encoder (models/_layers.py:134-193): per stage a strided patchify conv + LN
    and `stage_blocks` ConvNeXt blocks."""

    def __init__(self, stage_blocks=0, strds=(2, 2, 2, 2), dims=(96, 192, 384, 768), in_chans=3, drop_path_rate=0.0,
                 layer_scale_init_value=1e-6):
        super().__init__()
        self.downsample_layers = nn.ModuleList()
        self.stages = nn.ModuleList()
        self.stage_num = len(dims)
        for i in range(self.stage_num):
            if i > 0:
                layer = nn.Sequential(LayerNorm(dims[i - 1], eps=1e-6, data_format="channels_first"),
                                      nn.Conv2d(dims[i - 1], dims[i], kernel_size=strds[i], stride=strds[i]))
            else:
                layer = nn.Sequential(nn.Conv2d(in_chans, dims[0], kernel_size=strds[i], stride=strds[i]),
                                      LayerNorm(dims[0], eps=1e-6, data_format="channels_first"))
            self.downsample_layers.append(layer)
            self.stages.append(nn.Sequential(*[Block(dim=dims[i], layer_scale_init_value=layer_scale_init_value)
                                               for _ in range(stage_blocks)]))
        self.apply(self._init_weights)

    @staticmethod
    def _init_weights(m):
        if isinstance(m, (nn.Conv2d, nn.Linear)):
            nn.init.trunc_normal_(m.weight, std=0.02)
            nn.init.constant_(m.bias, 0)

    def forward(self, x):
        for i in range(self.stage_num):
            x = self.stages[i](self.downsample_layers[i](x))
        return x
