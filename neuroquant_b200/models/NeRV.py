"""NeRV (reference: models/NeRV.py): positional-encoding input + NeRV decoder on the decoder engine."""
import time

import numpy as np
import torch
import torch.nn as nn

from ._layers import NeRVBlock, PositionEncoding
from ..runner import DecoderRunner, EmbedList


class NeRV(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.fc_h = cfg["crop_h"] // int(np.prod(cfg["dec_strides"]))
        self.fc_w = cfg["crop_w"] // int(np.prod(cfg["dec_strides"]))
        self.encoder = PositionEncoding(cfg["base"], cfg["level"])
        layers = []
        c = cfg["dec_in_channel"]
        layers.append(nn.Conv2d(int(cfg["level"] * 2), c * self.fc_h * self.fc_w, 1, 1, 0))
        for ks, stride in zip(cfg["dec_kernels"], cfg["dec_strides"]):
            co = int(max(round(c / cfg["channel_reduce"]), cfg["channel_lbound"]))
            layers.append(NeRVBlock(c, co, ks, stride, bias=True, norm=cfg["dec_norm"], act=cfg["dec_acts"]))
            c = co
        self.decoder = nn.ModuleList(layers)
        self.head_layer = nn.Conv2d(c, 3, 3, 1, 1)
        self.out_bias = cfg["out_bias"]

    def __getstate__(self):
        state = self.__dict__.copy()
        state.pop(DecoderRunner.KEY, None)  # the engine binding (device buffers, ctypes) is rebuilt on demand
        return state

    def encode(self, img):
        return self.encoder(img[:, None]).float()

    def decode(self, img_embed):
        dec_start = time.time()
        img_out = DecoderRunner.of(self).decode(img_embed)
        if torch.cuda.is_available():
            torch.cuda.synchronize()  # NeRV.py:61-62
        return img_out, EmbedList(DecoderRunner.of(self), img_embed, False), time.time() - dec_start

    def forward(self, input):
        return self.decode(self.encode(input))
