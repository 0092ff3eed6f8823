"""HNeRV (reference: models/HNeRV.py): ConvNeXt frame encoder + NeRV decoder.  decode() -- the hot
path -- runs on the libnq_sm100 decoder engine, for the FP model and, once wrapped by QuantModel, for
the fake-quantised one."""
import time

import numpy as np
import torch
import torch.nn as nn

from ._layers import ConvNeXt, NeRVBlock
from ..runner import DecoderRunner, EmbedList


class HNeRV(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        assert cfg["enc_strides"] == cfg["dec_strides"]
        self.fc_h = int(np.prod(cfg["enc_strides"]) // np.prod(cfg["dec_strides"]))
        self.fc_w = self.fc_h
        self.encoder = ConvNeXt(stage_blocks=cfg["stage_block"], strds=cfg["enc_strides"], dims=cfg["enc_channel"],
                                drop_path_rate=0)
        layers = []
        c = cfg["dec_in_channel"]
        layers.append(nn.Conv2d(cfg["enc_channel"][-1], c, 1, 1, 0))
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
        return self.encoder(img)

    def decode(self, img_embed):
        """Returns (img_out, embed_list, dec_time) as HNeRV.py:49-71.  embed_list[0] is the input embedding; the stem's and
        the blocks' outputs (entries 1..) live in the engine's NHWC buffers and are converted on first access
        (runner.EmbedList) -- nothing on the hot path consumes them."""
        dec_start = time.time()
        img_out = DecoderRunner.of(self).decode(img_embed)
        if torch.cuda.is_available():
            torch.cuda.synchronize()  # HNeRV.py:67-68: dec_time is a host wall-clock measurement
        return img_out, EmbedList(DecoderRunner.of(self), img_embed, True), time.time() - dec_start

    def forward(self, input):
        return self.decode(self.encode(input))
