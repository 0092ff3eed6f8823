"""The quantised artefact: integer weight codes packed at their bit-width + per-channel scales, and a decoder that runs
straight from it (the reference stops at fp32 "codes" inside a pickled QuantModel and leaves the
bitstream "implementation-agnostic", readme.md:125-127, quant_model.py:74-80).

File layout (little endian):  b"NQB1" | u32 header_len | JSON header | payload.  The header lists, per decoder stage:
geometry (cin, cout, k, rh, rw, act), n_bits, hadamard flag, the shape of the code tensor (C_out, C_in or its power-of-two
pad when rotated, k, k), and byte offsets of its blobs in the payload:
  codes   dense bit stream of the integer codes (nq_pack_codes: element i in bits [i*b, (i+1)*b))
  delta   fp16 per output channel -- exact: AdaRound scales are fp16-representable (quantizer.py:264-265)
  zp      u8 per output channel (an integer in 0 .. 2^b - 1)
  bias    fp32 de-quantised bias (the network-wise calibration leaves bias rounding soft, SURVEY Q3, so the bias has no
          integer code; C_out values)
Decoding from the artefact is bit-identical to the calibrated QuantModel's decode: the tensor-core stages multiply the
integer codes - zero point (one exact bf16 plane) and apply the scale in the epilogue, exactly as the live model does.
"""
from __future__ import annotations

import json
import struct
from typing import List

import numpy as np
import torch

from . import _lib as L
from .engine import DecoderEngine, QuantStage, StageGeom

MAGIC = b"NQB1"


def pack_codes(codes: torch.Tensor, n_bits: int) -> torch.Tensor:
    """fp32 integer codes (CUDA) -> uint8 bit stream (CUDA); raises if any value is not an integer in range."""
    codes = codes.contiguous()
    n = codes.numel()
    out = torch.empty(int(L.lib.nq_packed_bytes(n, n_bits)), dtype=torch.uint8, device=codes.device)
    flag = torch.zeros(1, dtype=torch.int32, device=codes.device)
    L.check(L.lib.nq_pack_codes(L.ptr(codes), n, n_bits, out.data_ptr(), flag.data_ptr(), L.stream()), "nq_pack_codes")
    if int(flag):
        raise L.NqError("pack_codes: the tensor holds non-integer or out-of-range codes (soft rounding still on?)")
    return out


def unpack_codes(packed: torch.Tensor, numel: int, n_bits: int) -> torch.Tensor:
    out = torch.empty(numel, dtype=torch.float32, device=packed.device)
    L.check(L.lib.nq_unpack_codes(packed.data_ptr(), numel, n_bits, L.ptr(out), L.stream()), "nq_unpack_codes")
    return out


def save_artefact(engine: DecoderEngine, path: str) -> int:
    """Write the artefact of a decoder whose weight quantisers are hard (AdaRound after calibration) or nearest-rounded.
    The codes are those of the engine's LAST forward (quantizer.py:297).  Returns the file size in bytes."""
    if engine.mode == "off":
        raise L.NqError("save_artefact needs a quantised decoder")
    if engine.mode == "ada" and engine.soft_w:
        raise L.NqError("save_artefact: weight rounding is still soft; finish calibration first")
    if not engine._weights_valid:
        raise L.NqError("save_artefact: run a (hard-rounded) forward first so that the codes are current")
    stages, blobs, off = [], [], 0

    def add(b: bytes):
        nonlocal off
        blobs.append(b)
        o = off
        off += len(b)
        return [o, len(b)]

    torch.cuda.synchronize()
    for s, (_, _, _, deq_w, deq_b) in zip(engine.stages, engine._packed):
        g = s.geom
        d16 = s.delta_w.reshape(-1).half()
        if not torch.equal(d16.float(), s.delta_w.reshape(-1)):
            raise L.NqError("save_artefact: step sizes are not fp16-representable (plain UAQ scales): run start_adaround() / calibrate first")
        zp = s.zp_w.reshape(-1)
        if not torch.equal(zp, zp.round()) or float(zp.min()) < 0 or float(zp.max()) > 255:
            raise L.NqError("save_artefact: zero points are not integers in 0..255")
        rec = {"cin": g.cin, "cout": g.cout, "k": g.k, "rh": g.rh, "rw": g.rw, "act": g.act, "n_bits": s.n_bits,
               "hadamard": bool(s.hadamard), "code_shape": list(s.codes_w.shape), "per_channel": s.delta_w.numel() > 1}
        rec["codes"] = add(pack_codes(s.codes_w, s.n_bits).cpu().numpy().tobytes())
        rec["delta"] = add(d16.cpu().numpy().tobytes())
        rec["zp"] = add(zp.to(torch.uint8).cpu().numpy().tobytes())
        rec["bias"] = add(deq_b.cpu().numpy().astype("<f4").tobytes())
        stages.append(rec)
    header = json.dumps({"version": 1, "stages": stages}).encode()
    with open(path, "wb") as f:
        f.write(MAGIC + struct.pack("<I", len(header)) + header + b"".join(blobs))
    return len(MAGIC) + 4 + len(header) + off


def read_artefact(path: str):
    with open(path, "rb") as f:
        raw = f.read()
    if raw[:4] != MAGIC:
        raise L.NqError(f"{path}: not a neuroquant_b200 artefact")
    (hl,) = struct.unpack("<I", raw[4:8])
    header = json.loads(raw[8:8 + hl].decode())
    if header.get("version") != 1:
        raise L.NqError(f"{path}: unsupported artefact version {header.get('version')}")
    return header, memoryview(raw)[8 + hl:]


class PackedDecoder:
    """Quantised decode straight from an artefact: no full-precision weights exist on this path."""

    def __init__(self, path: str, device="cuda"):
        header, payload = read_artefact(path)
        stages: List[QuantStage] = []
        for rec in header["stages"]:
            g = StageGeom(rec["cin"], rec["cout"], rec["k"], rec["rh"], rec["rw"], rec["act"])

            def blob(name, dtype):
                o, n = rec[name]
                return torch.from_numpy(np.frombuffer(payload[o:o + n], dtype=dtype).copy()).to(device)

            shape = tuple(rec["code_shape"])
            numel = int(np.prod(shape))
            dummy_w = torch.zeros(g.cout, g.cin, g.k, g.k, device=device)
            s = QuantStage(g, dummy_w, blob("bias", "<f4"), rec["n_bits"], rec["hadamard"])
            s.codes_w = unpack_codes(blob("codes", np.uint8), numel, rec["n_bits"]).view(shape)
            dshape = (-1, 1, 1, 1) if rec["per_channel"] else (1,)
            s.delta_w = blob("delta", "<f2").float().view(dshape).contiguous()
            s.zp_w = blob("zp", np.uint8).float().view(dshape).contiguous()
            s.w_src = s.codes_w  # shape carrier only: nothing is quantised on this path
            stages.append(s)
        self.header = header
        self.engine = DecoderEngine(stages)  # (the stages' zero `weight` tensors only carry geometry and device)
        self.engine.mode = "packed"

    def decode(self, embed: torch.Tensor) -> torch.Tensor:
        """(n, C0, h0, w0) embeddings -> (n, 3, H, W) frames."""
        return self.engine.forward(embed, reuse_weights=True).clone()

    def weight_bits(self) -> int:
        return sum(int(np.prod(r["code_shape"])) * r["n_bits"] for r in self.header["stages"])


def save_model_artefact(qnn, path: str, embed: torch.Tensor) -> int:
    """Artefact of a calibrated QuantModel: runs one quantised (hard-rounded) decode of `embed` so that the codes are
    those of the deliverable (quantizer.py:297), then writes them."""
    from .runner import DecoderRunner
    qnn.eval()
    qnn.set_quant_state(True)
    runner = DecoderRunner.of(qnn.model)
    runner.decode(embed)
    return save_artefact(runner.engine, path)
