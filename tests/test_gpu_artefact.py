"""The packed quantised artefact: bit-stream kernels against the numpy statement in the oracle, and
decode straight from the artefact against the live calibrated model (bit-identical frames)."""
import os

import numpy as np
import pytest
import torch

from oracle import nq_oracle as O
from tests.helpers import CASES, load, t

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("bits", [2, 3, 4, 5, 6, 7, 8])
def test_pack_unpack_kernels_match_numpy_statement(bits):
    from neuroquant_b200.artefact import pack_codes, unpack_codes
    rng = np.random.default_rng(100 + bits)
    for n in (1, 7, 8, 9, 4097, 1356800):
        c = rng.integers(0, 2 ** bits, n).astype(np.float32)
        p = pack_codes(torch.from_numpy(c).cuda(), bits)
        assert np.array_equal(p.cpu().numpy(), O.pack_codes_np(c, bits))          # bit-exact stream
        assert np.array_equal(unpack_codes(p, n, bits).cpu().numpy(), c)           # round trip
    from neuroquant_b200 import _lib as L
    with pytest.raises(L.NqError):
        pack_codes(torch.tensor([0.5, 1.0]).cuda(), bits)                          # soft (non-integer) codes are refused
    with pytest.raises(L.NqError):
        pack_codes(torch.tensor([float(2 ** bits)]).cuda(), bits)                  # out of range


@pytest.mark.parametrize("tag", list(CASES))
def test_decode_from_artefact_is_bit_identical(tag, tmp_path):
    """Calibrate briefly, write the artefact, decode from it with no weights in memory: same frames bit for bit as the
    live model, and the file is the sum of the code bits plus small per-channel tables."""
    import neuroquant_b200 as nq
    from neuroquant_b200.artefact import PackedDecoder, save_artefact
    from tests.test_gpu_kernels import dev, make_engine
    g, arch, cfg, stages, eng = make_engine(nq, tag, "uaq")
    eng.init_scales()
    cali, frames = dev(t(g["cali"])), dev(t(g["frames"]))
    order = g["order"].tolist()

    def fetch(idx):
        idx = torch.as_tensor(idx, device="cuda")
        return cali[idx], frames[idx]

    loop = nq.CalibrationLoop(eng, fetch, len(order), iters=40, weight=0.01, b_range=(20, 2), warmup=0.2, p=2.0, lr=0.003)
    loop.run(lambda: order)
    want = eng.forward(cali[:3]).clone()          # hard-rounded weights: the deliverable
    path = os.path.join(tmp_path, "model.nqb")
    size = save_artefact(eng, path)
    assert size == os.path.getsize(path)
    code_bytes = sum((s.codes_w.numel() + 7) // 8 * s.n_bits for s in eng.stages)
    tables = sum(s.delta_w.numel() * 3 + s.bias.numel() * 4 for s in eng.stages)
    assert code_bytes + tables < size < code_bytes + tables + 4096      # + JSON header
    if not bool(g["hadamard"]):  # (rotated layers store codes of the power-of-two padded channel count)
        fp32_bytes = sum(s.weight.numel() * 4 for s in eng.stages)
        assert size < fp32_bytes * (max(s.n_bits for s in eng.stages) / 32 + 0.08)
    dec = PackedDecoder(path)
    got = dec.decode(cali[:3])
    assert torch.equal(got, want)
    for s, s2 in zip(eng.stages, dec.engine.stages):
        assert torch.equal(s.codes_w, s2.codes_w) and torch.equal(s.delta_w.reshape(-1), s2.delta_w.reshape(-1))
    # soft weights are refused
    eng.soft_w = True
    eng.invalidate()
    eng.forward(cali[:2])
    from neuroquant_b200 import _lib as L
    with pytest.raises(L.NqError):
        save_artefact(eng, path)
