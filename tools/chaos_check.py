"""How sensitive is the calibration itself?  Runs the exact-fp32 (SIMT) engine twice on HNeRV-Bunny-3M, the second time
with the embeddings perturbed by one part in 1e7, and prints how many final integer codes differ -- the yardstick for the
tensor-core-vs-fp32 `code_mismatch` that tests/test_gpu_fullsize.py reports.  GPU only."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tests.test_gpu_fullsize import _calibrate  # noqa: E402


class _Env:
    def setenv(self, k, v):
        os.environ[k] = v


if __name__ == "__main__":
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 240
    p0, l0, c0, _ = _calibrate(_Env(), "simt", iters)
    p1, l1, c1, _ = _calibrate(_Env(), "simt", iters, perturb=1e-7)
    mism = sum(int((a != b).sum()) for a, b in zip(c0, c1)) / sum(a.numel() for a in c0)
    print({"engine": "simt vs simt(+1e-7)", "iters": iters, "code_mismatch": mism,
           "psnr": (float(p0.mean()), float(p1.mean())), "final_loss": (l0[-1][2], l1[-1][2])})
