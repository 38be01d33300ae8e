"""TEST INFRASTRUCTURE ONLY.  Pins the gradient of the mean-shift iterations that the training loss differentiates
(reference src/segment_loss.py:50-56 -> src/mean_shift.py:19-79 with nms=False), executed UNMODIFIED from /root/reference
through oracle/ref_shim.py: for seeded unit-row embeddings X, bandwidth b and a linear probe W, the gradient of
sum(mean_shift_(X, b, iterations)[0] * W) with respect to X as the reference's autograd returns it.
Writes tests/golden/ms_backward.npz and prints oracle-vs-reference diffs.   python oracle/make_golden_ms_backward.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [HERE, ROOT]

import oracle as O  # noqa: E402
import ref_shim  # noqa: E402
from sednet_b200 import synth  # noqa: E402

CASES = (("a", 400, 128, 8, 0.02, 0.25, 3), ("b", 300, 64, 5, 0.03, 0.15, 5), ("c", 200, 128, 3, 0.01, 0.08, 2))


def make_case(tag, n, d, n_patches, sigma, b, iters):
    """(X (n,d) unit rows around n_patches centroids, probe W (n,d))."""
    rng = np.random.default_rng(1000 + n + d)
    lab = rng.integers(0, n_patches, n)
    X = synth.make_embedding(lab, d, sigma, 31 + n)
    W = rng.normal(size=(n, d)).astype(np.float32)
    return X.astype(np.float32), W


def main():
    ref = ref_shim.load()
    out = {}
    for tag, n, d, npatch, sigma, b, iters in CASES:
        x, w = make_case(tag, n, d, npatch, sigma, b, iters)
        W = torch.from_numpy(w)
        X = torch.from_numpy(x).requires_grad_(True)
        new_X, _ = ref.mean_shift.MeanShift().mean_shift_(X, b=torch.tensor(b), iterations=iters)
        (new_X * W).sum().backward()
        g_ref = X.grad.detach().clone()
        Xo = torch.from_numpy(x).requires_grad_(True)
        new_o = O.ms_shift(Xo, torch.tensor(b), iters)
        (new_o * W).sum().backward()
        print(f"ms_backward[{tag}]: N {n} d {d} b {b} it {iters}: |grad| max {float(g_ref.abs().max()):.3e}, "
              f"forward oracle diff {float((new_X - new_o).abs().max()):.2e}, grad oracle diff "
              f"{float((g_ref - Xo.grad).abs().max()):.2e}")
        out[f"{tag}_cfg"] = np.array([n, d, npatch, iters], np.int64)
        out[f"{tag}_sigma_b"] = np.array([sigma, b], np.float64)
        out[f"{tag}_x"] = x
        out[f"{tag}_w"] = w
        out[f"{tag}_newx"] = new_X.detach().numpy()
        out[f"{tag}_grad"] = g_ref.numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ms_backward.npz"), **out)
    print("wrote tests/golden/ms_backward.npz")


if __name__ == "__main__":
    main()
