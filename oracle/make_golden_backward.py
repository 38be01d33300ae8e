"""TEST INFRASTRUCTURE ONLY.  Pins the backward of customsvd (reference src/fitting_utils.py:385-455: CustomSVD.backward ->
compute_grad_V -> svd_grad_K), executed UNMODIFIED from /root/reference through oracle/ref_shim.py: for seeded (m,3) inputs
and sign-invariant scalar losses of V, the gradient with respect to the input that the reference's autograd Function
returns.  Writes tests/golden/backward.npz and prints oracle-vs-reference diffs.   python oracle/make_golden_backward.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)

import oracle as O  # noqa: E402
import ref_shim  # noqa: E402


def cases():
    """(tag, input (m,3)): a generic cloud, a thin slab (plane-like: one small singular value) and one with two nearly
    equal singular values (where the 1e-6 floor of svd_grad_K matters)."""
    rng = np.random.default_rng(77)
    a = rng.normal(size=(300, 3)).astype(np.float32)
    b = (rng.normal(size=(500, 3)) * np.array([1.0, 0.7, 0.01])).astype(np.float32)
    q, _ = np.linalg.qr(rng.normal(size=(400, 3)))
    c = (q * np.array([2.0, 2.0 + 2e-7, 0.5])) @ np.linalg.qr(rng.normal(size=(3, 3)))[0]
    return (("generic", a), ("slab", b), ("near_equal", c.astype(np.float32)))


def loss_of_V(V, coef, a):
    """Sign-invariant in every column of V: sum(coef * V^2) + (a . V[:, -1])^2 (the second term is the squared cosine between
    a fixed direction and the fitted plane normal, the way fit_plane_torch consumes V)."""
    return (coef * V ** 2).sum() + (a @ V[:, -1]) ** 2


def main():
    ref = ref_shim.load()
    out = {}
    rng = np.random.default_rng(5)
    for tag, x in cases():
        coef = torch.from_numpy(rng.normal(size=(3, 3)).astype(np.float32))
        a = torch.from_numpy(rng.normal(size=3).astype(np.float32))
        X = torch.from_numpy(x).requires_grad_(True)
        U, S, V = ref.fitting_utils.customsvd(X)
        loss_of_V(V, coef, a).backward()
        g_ref = X.grad.detach().clone()
        # oracle: same forward (torch.svd), explicit backward formula
        with torch.no_grad():
            Uo, So, Vo = torch.svd(torch.from_numpy(x), some=True)
        Vl = Vo.clone().requires_grad_(True)
        loss_of_V(Vl, coef, a).backward()
        g_or = O.compute_grad_V(Uo, So, Vo, Vl.grad)
        print(f"backward[{tag}]: S {S.detach().numpy()}, |grad| max {float(g_ref.abs().max()):.3e}, oracle diff "
              f"{float((g_ref - g_or).abs().max()):.2e}")
        out[f"{tag}_x"] = x
        out[f"{tag}_coef"] = coef.numpy()
        out[f"{tag}_a"] = a.numpy()
        out[f"{tag}_grad"] = g_ref.numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "backward.npz"), **out)
    print("wrote tests/golden/backward.npz")


if __name__ == "__main__":
    main()
