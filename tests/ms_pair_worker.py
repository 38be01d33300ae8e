"""Worker of test_meanshift_pair_kernel_matches_single: runs sed_ms_shift on seeded inputs under whatever SEDNET_B200_MS_PAIR
the parent set and saves the results (the switch is read once per process, hence the subprocess)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from sednet_b200 import synth  # noqa: E402
from sednet_b200.src import _lib  # noqa: E402

CASES = ((1, 256, 2, 1), (2, 1000, 3, 1), (2, 1000, 3, 3), (3, 4100, 4, 1), (10, 2100, 2, 1))   # B, N, iterations, mode


def run_cases(dev):
    res = {}
    for (B, N, iters, mode) in CASES:
        _, _, lab, _, _ = synth.make_cloud(400 + N, N, n_patches=8, min_pts=20)
        X1 = torch.from_numpy(synth.make_embedding(lab, 128, 0.02, 5)).to(dev)
        X = torch.stack([torch.roll(X1, b * 17, 0) for b in range(B)]).contiguous()
        bw = torch.full((B,), 0.3, device=dev) + 0.01 * torch.arange(B, device=dev)
        out, tmp = torch.empty_like(X), torch.empty_like(X)
        _lib.call("sed_ms_shift", _lib.ptr(X), _lib.ptr(bw), B, N, 128, iters, 0, mode, _lib.ptr(out), _lib.ptr(tmp),
                  _lib.stream())
        torch.cuda.synchronize()
        res[f"B{B}_N{N}_it{iters}_m{mode}"] = out.cpu()
    return res


if __name__ == "__main__":
    torch.save(run_cases(torch.device("cuda", 0)), sys.argv[1])
