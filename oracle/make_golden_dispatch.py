"""TEST INFRASTRUCTURE ONLY.  Pins row a18 of SURVEY.md section 8: the per-segment dispatcher
``fit_one_shape_torch`` (src/primitive_forward.py:929-1051) driving ``FittingModule.forward_pass_*``
(src/fitting_optimization.py:160-245), executed UNMODIFIED from /root/reference through oracle/ref_shim.py on seeded
synthetic segments, with ``data`` / ``weights`` built the way Evaluation.residual_eval_mode builds them
(Fitting_patches_and_edges/residual_utils.py:210-321).  Writes tests/golden/dispatch.npz (the reference's
``fitter.fitting.parameters`` and ``ResidualLoss.residual_loss`` outputs) and prints oracle-vs-reference diffs.

    python oracle/make_golden_dispatch.py
"""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)

import oracle as O  # noqa: E402
import ref_shim  # noqa: E402
from make_golden import synth, t, maxdiff, GOLD  # noqa: E402

from dispatch_cases import build_case  # noqa: E402


def eval_weights(ref, cluster_ids, bw):
    """residual_utils.py:237-239,300-308: one-hot -> weights_normalize -> argmax one-hot, (N, C)."""
    C = np.unique(cluster_ids).shape[0]
    w0 = ref.segment_utils.to_one_hot(t(cluster_ids), C).numpy().T
    w = ref.fitting_utils.weights_normalize(t(w0.astype(np.float32)), float(bw))
    w = torch.transpose(w, 1, 0)
    return ref.segment_utils.to_one_hot(torch.max(w, 1)[1].numpy(), w.shape[1])


def flat(v):
    return np.concatenate([np.asarray(x.detach().numpy() if isinstance(x, torch.Tensor) else x, np.float64).ravel()
                           for x in v[1:]]).astype(np.float32)


def main():
    torch.set_num_threads(os.cpu_count())
    ref = ref_shim.load()
    fo = importlib.import_module("src.fitting_optimization")
    fitter = fo.FittingModule("unused_closed.pth", "unused_open.pth")
    res_loss = ref.primitives.ResidualLoss(reduce=True)
    out = {}
    for tag, seed, mode_eval, noise in (("eval", 400, True, 0.0), ("eval2", 401, True, 0.003), ("train", 402, False, 0.002)):
        p, nrm, lab, typ, data, W = build_case(seed, mode_eval, noise)
        if mode_eval:
            assert torch.equal(W, eval_weights(ref, lab, 0.1))      # the reference's own weight chain gives the same
        with torch.no_grad():
            gt_points, _ = ref.primitive_forward.fit_one_shape_torch(data, fitter, W, 0.1, eval=mode_eval)
            params_ref = dict(fitter.fitting.parameters)
            dist_ref = res_loss.residual_loss(gt_points, params_ref, sqrt=True)
            params_or = O.fit_one_shape(data, W, eval=mode_eval)
            dist_or = O.residual_loss({d[5][1]: d[3] for d in data}, params_or, sqrt=True)
        assert set(params_ref) == set(params_or), (set(params_ref), set(params_or))
        ids, names = [], []
        for k in sorted(params_ref):
            v, o = params_ref[k], params_or[k]
            ids.append(k)
            names.append("none" if v is None else v[0])
            assert (v is None) == (o is None)
            if v is None:
                continue
            assert v[0] == o[0] and [tuple(np.shape(x)) for x in v[1:]] == [tuple(np.shape(x)) for x in o[1:]], (v, o)
            print(f"dispatch[{tag}] id {k} {v[0]:9s} shapes {[tuple(np.shape(x)) for x in v[1:]]} params diff "
                  f"{maxdiff(flat(v), flat(o)):.2e} residual {float(dist_ref[k][1]):.3e} diff "
                  f"{abs(float(dist_ref[k][1]) - float(dist_or[k][1])):.2e}")
            out[f"{tag}_{k}_params"] = flat(v)
            out[f"{tag}_{k}_residual"] = np.float32(dist_ref[k][1])
        out[f"{tag}_ids"] = np.array(ids)
        out[f"{tag}_names"] = np.array(names)
        out[f"{tag}_cfg"] = np.array([seed, int(mode_eval)])
        out[f"{tag}_noise"] = np.float64(noise)
    np.savez_compressed(os.path.join(GOLD, "dispatch.npz"), **out)


if __name__ == "__main__":
    main()
