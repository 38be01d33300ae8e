"""Row a18 of SURVEY.md section 8: the per-segment dispatcher fit_one_shape_torch (reference
src/primitive_forward.py:929-1051) + FittingModule.forward_pass_* (src/fitting_optimization.py:160-245) against
tests/golden/dispatch.npz, recorded from the UNMODIFIED reference by oracle/make_golden_dispatch.py (eval mode twice,
training mode once; a 15-point cluster and a 60-point spline cluster are dropped as the reference drops them)."""
import numpy as np
import pytest
import torch

import oracle as O
from dispatch_cases import build_case
from util import comparable_params, cylinder_fp64, flat_params, rel_err, sign_align

CASES = ("eval", "eval2", "train")


def _check(tag, g, params, residual_of, tol, res_tol, cyl_inputs=None):
    """cyl_inputs (GPU runs): ids -> (points, normals, weights) of the cylinder segments.  The reference solves the
    cylinder's circle through an FP32 explicit inverse of a cond ~ 1e6 system (src/fitting_utils.py:50-64); its own centre
    comes out quantised to multiples of 2^-7 on noisy patches (eval2, id 1: residual 1.5e-2 where the FP64 evaluation of
    the same formulas reaches 3.7e-3).  The kernel accumulates in FP64, so for cylinders: axis against the recorded
    reference, centre / radius against the FP64 evaluation of the reference's formulas (tests/util.py cylinder_fp64),
    and a residual (measured on the unweighted segment points) within 5 % of the reference's or better."""
    ids = [int(i) for i in g[tag + "_ids"]]
    names = [str(n) for n in g[tag + "_names"]]
    assert sorted(params) == ids
    for k, name in zip(ids, names):
        if name == "none":
            assert params[k] is None
            continue
        v = params[k]
        assert v[0] == name
        shapes = [tuple(x.shape) for x in v[1:]]
        assert shapes == {"plane": [(3, 1), ()], "cone": [(1, 3), (3, 1), ()], "cylinder": [(3, 1), (1, 3), ()],
                          "sphere": [(1, 3), ()]}[name], (name, shapes)                # the reference's layouts
        got, ref = comparable_params(name, flat_params(v), g[f"{tag}_{k}_params"])
        ref_res = float(g[f"{tag}_{k}_residual"])
        if name == "cylinder" and cyl_inputs is not None:
            assert rel_err(got[:3], ref[:3]) < 1e-5, (tag, k, got, ref)
            a64, c64, r64 = cylinder_fp64(*cyl_inputs[k])
            q = flat_params(v)
            assert rel_err(np.concatenate([sign_align(q[:3], a64), q[3:7]]), np.concatenate([a64, c64, [r64]])) < 1e-4
            assert residual_of(k) < ref_res * 1.05 + 1e-5, (tag, k, residual_of(k), ref_res)
            continue
        assert rel_err(got, ref) < tol, (tag, k, name, got, ref)
        assert abs(residual_of(k) - ref_res) < res_tol, (tag, k, name)


@pytest.mark.parametrize("tag", CASES)
def test_oracle_dispatch_golden(golden, tag):
    g = golden("dispatch")
    seed, mode_eval = [int(v) for v in g[tag + "_cfg"]]
    _, _, _, _, data, W = build_case(seed, bool(mode_eval), float(g[tag + "_noise"]))
    with torch.no_grad():
        params = O.fit_one_shape(data, W, eval=bool(mode_eval))
        dist = O.residual_loss({d[5][1]: d[3] for d in data}, params, sqrt=True)
    # the oracle equals the reference bit for bit where the vectors were recorded; 1e-5 for another CPU's BLAS
    _check(tag, g, params, lambda k: float(dist[k][1]), 1e-5, 1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", CASES)
def test_gpu_fit_one_shape_torch_golden(golden, tag):
    """The drop-in fit_one_shape_torch (ONE batched launch for all segments) and ResidualLoss against the reference's
    recorded fitter.fitting.parameters / residuals: 1e-4 relative (the cylinder's centre / radius carry the ~3e-3 noise
    of the reference's own FP32 explicit-inverse solve, see test_gpu_parity.py; its residual pins it)."""
    from sednet_b200.src.fitting_optimization import FittingModule
    from sednet_b200.src.primitive_forward import fit_one_shape_torch
    from sednet_b200.src.primitives import ResidualLoss
    g = golden("dispatch")
    dev = torch.device("cuda", 0)
    seed, mode_eval = [int(v) for v in g[tag + "_cfg"]]
    _, _, _, _, data, W = build_case(seed, bool(mode_eval), float(g[tag + "_noise"]))
    cyl = {}
    for d in data:                       # the rows and weights the dispatcher hands to the cylinder fit
        if d[2] == 4:
            pts, nrm, w = d[0].numpy(), d[1].numpy(), None
            if mode_eval:
                w = (W[torch.as_tensor(d[4]), d[5][0]] + O.EPS).numpy()
            else:
                pts, nrm, w = pts[::2][::2], nrm[::2][::2], (W[:, d[5][0]] + O.EPS).numpy()[::2][::2]
            cyl[d[5][1]] = (pts, nrm, w)
    data = [[d[0].to(dev), d[1].to(dev), d[2], d[3].to(dev), d[4], d[5]] for d in data]
    fitter = FittingModule("unused_closed.pth", "unused_open.pth")
    gt_points, recon = fit_one_shape_torch(data, fitter, W.to(dev), 0.1, eval=bool(mode_eval))
    params = fitter.fitting.parameters
    assert len(recon) == len(data) and all(r is None for r in recon)
    assert all((gt_points[k] is None) == (params[k] is None) for k in params)
    dist = ResidualLoss(reduce=True).residual_loss(gt_points, params, sqrt=True)
    _check(tag, g, params, lambda k: float(dist[k][1]), 1e-4, 2e-5, cyl_inputs=cyl)


@pytest.mark.gpu
def test_gpu_fitting_module_forward_pass_golden(golden):
    """FittingModule.forward_pass_{plane,cone,cylinder,sphere} one segment at a time (the reference's own call pattern),
    weights = one-hot + EPS as fit_one_shape_torch passes them (:954)."""
    from sednet_b200.src.fitting_optimization import FittingModule
    g = golden("dispatch")
    dev = torch.device("cuda", 0)
    tag = "eval2"
    seed, mode_eval = [int(v) for v in g[tag + "_cfg"]]
    _, _, _, _, data, W = build_case(seed, True, float(g[tag + "_noise"]))
    fm = FittingModule(None, None)
    call = {1: fm.forward_pass_plane, 3: fm.forward_pass_cone, 4: fm.forward_pass_cylinder, 5: fm.forward_pass_sphere}
    for pts, nrm, l, _, mask, (part, ids) in data:
        if pts.shape[0] < 20 or l not in call:
            continue
        w = (W[torch.as_tensor(mask), part:part + 1] + O.EPS).to(dev)
        assert call[l](pts.to(dev), nrm.to(dev), w, ids=ids) is None
    for k, v in fm.fitting.parameters.items():
        got, ref = comparable_params(v[0], flat_params(v), g[f"{tag}_{k}_params"])
        if v[0] == "cylinder":          # centre / radius: see _check; here only the axis against the recorded reference
            assert rel_err(got[:3], ref[:3]) < 1e-5
        else:
            assert rel_err(got, ref) < 1e-4, (k, v[0], got, ref)
    with pytest.raises(NotImplementedError):
        fm.forward_pass_plane(data[0][0].to(dev), data[0][1].to(dev), torch.ones((data[0][0].shape[0], 1), device=dev),
                              ids=0, sample_points=True)
