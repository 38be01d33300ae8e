"""Randomised parity sweep of the clustering + fitting half against the oracle (GPU box): for each seed a cloud of random
size / patch count / cluster tightness, guarded mean-shift in the three precision modes, the per-segment vote and fits --
partition, bandwidth, shifted points, segment types and fitted parameters compared with oracle/oracle.py.
python tools/sweep_parity.py [first_seed] [n_seeds]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
import oracle as O
from sednet_b200 import synth
from sednet_b200.src.mean_shift import MeanShift
from sednet_b200.src import primitive_forward as PF
from util import canon, comparable_params, cylinder_fp64, rel_err, sign_align

s0 = int(sys.argv[1]) if len(sys.argv) > 1 else 0
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 20
only = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else None
dev = torch.device("cuda")
torch.set_num_threads(os.cpu_count())
bad = 0
for seed in (only or range(s0, s0 + ns)):
    rng = np.random.default_rng(seed)
    N = int(rng.integers(600, 3200))
    npatch = int(rng.integers(3, 17))
    sigma = float(rng.choice([0.005, 0.01, 0.02, 0.04]))
    iters = int(rng.choice([10, 25, 50]))
    if os.environ.get("SWEEP_N"):
        N = int(os.environ["SWEEP_N"])              # e.g. 10000: BASELINE's cloud size (the oracle then takes ~40 s per cloud)
    if os.environ.get("SWEEP_ITERS"):
        iters = int(os.environ["SWEEP_ITERS"])      # e.g. 50: the driver's setting for every cloud
    pts, nrm, lab, typ, _ = synth.make_cloud(9000 + seed, N, n_patches=npatch, min_pts=30)
    X = torch.from_numpy(synth.make_embedding(lab, 128, sigma, 100 + seed))
    with torch.no_grad():
        onew, ocen, obw, olab = O.mean_shift(X, 10000, 0.015, iters)
    ol = olab.numpy()
    msg = []
    for prec in (0, 1, 3, 4):
        np.random.seed(seed)
        newX, cen, bw, labels = MeanShift(prec_mode=prec).mean_shift(X.to(dev), 10000, 0.015, iters)
        l = labels.cpu().numpy()
        same = bool((canon(l) == canon(ol)).all())
        err = float((newX.cpu() - onew).abs().max())
        bwe = abs(float(bw) - float(obw)) / float(obw)
        if not same or err > {0: 5e-5, 1: 1e-4, 3: 2e-4, 4: 5e-5}[prec] or bwe > 1e-4:
            detail = ""
            if not same:    # how different: segment counts and the share of points that agree under the best one-to-one matching
                from scipy.optimize import linear_sum_assignment
                a, b_ = canon(l), canon(ol)
                C = np.zeros((a.max() + 1, b_.max() + 1), np.int64)
                np.add.at(C, (a, b_), 1)
                ri, ci = linear_sum_assignment(-C)
                detail = f" ({a.max() + 1} vs {b_.max() + 1} segments, {C[ri, ci].sum() / len(a):.4f} of the points agree)"
            msg.append(f"mode {prec}: partition {'same' if same else 'DIFFERENT' + detail} shifted {err:.2e} bw {bwe:.1e}")
    # fits on the oracle's segments: vote + fit through the product's dispatcher vs the oracle's
    n_seg = int(ol.max()) + 1
    st = O.segment_types(typ, ol, n_seg)
    ofits = O.fit_segments(torch.from_numpy(pts), torch.from_numpy(nrm), ol, st)
    P, Nn = torch.from_numpy(pts).to(dev), torch.from_numpy(nrm).to(dev)
    W = torch.nn.functional.one_hot(torch.from_numpy(ol), n_seg).float().to(dev)
    for s, v in ofits.items():
        name = v[0]
        m = torch.from_numpy(ol == s).to(dev)
        fit = PF.Fit()
        fn = dict(plane=fit.fit_plane_torch, sphere=fit.fit_sphere_torch, cylinder=fit.fit_cylinder_torch, cone=fit.fit_cone_torch)[name]
        out = fn(P[m], Nn[m], torch.ones(int(m.sum()), 1, device=dev))
        got = np.concatenate([np.asarray(x.detach().cpu().numpy(), np.float64).ravel() for x in out])
        want = np.concatenate([np.asarray(x.numpy() if isinstance(x, torch.Tensor) else x, np.float64).ravel() for x in v[1:]])
        g, w = comparable_params(name, got, want)
        if name == "cylinder":
            # the reference's FP32 explicit-inverse circle solve is noise-limited (tests/test_dispatch.py): axis against the
            # oracle, centre / radius against the FP64 evaluation of the same formulas
            mm = ol == s
            a64, c64, r64 = cylinder_fp64(pts[mm], nrm[mm], np.ones(int(mm.sum())))
            e = max(rel_err(g[:3], w[:3]) / 1e-5,
                    rel_err(np.concatenate([sign_align(got[:3], a64), got[3:7]]), np.concatenate([a64, c64, [r64]])) / 1e-4)
            if e > 1:
                msg.append(f"fit cylinder seg {s} ({int(m.sum())} pts): {e:.1f} x tolerance")
        elif rel_err(g, w) > (2e-4 if name == "cone" else 1e-4):
            msg.append(f"fit {name} seg {s} ({int(m.sum())} pts): rel err {rel_err(g, w):.2e}")
    print(f"seed {seed}: N {N} patches {npatch} sigma {sigma} it {iters} segments {n_seg} bw {float(obw):.4f} -> "
          + ("ok" if not msg else "; ".join(msg)), flush=True)
    bad += bool(msg)
print(f"{ns} seeds, {bad} with deviations")
