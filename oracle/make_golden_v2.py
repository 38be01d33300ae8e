"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/stage2.npz by executing the UNMODIFIED stage-2 reference modules
(/root/reference/Fitting_patches_and_edges: primitive_forward_v2.Fit, circle_fit_utils, proj_2_edge_utils) on seeded
synthetic segments, and checks oracle/oracle_v2.py against them on the way.

    python oracle/make_golden_v2.py

Compat patches on top of oracle/ref_shim.py (none touches the arithmetic of the path):
  * pointnet2._ext (a CUDA extension that cannot be built or run in this CPU-only container) is replaced by a module
    whose three_nn is oracle_v2.three_nn -- the golden adjacency maps therefore pin the reference's Python logic ON TOP
    of our restatement of that kernel (interpolate_gpu.cu:14-66), which tests/test_oracle_golden.py checks separately
    against a literal triple loop;
  * torch.lstsq (removed from torch) -> torch.linalg.lstsq, padded to the old return convention.
"""
import importlib
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)

import oracle_v2 as O2  # noqa: E402
import ref_shim  # noqa: E402

STAGE2 = os.path.join(ref_shim.REFERENCE_ROOT, "Fitting_patches_and_edges")
GOLD = os.path.join(ROOT, "tests", "golden")


def _load_synth():
    spec = importlib.util.spec_from_file_location("sednet_synth", os.path.join(ROOT, "sed-net_b200", "synth.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


synth = _load_synth()


def load_stage2():
    ref_shim.install()
    ext = types.ModuleType("pointnet2._ext")

    def three_nn(unknown, known):
        d, i = zip(*(O2.three_nn(u.numpy(), k.numpy()) for u, k in zip(unknown, known)))
        return torch.from_numpy(np.stack(d)), torch.from_numpy(np.stack(i).astype(np.int32))

    ext.three_nn = three_nn
    sys.path.insert(0, STAGE2)
    for m in ("fitting_utils", "utils", "guard", "curve_utils", "VisUtils", "approximation", "primitive_forward",
              "primitives", "PointNet", "pointnet2"):
        sys.modules.pop(m, None)      # stage 2 has its own copies of these flat module names
    pkg = importlib.import_module("pointnet2")
    sys.modules["pointnet2._ext"] = ext
    pkg._ext = ext

    def lstsq(B, A=None):
        sol = torch.linalg.lstsq(A, B).solution
        pad = torch.zeros((max(A.shape[0] - sol.shape[0], 0), sol.shape[1]), dtype=sol.dtype)
        return torch.cat([sol, pad], 0), None

    torch.lstsq = lstsq
    ns = types.SimpleNamespace()
    ns.pf = importlib.import_module("primitive_forward_v2")
    ns.circle = importlib.import_module("circle_fit_utils")
    ns.edge = importlib.import_module("proj_2_edge_utils")
    return ns


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def md(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max())


def main():
    ref = load_stage2()
    fit = ref.pf.Fit()
    out = {}
    # ---- fits on the ground-truth segments of seeded clouds (noise on the normals and points so that nothing is exact)
    cases = []
    for seed in (301, 302, 303):
        pts, nrm, lab, typ, _ = synth.make_cloud(seed, 6000, n_patches=8, normal_jitter=0.01)
        rng = np.random.default_rng(seed)
        pts = (pts + 2e-4 * rng.normal(size=pts.shape)).astype(np.float32)
        for s in range(int(lab.max()) + 1):
            m = lab == s
            cases.append((seed, s, int(typ[m][0]), pts[m], nrm[m]))
    worst = {}
    for seed, s, ty, p, n in cases:
        P, Nn, W = t(p), t(n), torch.ones((p.shape[0], 1))
        key = f"c{seed}_s{s}"
        with torch.no_grad():
            if ty == 1:
                for ratio in (0.5, 0.25):
                    a, d = fit.fit_plane_torch(P, Nn, W, filter_ratio=ratio)
                    oa, od = O2.fit_plane_v2(P, Nn, W, filter_ratio=ratio)
                    sg = 1.0 if float((a * oa).sum()) > 0 else -1.0
                    worst["plane"] = max(worst.get("plane", 0), md(a, sg * oa), md(d, sg * od))
                    out[f"{key}_plane{int(ratio * 100)}"] = np.concatenate([a.numpy().ravel(), [float(d)]])
            elif ty == 5:
                c, r = fit.fit_sphere_torch(P, Nn, W)
                oc, orr = O2.fit_sphere_v2(P, Nn, W)
                worst["sphere"] = max(worst.get("sphere", 0), md(c, oc), md(r, orr))
                out[f"{key}_sphere"] = np.concatenate([c.numpy().ravel(), [float(r)]])
            elif ty == 4:
                a, c, r = fit.fit_cylinder_torch(P, Nn, W)
                oa, oc, orr = O2.fit_cylinder_v2(P, Nn, W)
                sg = 1.0 if float((a * oa).sum()) > 0 else -1.0
                worst["cylinder"] = max(worst.get("cylinder", 0), md(a, sg * oa), md(c, oc), md(r, orr))
                out[f"{key}_cylinder"] = np.concatenate([a.numpy().ravel(), np.asarray(c).ravel(), [float(r)]])
            elif ty == 3:
                c, a, th = fit.fit_cone_torch(P.clone(), Nn.clone(), W.clone())
                oc, oa, oth = O2.fit_cone_v2(P.clone(), Nn.clone(), W.clone())
                worst["cone"] = max(worst.get("cone", 0), md(c, oc), md(a, oa), md(th, oth))
                out[f"{key}_cone"] = np.concatenate([c.numpy().ravel(), a.numpy().ravel(), [float(th)]])
    print("oracle_v2 vs reference, max abs diff per primitive:", worst)
    out["fit_cfg"] = np.array([301, 302, 303, 6000, 8])
    # ---- circle_segmentation alone (known answer: r = 0.37 circle in a tilted plane)
    rng = np.random.default_rng(5)
    ang = rng.uniform(0, 2 * np.pi, 400)
    nrm_ = np.array([1.0, 2.0, 2.0]) / 3
    u = np.cross(nrm_, [1.0, 0, 0]); u /= np.linalg.norm(u); v = np.cross(nrm_, u)
    circ = (np.array([0.1, -0.2, 0.3]) + 0.37 * (np.cos(ang)[:, None] * u + np.sin(ang)[:, None] * v)).astype(np.float32)
    _, C, r = ref.circle.circle_segmentation(circ)
    oC, orr = O2.circle_segmentation(circ)
    print("circle_segmentation: ref", C, r, "oracle diff", md(C, oC), abs(r - orr))
    out["circle_center"], out["circle_radius"] = np.asarray(C), np.float64(r)
    # ---- adjacency maps (three_nn = our restatement, Python logic = the reference's)
    pts, lab = synth.make_touching_instances(404, 3000)
    ids = np.arange(int(lab.max()) + 1)
    e_strict = ref.edge.get_edges_between_insts(t(pts), t(lab), strict=True).numpy()
    e_loose = ref.edge.get_edges_between_insts(t(pts), t(lab), strict=False).numpy()
    mat = ref.edge.face_face_inter_map(t(pts), t(lab), t(ids), nn_num_thresh=3).numpy()
    assert (e_strict == O2.edges_between_insts(pts, lab, True)).all() and (e_loose == O2.edges_between_insts(pts, lab, False)).all()
    assert (mat == O2.face_face_inter_map(pts, lab, ids, 3)).all()
    print("adjacency: edge points", int(e_strict.sum()), int(e_loose.sum()), "pairs", int(mat.sum()))
    out["adj_cfg"] = np.array([404, 3000])
    out["edge_strict"], out["edge_loose"], out["face_mat"] = e_strict, e_loose, mat
    np.savez_compressed(os.path.join(GOLD, "stage2.npz"), **out)
    print("wrote", os.path.join(GOLD, "stage2.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()
