"""CPU: the stage-2 oracle (oracle/oracle_v2.py) against the vectors recorded from the unmodified reference
(tests/golden/stage2.npz, written by oracle/make_golden_v2.py), its three_nn restatement against a literal loop, and
the stage-1 -> stage-2 text files."""
import os

import numpy as np
import torch

import oracle_v2 as O2
from conftest import ROOT
from sednet_b200 import synth
from util import t


def _golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "stage2.npz"))


def test_stage2_fits_match_reference():
    g = _golden()
    n_checked = 0
    for key, ty, p, n in O2.stage2_golden_cases(synth):
        P, Nn, W = t(p), t(n), torch.ones((p.shape[0], 1))
        if ty == 1:
            for ratio in (0.5, 0.25):
                a, d = O2.fit_plane_v2(P, Nn, W, filter_ratio=ratio)
                ref = g[f"{key}_plane{int(ratio * 100)}"]
                got = np.concatenate([a.numpy().ravel(), [float(d)]])
                got = got if got[:3] @ ref[:3] > 0 else -got
                assert np.abs(got - ref).max() < 1e-6
        elif ty == 5:
            c, r = O2.fit_sphere_v2(P, Nn, W)
            assert np.abs(np.concatenate([c.numpy().ravel(), [float(r)]]) - g[f"{key}_sphere"]).max() < 1e-6
        elif ty == 4:
            a, c, r = O2.fit_cylinder_v2(P, Nn, W)
            ref = g[f"{key}_cylinder"]
            a = a.numpy().ravel()
            a = a if a @ ref[:3] > 0 else -a
            assert np.abs(np.concatenate([a, c.numpy().ravel(), [r]]) - ref).max() < 1e-6
        else:
            c, a, th = O2.fit_cone_v2(P.clone(), Nn.clone(), W.clone())
            got = np.concatenate([c.numpy().ravel(), a.numpy().ravel(), [float(th)]])
            assert np.abs(got - g[f"{key}_cone"]).max() < 2e-6
        n_checked += 1
    assert n_checked == 24


def test_circle_segmentation_known_answer():
    g = _golden()
    rng = np.random.default_rng(5)
    ang = rng.uniform(0, 2 * np.pi, 400)
    nrm = np.array([1.0, 2.0, 2.0]) / 3
    u = np.cross(nrm, [1.0, 0, 0]); u /= np.linalg.norm(u); v = np.cross(nrm, u)
    circ = (np.array([0.1, -0.2, 0.3]) + 0.37 * (np.cos(ang)[:, None] * u + np.sin(ang)[:, None] * v)).astype(np.float32)
    C, r = O2.circle_segmentation(circ)
    assert np.abs(C - g["circle_center"]).max() < 1e-9 and abs(r - float(g["circle_radius"])) < 1e-9
    assert np.abs(C - [0.1, -0.2, 0.3]).max() < 1e-6 and abs(r - 0.37) < 1e-6


def test_three_nn_restates_the_kernel_loop():
    """interpolate_gpu.cu:31-55 as a literal loop (small case), including exact duplicates (ties -> lowest index)."""
    rng = np.random.default_rng(2)
    k = rng.normal(size=(60, 3)).astype(np.float32)
    k[17] = k[3]; k[40] = k[3]                     # duplicates
    u = np.concatenate([k[:20], rng.normal(size=(10, 3)).astype(np.float32)])
    d2, idx = O2.three_nn(u, k)
    for j in range(u.shape[0]):
        best = [np.float32(np.inf)] * 3
        bi = [0, 0, 0]
        for kk in range(k.shape[0]):
            dx, dy, dz = u[j, 0] - k[kk, 0], u[j, 1] - k[kk, 1], u[j, 2] - k[kk, 2]
            d = np.float32(np.float32(dx * dx + dy * dy) + dz * dz)
            if d < best[0]:
                best = [d, best[0], best[1]]; bi = [kk, bi[0], bi[1]]
            elif d < best[1]:
                best = [best[0], d, best[1]]; bi = [bi[0], kk, bi[1]]
            elif d < best[2]:
                best[2] = d; bi[2] = kk
        assert list(idx[j]) == bi and np.array_equal(d2[j], np.array(best, np.float32))
    assert list(idx[3]) == [3, 17, 40]


def test_adjacency_maps_match_reference():
    g = _golden()
    seed, n = [int(v) for v in g["adj_cfg"]]
    pts, lab = synth.make_touching_instances(seed, n)
    ids = np.arange(int(lab.max()) + 1)
    assert np.array_equal(O2.edges_between_insts(pts, lab, True), g["edge_strict"])
    assert np.array_equal(O2.edges_between_insts(pts, lab, False), g["edge_loose"])
    mat = O2.face_face_inter_map(pts, lab, ids, 3)
    assert np.array_equal(mat, g["face_mat"]) and mat[7].sum() == 1 and g["edge_loose"].sum() > g["edge_strict"].sum() > 0


def test_stage1_text_files_round_trip(tmp_path):
    from sednet_b200.Fitting_patches_and_edges import wire
    pts, nrm, lab, typ, _ = synth.make_cloud(9, 500, n_patches=3, min_pts=100)
    edges = np.random.default_rng(0).normal(size=(2, 500)).astype(np.float32)
    paths = wire.write_stage1(str(tmp_path), 42, pts, nrm, lab, typ, edges)
    sm = torch.softmax(t(edges)[None], dim=1).transpose(1, 2).squeeze(0).numpy()
    want = O2.format_stage1(pts, nrm, lab, typ, sm)
    for p in paths:
        suffix = os.path.basename(p)[len("42"):]
        assert open(p).read() == want[suffix], suffix          # byte-identical to the reference's np.savetxt calls
    back = wire.read_stage1(str(tmp_path), 42)
    assert np.abs(back["points"] - pts).max() <= 5e-5 and np.array_equal(back["inst"], lab) and np.array_equal(back["types"], typ)
    assert back["edges"].shape == (500, 2) and np.abs(back["edges"].sum(1) - 1).max() < 2e-4


def test_metrics_oracle_matches_reference():
    """oracle/oracle_metrics.py against the values recorded from the unmodified src/segment_utils.py / src/utils.py."""
    import oracle_metrics as OM
    g = np.load(os.path.join(ROOT, "tests", "golden", "metrics.npz"))
    *seeds, n = [int(v) for v in g["cfg"]]
    for case, seed in enumerate(seeds):
        pts, gt, typ_gt, pred, typ_pred = synth.make_metric_case(seed, n)
        if case == 3:
            pts = (pts * 3.0).astype(np.float32)
        w = OM.one_hot(pred, int(np.unique(pred).shape[0]))
        for usecd in (0, 1):
            o = OM.siou_matched_segments(gt.copy(), pred.copy(), typ_pred.copy(), typ_gt.copy(), w, pts if usecd else None)
            assert np.array_equal(np.array([o[0], o[1], o[4]]), g[f"c{case}_u{usecd}"])
            assert np.array_equal(np.array(o[3]), g[f"c{case}_u{usecd}_pairs"])
        logits = np.random.default_rng(seed).normal(size=(1, n, 10)).astype(np.float32)
        logits[0, np.arange(n), typ_pred] += 3.0
        I_gt = gt.copy()
        I_gt[:50] = -1
        assert abs(float(OM.compute_type_miou_abc(logits[0], typ_gt.copy(), pred.copy(), I_gt)) - float(g[f"c{case}_abc"])) < 1e-7
        cd = OM.chamfer_distance(pts[pred == 0], pts[gt == 1])
        assert abs(cd - float(g[f"c{case}_cd"])) < 1e-7 * max(1.0, float(g[f"c{case}_cd"]))


def test_hpnet_oracle_matches_reference():
    """oracle/oracle_hpnet.py (compute_entropy, hpnet_process cache-hit branch) against the recorded reference values."""
    import oracle_hpnet as OH
    g = np.load(os.path.join(ROOT, "tests", "golden", "hpnet.npz"))
    seed, n, chunk = [int(v) for v in g["c0_cfg"]]
    feat, v, types, edges = OH.hpnet_case(seed, n)
    with torch.no_grad():
        e = [float(OH.compute_entropy(x, CHUNK=chunk)) for x in (feat, v)]
        emb = OH.hpnet_combine(feat, v, torch.tensor(float(g["c0_ent"][1])), types, edges, 0.5, chunk)
    assert np.allclose(e, g["c0_ent"], rtol=1e-6, atol=0)
    assert emb.shape == (1, n, 148) and np.abs(emb[0, ::50].numpy() - g["c0_emb_sample"]).max() < 1e-6
