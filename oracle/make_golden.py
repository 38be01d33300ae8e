"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz by executing the UNMODIFIED
reference (/root/reference, through oracle/ref_shim.py) on seeded synthetic inputs, and
checks the CPU restatement (oracle/oracle.py) against it on the way ("pinning" the oracle).

Run in the build container (the GPU box has no /root/reference):
    python oracle/make_golden.py            # writes tests/golden/*.npz, prints the oracle-vs-reference diffs

Inputs are NOT stored when they can be regenerated from a seed by
``sed-net_b200/synth.py`` (numpy Generator streams are stable); only reference OUTPUTS are.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)

import oracle as O  # noqa: E402
import ref_shim  # noqa: E402


def _load_synth():
    spec = importlib.util.spec_from_file_location("sednet_synth", os.path.join(ROOT, "sed-net_b200", "synth.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


synth = _load_synth()
GOLD = os.path.join(ROOT, "tests", "golden")


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def maxdiff(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b))) if a.size else 0.0


def golden_knn(ref):
    rng = np.random.default_rng(11)
    x = rng.normal(size=(2, 64, 700)).astype(np.float32)
    idx_ref = ref.PointNet.knn(t(x), 20, 20).numpy()
    idx_or = O.knn_l2(t(x), 20).numpy()
    p, n, _, _, _ = synth.make_cloud(21, 900)
    x6 = np.concatenate([p, n], 1).T[None].copy()
    idx6_ref = ref.PointNet.knn_points_normals(t(x6), 16, 16, 1.0).numpy()
    idx6_or = O.knn_points_normals(t(x6), 16, 1.0).numpy()
    gf_ref = ref.PointNet.get_graph_feature(t(x), 20, 20).numpy()
    gf_or = O.graph_feature(t(x), t(idx_or)).numpy()
    print("knn: idx equal", (idx_ref == idx_or).all(), (idx6_ref == idx6_or).all(), "graph feature diff", maxdiff(gf_ref, gf_or))
    np.savez_compressed(os.path.join(GOLD, "knn.npz"), seed_x=11, idx_l2=idx_ref.astype(np.int32),
                        seed_cloud=21, idx_pn=idx6_ref.astype(np.int32),
                        graph_feature_checksum=np.float64(gf_ref.astype(np.float64).sum()),
                        graph_feature_sample=gf_ref[:, :, ::50, ::5])


def _ref_model(ref, sd_np, k):
    m = ref.SEDNet.SEDNet(embedding=True, emb_size=128, primitives=True, num_primitives=6, mode=5, num_channels=6,
                          combine_label_prim=True, edge_module=True, late_fusion=True, nn_nb=k)
    m.load_state_dict({kk: t(v) for kk, v in sd_np.items()})
    return m.eval()


def golden_forward(ref):
    out = {}
    for tag, seed, rgn, n, k in (("plain", 0, False, 600, 16), ("gnrand", 1, True, 512, 20)):
        sd_np = synth.make_state_dict(seed, randomize_gn=rgn)
        p, nrm, _, _, _ = synth.make_cloud(100 + seed, n)
        x = t(np.concatenate([p, nrm], 1).T[None].copy())
        with torch.no_grad():
            r = _ref_model(ref, sd_np, k)(x, None, False)
            o, inter = O.sednet_forward({kk: t(v) for kk, v in sd_np.items()}, x, k, return_intermediates=True)
            x4_ref, feats_ref = _ref_model(ref, sd_np, k).encoder(x)
        print(f"forward[{tag}]: emb {maxdiff(r[0], o[0]):.2e} logp {maxdiff(r[1], o[1]):.2e} edges {maxdiff(r[3], o[3]):.2e}"
              f" x4 {maxdiff(x4_ref, inter['x4']):.2e} feats {maxdiff(feats_ref, inter['feats']):.2e}")
        out.update({f"{tag}_emb": r[0].numpy(), f"{tag}_logp": r[1].numpy(), f"{tag}_edges": r[3].numpy(),
                    f"{tag}_x4": x4_ref.numpy(), f"{tag}_feats": feats_ref.numpy(),
                    f"{tag}_cfg": np.array([seed, int(rgn), n, k, 100 + seed])})
    np.savez_compressed(os.path.join(GOLD, "forward.npz"), **out)


def golden_meanshift(ref):
    out = {}
    ms = ref.mean_shift.MeanShift()
    for tag, seed, n, npatch, sigma in (("a", 5, 1500, 6, 0.01), ("b", 6, 2048, 9, 0.02)):
        _, _, lab, _, _ = synth.make_cloud(200 + seed, n, n_patches=npatch)
        X = t(synth.make_embedding(lab, 128, sigma, seed))
        np.random.seed(0)
        with torch.no_grad():
            newX, center, bw, labels = ms.mean_shift(X, 10000, 0.015, 50)
        onew, ocen, obw, olab = O.mean_shift(X, 10000, 0.015, 50)
        same = (O.canonical_labels(labels.numpy()) == O.canonical_labels(olab.numpy())).all()
        pure = (O.canonical_labels(labels.numpy()) == O.canonical_labels(lab)).all()
        print(f"meanshift[{tag}]: bw ref {float(bw):.6f} oracle {float(obw):.6f} newX diff {maxdiff(newX, onew):.2e} "
              f"labels(canon) equal {same} raw equal {(labels.numpy() == olab.numpy()).all()} matches GT {pure} n_clusters {center.shape[0]}")
        out.update({f"{tag}_cfg": np.array([seed, n, npatch, 200 + seed]), f"{tag}_sigma": np.float64(sigma),
                    f"{tag}_bw": np.float32(bw), f"{tag}_labels": labels.numpy().astype(np.int32),
                    f"{tag}_newX_sample": newX.numpy()[::25], f"{tag}_n_clusters": np.int64(center.shape[0])})
    np.savez_compressed(os.path.join(GOLD, "meanshift.npz"), **out)


def _segments(seed, n, noise=0.0):
    """One segment per analytic type, from a synthetic cloud; returns {type: (pts, normals)}."""
    segs = {}
    s = seed
    while len(segs) < 4:
        p, nrm, lab, typ, _ = synth.make_cloud(s, n, n_patches=8)
        for l in np.unique(lab):
            m = lab == l
            ty = int(typ[m][0])
            if ty not in segs and m.sum() >= 100:
                pp = p[m] + (noise * np.random.default_rng(s).normal(size=p[m].shape)).astype(np.float32)
                segs[ty] = (pp.astype(np.float32), nrm[m])
        s += 1
    return segs


def golden_fits(ref):
    fit = ref.primitive_forward.Fit()
    dist = ref.primitives.ComputePrimitiveDistance(reduce=False)
    out = {}
    for tag, seed, noise, soft in (("clean", 300, 0.0, False), ("noisy", 310, 0.002, False), ("soft", 320, 0.001, True)):
        segs = _segments(seed, 4000, noise)
        for ty, (p, nrm) in segs.items():
            P, Nn = t(p), t(nrm)
            if soft:
                w = np.random.default_rng(seed + ty).uniform(0.05, 1.0, (p.shape[0], 1)).astype(np.float32)
            else:
                w = np.ones((p.shape[0], 1), np.float32)
            W = t(w)
            key = f"{tag}_{ty}"
            out[key + "_pts"], out[key + "_nrm"], out[key + "_w"] = p, nrm, w
            if ty == synth.PLANE:
                a, d = fit.fit_plane_torch(P, Nn, W)
                oa, od = O.fit_plane(P, Nn, W)
                sgn = 1.0 if float((a * oa).sum()) > 0 else -1.0
                print(f"fit[{key}] plane: a diff {maxdiff(a, sgn * oa):.2e} d diff {abs(float(d) - sgn * float(od)):.2e}")
                out[key + "_params"] = np.concatenate([a.numpy().ravel(), [float(d)]]).astype(np.float32)
                out[key + "_dist"] = dist.distance_from_plane(P, [a.reshape(3, 1), d], sqrt=True).numpy()
                print("    dist diff", maxdiff(out[key + "_dist"], O.distance_from_plane(P, a, d, sqrt=True, reduce=False)))
            elif ty == synth.SPHERE:
                c, r = fit.fit_sphere_torch(P, Nn, W)
                oc, orr = O.fit_sphere(P, Nn, W)
                print(f"fit[{key}] sphere: c diff {maxdiff(c, oc):.2e} r diff {abs(float(r) - float(orr)):.2e}")
                out[key + "_params"] = np.concatenate([c.numpy().ravel(), [float(r)]]).astype(np.float32)
                out[key + "_dist"] = dist.distance_from_sphere(P, [c, r], sqrt=True).numpy()
                print("    dist diff", maxdiff(out[key + "_dist"], O.distance_from_sphere(P, c, r, sqrt=True, reduce=False)))
            elif ty == synth.CYLINDER:
                a, c, r = fit.fit_cylinder_torch(P, Nn, W)
                oa, oc, orr = O.fit_cylinder(P, Nn, W)
                sgn = 1.0 if float((a * oa).sum()) > 0 else -1.0
                print(f"fit[{key}] cylinder: a diff {maxdiff(a, sgn * oa):.2e} c diff {maxdiff(c, oc):.2e} r diff {abs(float(r) - float(orr)):.2e}")
                out[key + "_params"] = np.concatenate([a.numpy().ravel(), c.numpy().ravel(), [float(r)]]).astype(np.float32)
                out[key + "_dist"] = dist.distance_from_cylinder(P, [a, c, r], sqrt=True).numpy()
                print("    dist diff", maxdiff(out[key + "_dist"], O.distance_from_cylinder(P, a, c, r, sqrt=True, reduce=False)))
            elif ty == synth.CONE:
                c, a, th = fit.fit_cone_torch(P, Nn, W)
                oc, oa, oth = O.fit_cone(P, Nn, W)
                print(f"fit[{key}] cone: apex diff {maxdiff(c, oc):.2e} a diff {maxdiff(a, oa):.2e} theta diff {abs(float(th) - float(oth)):.2e}")
                out[key + "_params"] = np.concatenate([c.numpy().ravel(), a.numpy().ravel(), [float(th)]]).astype(np.float32)
                out[key + "_dist"] = dist.distance_from_cone(P, [c, a, th], sqrt=True).numpy()
                print("    dist diff", maxdiff(out[key + "_dist"], O.distance_from_cone(P, c, a, th, sqrt=True, reduce=False)))
    # analytic known-answer inputs of the reference's own (assert-free) smoke tests,
    # Fitting_patches_and_edges/test_fitting_utils.py:12-13,28,47: exact surfaces, known parameters
    np.savez_compressed(os.path.join(GOLD, "fits.npz"), **out)


def golden_misc(ref):
    lab = np.random.default_rng(3).integers(0, 7, 300)
    oh_ref = ref.segment_utils.to_one_hot(t(lab), 7).numpy()
    oh_or = O.to_one_hot(lab, 7).numpy()
    w = np.random.default_rng(4).uniform(-1, 1, (5, 300)).astype(np.float32)
    wn_ref = ref.fitting_utils.weights_normalize(t(w), 0.3).numpy()
    wn_or = O.weights_normalize(t(w), 0.3).numpy()
    A = np.random.default_rng(5).normal(size=(200, 3)).astype(np.float32)
    A2 = A.copy(); A2[:, 2] = A2[:, 0] * 0.5 - A2[:, 1]  # rank 2
    Y = np.random.default_rng(6).normal(size=(200, 1)).astype(np.float32)
    ls = ref.fitting_utils.LeastSquares()
    x_full, x_def = ls.lstsq(t(A), t(Y), 0.0).numpy(), ls.lstsq(t(A2), t(Y), 0.0).numpy()
    lam = ref.fitting_utils.best_lambda(t(A2).T @ t(A2))
    print("misc: one_hot eq", (oh_ref == oh_or).all(), "weights_normalize diff", maxdiff(wn_ref, wn_or),
          "lstsq full diff", maxdiff(x_full, O.lstsq(t(A), t(Y))), "deficient diff", maxdiff(x_def, O.lstsq(t(A2), t(Y))),
          "lambda", lam, O.best_lambda(t(A2).T @ t(A2)))
    np.savez_compressed(os.path.join(GOLD, "misc.npz"), one_hot=oh_ref, wn=wn_ref, lstsq_full=x_full, lstsq_def=x_def,
                        best_lambda=np.float64(lam))


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    ref = ref_shim.load()
    golden_knn(ref)
    golden_misc(ref)
    golden_fits(ref)
    golden_meanshift(ref)
    golden_forward(ref)


if __name__ == "__main__":
    main()
