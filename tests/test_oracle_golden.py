"""CPU: the oracle (oracle/oracle.py) replayed against the golden vectors that oracle/make_golden.py recorded from
the UNMODIFIED reference (tests/golden/*.npz).  This is what pins the oracle."""
import numpy as np
import torch

import oracle as O
from sednet_b200 import synth
from util import canon, cloud_input, rel_err, sign_align, t


def test_knn_golden(golden):
    g = golden("knn")
    x = np.random.default_rng(int(g["seed_x"])).normal(size=(2, 64, 700)).astype(np.float32)
    assert (O.knn_l2(t(x), 20).numpy() == g["idx_l2"]).all()
    _, _, _, _, x6 = cloud_input(int(g["seed_cloud"]), 900)
    assert (O.knn_points_normals(t(x6), 16, 1.0).numpy() == g["idx_pn"]).all()
    gf = O.graph_feature(t(x), t(g["idx_l2"].astype(np.int64))).numpy()
    assert abs(gf.astype(np.float64).sum() - float(g["graph_feature_checksum"])) < 1e-3
    assert np.array_equal(gf[:, :, ::50, ::5], g["graph_feature_sample"])


def test_forward_golden(golden):
    g = golden("forward")
    for tag in ("plain", "gnrand"):
        seed, rgn, n, k, cseed = [int(v) for v in g[tag + "_cfg"]]
        sd = {kk: t(v) for kk, v in synth.make_state_dict(seed, randomize_gn=bool(rgn)).items()}
        _, _, _, _, x = cloud_input(cseed, n)
        with torch.no_grad():
            out, inter = O.sednet_forward(sd, t(x), k, return_intermediates=True)
        assert np.max(np.abs(out[0].numpy() - g[tag + "_emb"])) < 2e-5
        assert np.max(np.abs(out[1].numpy() - g[tag + "_logp"])) < 2e-5
        assert np.max(np.abs(out[3].numpy() - g[tag + "_edges"])) < 2e-5
        assert np.max(np.abs(inter["x4"].numpy() - g[tag + "_x4"])) < 2e-5
        assert np.max(np.abs(inter["feats"].numpy() - g[tag + "_feats"])) < 2e-5


def test_meanshift_golden(golden):
    g = golden("meanshift")
    for tag in ("a", "b"):
        seed, n, npatch, cseed = [int(v) for v in g[tag + "_cfg"]]
        _, _, lab, _, _ = synth.make_cloud(cseed, n, n_patches=npatch)
        X = t(synth.make_embedding(lab, 128, float(g[tag + "_sigma"]), seed))
        with torch.no_grad():
            newX, center, bw, labels = O.mean_shift(X, 10000, 0.015, 50)
        assert abs(float(bw) - float(g[tag + "_bw"])) < 1e-5
        assert center.shape[0] == int(g[tag + "_n_clusters"])
        assert (canon(labels.numpy()) == canon(g[tag + "_labels"])).all()
        assert np.max(np.abs(newX.numpy()[::25] - g[tag + "_newX_sample"])) < 1e-4


def test_fits_golden(golden):
    g = golden("fits")
    keys = sorted({k.rsplit("_", 1)[0] for k in g.files if k.endswith("_pts")})
    assert len(keys) == 12
    for key in keys:
        ty = int(key.split("_")[1])
        P, Nn, W = t(g[key + "_pts"]), t(g[key + "_nrm"]), t(g[key + "_w"])
        ref = g[key + "_params"]
        if ty == synth.PLANE:
            a, d = O.fit_plane(P, Nn, W)
            got = np.concatenate([a.numpy().ravel(), [float(d)]])
            got = got if got[:3] @ ref[:3] > 0 else -got
            dist = O.distance_from_plane(P, t(ref[:3]), t(ref[3:4]), sqrt=True, reduce=False)
        elif ty == synth.SPHERE:
            c, r = O.fit_sphere(P, Nn, W)
            got = np.concatenate([c.numpy().ravel(), [float(r)]])
            dist = O.distance_from_sphere(P, t(ref[:3]), t(ref[3:4]), sqrt=True, reduce=False)
        elif ty == synth.CYLINDER:
            a, c, r = O.fit_cylinder(P, Nn, W)
            a = sign_align(a.numpy(), ref[:3])
            # the centre component along the axis is noise / lambda: compare the orthogonal part
            cref, cg = ref[3:6].astype(np.float64), c.numpy().ravel().astype(np.float64)
            ax = ref[:3].astype(np.float64)
            cref, cg = cref - (cref @ ax) * ax, cg - (cg @ ax) * ax
            got = np.concatenate([a, cg, [float(r)]])
            ref = np.concatenate([ref[:3], cref, ref[6:7]])
            dist = O.distance_from_cylinder(P, t(g[key + "_params"][:3]), t(g[key + "_params"][3:6]),
                                            t(g[key + "_params"][6:7]), sqrt=True, reduce=False)
        else:
            c, a, th = O.fit_cone(P, Nn, W)
            got = np.concatenate([c.numpy().ravel(), a.numpy().ravel(), [float(th)]])
            dist = O.distance_from_cone(P, t(ref[:3]), t(ref[3:6]), t(ref[6:7]), sqrt=True, reduce=False)
        # oracle == unmodified reference bit for bit in the container that recorded the vectors (make_golden.py prints 0);
        # 1e-5 leaves room for another CPU's BLAS kernels only
        assert rel_err(got, ref) < 1e-5, (key, got, ref)
        assert np.max(np.abs(dist.numpy() - g[key + "_dist"])) < 1e-5, key


def test_misc_golden(golden):
    g = golden("misc")
    lab = np.random.default_rng(3).integers(0, 7, 300)
    assert np.array_equal(O.to_one_hot(lab, 7).numpy(), g["one_hot"])
    w = np.random.default_rng(4).uniform(-1, 1, (5, 300)).astype(np.float32)
    assert np.max(np.abs(O.weights_normalize(t(w), 0.3).numpy() - g["wn"])) < 1e-6
    A = np.random.default_rng(5).normal(size=(200, 3)).astype(np.float32)
    A2 = A.copy(); A2[:, 2] = A2[:, 0] * 0.5 - A2[:, 1]
    Y = np.random.default_rng(6).normal(size=(200, 1)).astype(np.float32)
    assert np.max(np.abs(O.lstsq(t(A), t(Y)).numpy() - g["lstsq_full"])) < 1e-5
    assert rel_err(O.lstsq(t(A2), t(Y)).numpy(), g["lstsq_def"]) < 1e-3
    assert abs(O.best_lambda(t(A2).T @ t(A2)) - float(g["best_lambda"])) < 1e-12
