"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/hpnet.npz by executing the UNMODIFIED src/smooth_normal_matrix.py of
the reference (compute_entropy, and hpnet_process in its cache-hit branch: the cached spectral vectors are written to a
scratch src/normal_smooth_cache/ first) and checks oracle/oracle_hpnet.py against it.   python oracle/make_golden_hpnet.py"""
import importlib
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import oracle_hpnet as OH  # noqa: E402
import ref_shim  # noqa: E402


def main():
    ref_shim.install()
    snm = importlib.import_module("src.smooth_normal_matrix")
    out = {}
    for case, (seed, n, chunk) in enumerate(((1, 2000, 400), (2, 2600, 400))):     # second case: N > 5 * CHUNK
        feat, v, types, edges = OH.hpnet_case(seed, n)
        with torch.no_grad():
            e_ref = [float(snm.compute_entropy(x, CHUNK=chunk)) for x in (feat, v)]
            e_or = [float(OH.compute_entropy(x, CHUNK=chunk)) for x in (feat, v)]
        print(f"case {case}: compute_entropy ref {e_ref} oracle {e_or}")
        ent_v = torch.tensor(e_ref[1])
        cwd = os.getcwd()
        with tempfile.TemporaryDirectory() as d:
            os.makedirs(os.path.join(d, "src", "normal_smooth_cache"))
            torch.save(v, os.path.join(d, "src", "normal_smooth_cache", "Us_7_0.1_50.pt"))
            torch.save(ent_v, os.path.join(d, "src", "normal_smooth_cache", "WUs_7_0.1_50.pt"))
            os.chdir(d)
            try:
                with torch.no_grad():
                    emb = snm.hpnet_process(feat, torch.zeros(1, n, 3), torch.zeros(1, n, 3), id=7, types=types, edges=edges,
                                            normal_smooth_w=0.5, CHUNK=chunk, gpu="cpu")
            finally:
                os.chdir(cwd)
        with torch.no_grad():
            emb_o = OH.hpnet_combine(feat, v, ent_v, types, edges, 0.5, chunk)
        print(f"case {case}: hpnet_process {tuple(emb.shape)} oracle max diff {float((emb - emb_o).abs().max()):.2e}")
        out[f"c{case}_cfg"] = np.array([seed, n, chunk])
        out[f"c{case}_ent"] = np.array(e_ref, np.float64)
        out[f"c{case}_emb_sample"] = emb[0, ::50].numpy()
        out[f"c{case}_emb_sum"] = np.float64(emb.double().sum())
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "hpnet.npz"), **out)
    print("wrote tests/golden/hpnet.npz")


if __name__ == "__main__" and "--spectral" not in sys.argv:
    main()


def spectral_case(seed, n):
    """Seeded inputs of the spectral-branch fixtures: a synthetic cloud (points, unit normals) and clustered features."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("sednet_synth", os.path.join(ROOT, "sed-net_b200", "synth.py"))
    synth = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(synth)
    p, nrm, lab, typ, _ = synth.make_cloud(seed, n)
    feat = torch.from_numpy(synth.make_embedding(lab, 128, 0.05, seed))[None] * 3.0      # "without L2": not unit rows
    return torch.from_numpy(p)[None], torch.from_numpy(nrm)[None], feat, lab


def main_spectral():
    """The branch of hpnet_process that BUILDS the spectral vectors (src/smooth_normal_matrix.py:190-196), executed
    unmodified: farthest-50 table, dense affinity matrix, torch.lobpcg(k=12, niter=10) from the torch.randn block drawn
    right after torch.manual_seed(seed) (reproducible from the seed), row normalisation, entropy, concatenation."""
    ref_shim.install()
    snm = importlib.import_module("src.smooth_normal_matrix")
    out = {}
    for case, (seed, n, chunk) in enumerate(((11, 1500, 300), (12, 2200, 400))):
        P, Nn, feat, lab = spectral_case(seed, n)
        g = torch.Generator().manual_seed(seed + 100)
        types = torch.log_softmax(2.0 * torch.randn((1, n, 6), generator=g), -1)
        edges = torch.randn((1, n, 2), generator=g)
        with torch.no_grad():
            idx_ref = snm.knn_idx(P, 50)
            A_ref = snm.construction_affinity_matrix_normal(P, Nn, sigma=0.1, knn=50)
            idx_or = OH.knn_idx(P, 50)
            A_or = OH.construction_affinity_matrix_normal(P, Nn, sigma=0.1, knn=50)
        print(f"spectral {case}: knn_idx equal {bool((idx_ref == idx_or).all())}, affinity max diff {float((A_ref - A_or).abs().max()):.2e}"
              f" (max {float(A_ref.max()):.3e})")
        cwd = os.getcwd()
        with tempfile.TemporaryDirectory() as d:
            os.makedirs(os.path.join(d, "src", "normal_smooth_cache"))
            os.chdir(d)
            try:
                with torch.no_grad():
                    torch.manual_seed(seed)
                    emb = snm.hpnet_process(feat, P, Nn, id=None, types=types, edges=edges, normal_smooth_w=0.5, CHUNK=chunk,
                                            gpu="cpu")
                v_ref = torch.load(os.path.join(d, "src", "normal_smooth_cache", "Us_None_0.1_50.pt"))
                ent_ref = float(torch.load(os.path.join(d, "src", "normal_smooth_cache", "WUs_None_0.1_50.pt")))
            finally:
                os.chdir(cwd)
        torch.manual_seed(seed)
        X0 = torch.randn((n, 12))
        with torch.no_grad():
            v_or, E_or = OH.spectral_vectors(P, Nn, X0[None])
            emb_or = OH.hpnet_combine(feat, v_or, OH.compute_entropy(v_or, chunk), types, edges, 0.5, chunk)
        print(f"spectral {case}: v max diff {float((v_ref - v_or).abs().max()):.2e}, entropy {ent_ref:.6f}, "
              f"emb max diff {float((emb - emb_or).abs().max()):.2e}, Ritz values {E_or[0, :4].tolist()}")
        out[f"s{case}_cfg"] = np.array([seed, n, chunk])
        out[f"s{case}_idx_sample"] = idx_ref[0, ::37].numpy().astype(np.int32)
        out[f"s{case}_A_rowsum"] = A_ref[0].double().sum(1).numpy()
        out[f"s{case}_A_fro"] = np.float64(A_ref[0].double().pow(2).sum().sqrt())
        out[f"s{case}_A_sample"] = A_ref[0, ::50, ::3].numpy()
        out[f"s{case}_v"] = v_ref[0].numpy()
        out[f"s{case}_ritz"] = E_or[0].numpy()
        out[f"s{case}_ent"] = np.float64(ent_ref)
        out[f"s{case}_emb_sample"] = emb[0, ::25].numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "hpnet_spectral.npz"), **out)
    print("wrote tests/golden/hpnet_spectral.npz")


if __name__ == "__main__" and "--spectral" in sys.argv:
    main_spectral()
