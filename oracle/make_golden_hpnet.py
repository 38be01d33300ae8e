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


if __name__ == "__main__":
    main()
