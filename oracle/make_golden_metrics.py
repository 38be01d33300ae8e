"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/metrics.npz by executing the UNMODIFIED evaluation functions of the
reference (src/segment_utils.py, src/utils.py through oracle/ref_shim.py) on seeded label arrays, and checks
oracle/oracle_metrics.py against them.      python oracle/make_golden_metrics.py"""
import importlib
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import oracle_metrics as OM  # noqa: E402
import ref_shim  # noqa: E402


def _load_synth():
    spec = importlib.util.spec_from_file_location("sednet_synth", os.path.join(ROOT, "sed-net_b200", "synth.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


synth = _load_synth()


def main():
    ref = ref_shim.load()
    su = ref.segment_utils
    out = {}
    for case, seed in enumerate((11, 12, 13, 14)):
        pts, gt, typ_gt, pred, typ_pred = synth.make_metric_case(seed, 4000)
        if case == 3:
            pts = (pts * 3.0).astype(np.float32)      # larger cloud: some matched pairs are farther than the 0.1 chamfer bound
        n_cl = int(np.unique(pred).shape[0])
        weights = su.to_one_hot(pred, n_cl).float()
        for usecd in (0, 1):
            args = (gt.copy(), pred.copy(), typ_pred.copy(), typ_gt.copy(), weights)
            with torch.no_grad():
                res = (su.SIOU_matched_segments_usecd(*args, torch.from_numpy(pts)) if usecd else su.SIOU_matched_segments(*args))
            s_iou, p_iou, matching, pairs, recall = res
            o = OM.siou_matched_segments(gt.copy(), pred.copy(), typ_pred.copy(), typ_gt.copy(), weights.numpy(),
                                         pts if usecd else None)
            print(f"case {case} usecd {usecd}: ref {s_iou:.6f} {p_iou:.6f} {recall:.6f} | oracle diff "
                  f"{abs(s_iou - o[0]):.1e} {abs(p_iou - o[1]):.1e} {abs(recall - o[4]):.1e} pairs equal {np.array_equal(np.array(pairs), np.array(o[3]))}")
            out[f"c{case}_u{usecd}"] = np.array([s_iou, p_iou, recall], np.float64)
            out[f"c{case}_u{usecd}_pairs"] = np.array(pairs, np.int64)
            out[f"c{case}_u{usecd}_rows"], out[f"c{case}_u{usecd}_cols"] = np.asarray(matching[0][0]), np.asarray(matching[0][1])
        logits = np.random.default_rng(seed).normal(size=(1, 4000, 10)).astype(np.float32)
        logits[0, np.arange(4000), typ_pred] += 3.0
        I_gt = gt.copy()
        I_gt[:50] = -1                                   # background points (:326-331)
        r = su.compute_type_miou_abc(torch.from_numpy(logits.copy()), torch.from_numpy(typ_gt.copy())[None],
                                     torch.from_numpy(pred.copy())[None], torch.from_numpy(I_gt.copy())[None])
        o = OM.compute_type_miou_abc(logits[0].copy(), typ_gt.copy(), pred.copy(), I_gt.copy())
        print(f"case {case} type miou abc: ref {float(r):.6f} oracle {float(o):.6f}")
        out[f"c{case}_abc"] = np.float64(float(r))
        a, b = pts[pred == 0], pts[gt == 1]
        cd = ref_shim_import_utils().chamfer_distance(torch.from_numpy(a)[None], torch.from_numpy(b)[None])
        print(f"case {case} chamfer: ref {float(cd):.8f} oracle {OM.chamfer_distance(a, b):.8f}")
        out[f"c{case}_cd"] = np.float64(float(cd))
    out["cfg"] = np.array([11, 12, 13, 14, 4000])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "metrics.npz"), **out)
    print("wrote tests/golden/metrics.npz", len(out), "arrays")


def ref_shim_import_utils():
    return importlib.import_module("src.utils")


if __name__ == "__main__":
    main()
