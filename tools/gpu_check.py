"""Stage-by-stage diagnostic of the CUDA path against the oracle (prints differences, never asserts).
Run on the GPU box:  python tools/gpu_check.py [stage ...]"""
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]

import oracle as O  # noqa: E402
from sednet_b200 import synth  # noqa: E402
from sednet_b200.src import PointNet, SEDNet, _lib, mean_shift, primitive_forward, primitives  # noqa: E402
from util import canon, cloud_input, knn_set_agreement, rel_err, sign_align, t  # noqa: E402

dev = torch.device("cuda")


def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        r = fn()
    torch.cuda.synchronize()
    return r, (time.perf_counter() - t0) / n * 1e3


def stage_knn():
    rng = np.random.default_rng(11)
    for (B, C, N, k) in ((2, 64, 700, 20), (1, 64, 2048, 64), (2, 64, 10000, 64), (1, 64, 10000, 20)):
        x = rng.normal(size=(B, C, N)).astype(np.float32)
        ref = O.knn_l2(t(x[:1, :, :]), k).numpy() if N > 4000 else O.knn_l2(t(x), k).numpy()
        got, ms = timed(lambda: PointNet.knn(t(x).to(dev), k, k))
        got = got.cpu().numpy()[: ref.shape[0]]
        rows, shared = knn_set_agreement(got, ref)
        print(f"knn_l2 B{B} C{C} N{N} k{k}: exact {np.mean(got == ref):.6f} rows {rows:.6f} shared {shared:.6f} self-first "
              f"{(got[:, :, 0] == np.arange(N)).mean():.4f}  {ms:.2f} ms")
    for (N, k) in ((900, 16), (10000, 64)):
        _, _, _, _, x6 = cloud_input(21, N)
        ref = O.knn_points_normals(t(x6), k, 1.0).numpy()
        got, ms = timed(lambda: PointNet.knn_points_normals(t(x6).to(dev), k, k, 1.0))
        rows, shared = knn_set_agreement(got.cpu().numpy(), ref)
        print(f"knn_pn N{N} k{k}: exact {np.mean(got.cpu().numpy() == ref):.6f} rows {rows:.6f} shared {shared:.6f}  {ms:.2f} ms")


def stage_forward():
    for tag, seed, rgn, n, k in (("plain", 0, False, 600, 16), ("gnrand", 1, True, 512, 20), ("big", 2, True, 2048, 64)):
        sd_np = synth.make_state_dict(seed, randomize_gn=rgn)
        sd = {kk: t(v) for kk, v in sd_np.items()}
        _, _, _, _, x = cloud_input(100 + seed, n)
        with torch.no_grad():
            ref, inter = O.sednet_forward(sd, t(x), k, return_intermediates=True)
        m = SEDNet.SEDNet(embedding=True, emb_size=128, primitives=True, num_primitives=6, mode=5, num_channels=6,
                          combine_label_prim=True, edge_module=True, late_fusion=True, nn_nb=k)
        m.load_state_dict(sd)
        m = m.to(dev).eval()
        out, ms = timed(lambda: m(t(x).to(dev), None, False))
        x4, feats = m.encode(t(x).to(dev))
        d = lambda a, b: float(np.max(np.abs(a.cpu().numpy() - b.numpy())))
        print(f"forward[{tag}] N{n} k{k}: emb {d(out[0], ref[0]):.2e} logp {d(out[1], ref[1]):.2e} edges {d(out[3], ref[3]):.2e} "
              f"x4 {d(x4, inter['x4']):.2e} feats {d(feats, inter['feats']):.2e} "
              f"x1 {d(feats[:, :64], inter['x1']):.2e} x2 {d(feats[:, 64:128], inter['x2']):.2e} x3 {d(feats[:, 128:], inter['x3']):.2e}"
              f"  argmax-type agree {(out[1].argmax(1).cpu() == ref[1].argmax(1)).float().mean():.5f}  {ms:.2f} ms")


def stage_meanshift():
    for tag, seed, n, npatch, sigma in (("a", 5, 1500, 6, 0.01), ("b", 6, 2048, 9, 0.02), ("c", 7, 10000, 14, 0.02)):
        _, _, lab, _, _ = synth.make_cloud(200 + seed, n, n_patches=npatch)
        X = t(synth.make_embedding(lab, 128, sigma, seed))
        t0 = time.perf_counter()
        if n <= 4096:
            with torch.no_grad():
                onew, ocen, obw, olab = O.mean_shift(X, 10000, 0.015, 50)
        else:
            obw = torch.clamp(O.ms_bandwidth(X, 10000, 0.015), min=0.003)
            onew = olab = None
        tcpu = time.perf_counter() - t0
        ms = mean_shift.MeanShift(prec_mode=0)
        Xd = X.to(dev)
        (newX, center, bw, labels), tm = timed(lambda: ms.mean_shift(Xd, 10000, 0.015, 50), n=1)
        msg = f"meanshift[{tag}] N{n}: bw gpu {float(bw):.6f} oracle {float(obw):.6f} n_centers {center.shape[0]} "
        if onew is not None:
            msg += (f"newX diff {float((newX.cpu() - onew).abs().max()):.2e} labels(canon) equal "
                    f"{(canon(labels.cpu().numpy()) == canon(olab.numpy())).all()} ")
        msg += f"matches GT {(canon(labels.cpu().numpy()) == canon(lab)).all()}  gpu {tm:.1f} ms cpu {tcpu * 1e3:.0f} ms"
        print(msg)


def stage_fits():
    g = np.load(os.path.join(ROOT, "tests", "golden", "fits.npz"))
    fit = primitive_forward.Fit()
    dist = primitives.ComputePrimitiveDistance(reduce=False)
    keys = sorted({k.rsplit("_", 1)[0] for k in g.files if k.endswith("_pts")})
    for key in keys:
        ty = int(key.split("_")[1])
        P, Nn, W = t(g[key + "_pts"]).to(dev), t(g[key + "_nrm"]).to(dev), t(g[key + "_w"]).to(dev)
        ref = g[key + "_params"].astype(np.float64)
        if ty == synth.PLANE:
            a, d = fit.fit_plane_torch(P, Nn, W)
            got = np.concatenate([a.cpu().numpy().ravel(), [float(d)]])
            got = got if got[:3] @ ref[:3] > 0 else -got
            dd = dist.distance_from_plane(P, [t(g[key + "_params"][:3]).reshape(3, 1), t(g[key + "_params"][3:4])], sqrt=True)
        elif ty == synth.SPHERE:
            c, r = fit.fit_sphere_torch(P, Nn, W)
            got = np.concatenate([c.cpu().numpy().ravel(), [float(r)]])
            dd = dist.distance_from_sphere(P, [t(g[key + "_params"][:3]), t(g[key + "_params"][3:4])], sqrt=True)
        elif ty == synth.CYLINDER:
            a, c, r = fit.fit_cylinder_torch(P, Nn, W)
            a = sign_align(a.cpu().numpy(), ref[:3])
            ax = ref[:3]
            cg = c.cpu().numpy().ravel().astype(np.float64)
            cref = ref[3:6] - (ref[3:6] @ ax) * ax
            cg = cg - (cg @ ax) * ax
            got = np.concatenate([a, cg, [float(r)]])
            ref = np.concatenate([ref[:3], cref, ref[6:7]])
            q = g[key + "_params"]
            dd = dist.distance_from_cylinder(P, [t(q[:3]), t(q[3:6]), t(q[6:7])], sqrt=True)
        else:
            c, a, th = fit.fit_cone_torch(P, Nn, W)
            got = np.concatenate([c.cpu().numpy().ravel(), a.cpu().numpy().ravel(), [float(th)]])
            q = g[key + "_params"]
            dd = dist.distance_from_cone(P, [t(q[:3]), t(q[3:6]), t(q[6:7])], sqrt=True)
        print(f"fit[{key}] n={P.shape[0]}: rel err {rel_err(got, ref):.2e} dist diff {np.max(np.abs(dd.cpu().numpy() - g[key + '_dist'])):.2e}")


def stage_pipeline():
    from sednet_b200.pipeline import Pipeline, launches
    B, N, k = 2, 2048, 64
    pts, nrm, lab, typ = synth.make_batch(B, N, seed0=4321)
    sd_t, sd_i = synth.make_state_dict(0), synth.make_state_dict(1, randomize_gn=True)
    pipe = Pipeline(B, N, k)
    pipe.set_weights(sd_t, sd_i)
    P, Nn = t(pts).pin_memory(), t(nrm).pin_memory()
    launches(reset=True)
    out, ms = timed(lambda: pipe.run_host(P, Nn, 0.015, 50, 0), n=1)
    print(f"pipeline B{B} N{N}: {ms:.1f} ms, launches {launches()}, n_labels {out['n_labels'].tolist()} bw {out['bw'].tolist()}")
    t0 = time.perf_counter()
    with torch.no_grad():
        ref = O.end_to_end({kk: t(v) for kk, v in sd_t.items()}, {kk: t(v) for kk, v in sd_i.items()}, t(pts), t(nrm), k)
    print(f"oracle end_to_end {time.perf_counter() - t0:.1f} s")
    Xg = pipe.device_tensor("X").cpu()
    embg = pipe.device_tensor("embedding").cpu()
    with torch.no_grad():
        inp = torch.cat([t(pts), t(nrm)], 2).permute(0, 2, 1).contiguous()
        emb_o = O.sednet_forward({kk: t(v) for kk, v in sd_i.items()}, inp, k)[0]
    print("  embedding diff", float((embg[:B] - emb_o).abs().max()), "emb abs max", float(emb_o.abs().max()))
    for b in range(B):
        Xo = torch.nn.functional.normalize(emb_o[b].T, p=2, dim=1)
        print(f"  X diff {float((Xg[b] - Xo).abs().max()):.3e} bw oracle(Xo) {float(O.ms_bandwidth(Xo, 10000, 0.015)):.6f} "
              f"bw oracle(Xgpu) {float(O.ms_bandwidth(Xg[b], 10000, 0.015)):.6f}")
        ms = mean_shift.MeanShift()
        print(f"  bw gpu(Xo) {float(ms.compute_bandwidth(Xo.to(dev), 10000, 0.015)):.6f} gpu(Xgpu) {float(ms.compute_bandwidth(Xg[b].to(dev), 10000, 0.015)):.6f}")
    for b in range(B):
        r = ref[b]
        same = (canon(out["labels"][b].numpy()) == canon(r["labels"])).all()
        types = (out["pred_type"][b].numpy() == r["types"]).mean()
        print(f"  cloud {b}: labels(canon) equal {same} n_seg gpu {int(out['n_labels'][b])} ref {len(np.unique(r['labels']))} "
              f"pred_type agree {types:.5f} bw {float(out['bw'][b]):.6f} vs {r['bw']:.6f} fits ref {len(r['fits'])} "
              f"gpu fitted {(out['status'][b].numpy() != 1).sum()}")


def stage_tc():
    """tcgen05 mean-shift (prec 1 = split FP16, prec 2 = single FP16) against the FP32 FFMA kernel."""
    for n, npatch, sigma in ((2048, 9, 0.02), (10000, 14, 0.02), (10000, 14, 0.005), (10000, 30, 0.04)):
        _, _, lab, _, _ = synth.make_cloud(400 + n, n, n_patches=npatch, min_pts=100)
        X = t(synth.make_embedding(lab, 128, sigma, n)).to(dev)
        bw = torch.clamp(mean_shift.MeanShift(0).compute_bandwidth(X, 10000, 0.015), min=0.003) if n >= 150 else 0.3
        for iters in (1, 50):
            ref, t0 = timed(lambda: mean_shift.MeanShift(0).mean_shift_(X, bw, iters)[0], n=1)
            for prec in (1, 3, 2):
                got, tm = timed(lambda: mean_shift.MeanShift(prec).mean_shift_(X, bw, iters)[0], n=1)
                diff = (got - ref).abs().max().item()
                nrm = (torch.linalg.norm(got, dim=1) - 1).abs().max().item()
                print(f"tc N{n} iters{iters} prec{prec}: max|tc - ffma| {diff:.3e} norm err {nrm:.1e} bw {float(bw):.4f} "
                      f"nan {bool(torch.isnan(got).any())}  tc {tm:.2f} ms ffma {t0:.2f} ms")
        for prec in (1, 3, 2):
            newX, center, bw2, labels = mean_shift.MeanShift(prec).mean_shift(X, 10000, 0.015, 50)
            print(f"   prec{prec}: labels match planted partition {(canon(labels.cpu().numpy()) == canon(lab)).all()} "
                  f"n_centers {center.shape[0]}")


STAGES = dict(tc=stage_tc,knn=stage_knn, forward=stage_forward, meanshift=stage_meanshift, fits=stage_fits, pipeline=stage_pipeline)

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), "lib version", _lib.load().sed_version())
    for name in (sys.argv[1:] or list(STAGES)):
        print(f"==== {name}")
        try:
            STAGES[name]()
        except Exception:
            traceback.print_exc()
        torch.cuda.synchronize()
