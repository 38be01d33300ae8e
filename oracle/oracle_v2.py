"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import anything under oracle/).

CPU restatement of the reference's STAGE-2 primitive fits and instance-adjacency maps:
    Fit.fit_plane_torch / fit_sphere_torch / fit_cylinder_torch / fit_cone_torch
        Fitting_patches_and_edges/primitive_forward_v2.py:716-891
    circle_segmentation, fit_circle_2d, rodrigues_rot
        Fitting_patches_and_edges/circle_fit_utils.py:11-113
    three_nn   Fitting_patches_and_edges/pointnet2/_ext_src/src/interpolate_gpu.cu:14-66 (a CUDA kernel: restated)
    get_edges_between_insts, face_face_inter_map
        Fitting_patches_and_edges/proj_2_edge_utils.py:45-110
    stage-1 -> stage-2 text files   generate_predictions_aug.py:424-437

Pinned by oracle/make_golden_v2.py against the unmodified reference modules executed in the build container (with
pointnet2._ext.three_nn -- an unbuildable-here CUDA extension -- replaced by three_nn() below, and the removed
torch.lstsq mapped to torch.linalg.lstsq); outputs recorded in tests/golden/stage2.npz.
"""
import numpy as np
import torch

import oracle as O

EPS = float(np.finfo(np.float32).eps)


# ------------------------------------------------------------------------------------------------ three_nn
def three_nn(unknown, known):
    """interpolate_gpu.cu:14-66: for every row of unknown (n,3) the three nearest rows of known (m,3) by the FP32
    direct-form squared distance, strict '<' updates in index order (ties -> lowest index).  Returns
    (dist2 (n,3) float32, idx (n,3) int64)."""
    u = np.asarray(unknown, np.float32)
    k = np.asarray(known, np.float32)
    n = u.shape[0]
    dist2 = np.empty((n, 3), np.float32)
    idx = np.empty((n, 3), np.int64)
    for j0 in range(0, n, 512):
        uu = u[j0:j0 + 512]
        dx = uu[:, None, 0] - k[None, :, 0]
        dy = uu[:, None, 1] - k[None, :, 1]
        dz = uu[:, None, 2] - k[None, :, 2]
        d = (dx * dx + dy * dy) + dz * dz
        order = np.argsort(d, axis=1, kind="stable")[:, :3]     # stable: ties -> lowest index, as the '<' chain
        idx[j0:j0 + 512] = order
        dist2[j0:j0 + 512] = np.take_along_axis(d, order, 1)
    return dist2, idx


def edges_between_insts(points, insts, strict=True):
    """proj_2_edge_utils.py:45-60: points whose first (and, strict, second) non-self neighbour belongs to another
    instance."""
    p = np.asarray(points, np.float32)[:, :3]
    insts = np.asarray(insts)
    idx = three_nn(p, p)[1]
    one = insts[idx[:, 1]] != insts
    two = insts[idx[:, 2]] != insts
    return (one & two) if strict else one


def face_face_inter_map(points, insts, primitive_ids, nn_num_thresh=3):
    """proj_2_edge_utils.py:63-110: 30 x 30 boolean adjacency of instances (row = instance id)."""
    p = np.asarray(points, np.float32)
    insts = np.asarray(insts).astype(np.int64)
    idx = three_nn(p, p)[1]
    one, two = idx[:, 1], idx[:, 2]
    mat = np.zeros((30, 30), bool)
    for _id in primitive_ids:
        _id = int(_id)
        sel = insts == _id
        a = insts[one[sel]]
        b = insts[two[sel]]
        diff = np.concatenate([a[a != _id], b[b != _id]])
        vals, counts = np.unique(diff, return_counts=True)
        for v, c in zip(vals, counts):
            if c >= nn_num_thresh:
                mat[_id, int(v)] = True
    ids = [int(v) for v in primitive_ids]
    for i in range(30):
        if not mat[i].any() and i in ids:
            sel = np.flatnonzero(insts == i)
            d = ((p - p[sel[0]][None]) ** 2).sum(-1)
            order = np.argsort(d, kind="stable")
            inst_sort = insts[order]
            mat[i, int(inst_sort[inst_sort != i][0])] = True
    return mat


# ------------------------------------------------------------------------------------------------ fits
def _crop_nearest(points, m):
    """index = argsort(|p - mean|^2)[:m] (primitive_forward_v2.py:723-725, 831-834, 854-856)."""
    center = points.mean(0).reshape((1, 3))
    return torch.argsort((points - center).pow(2).sum(-1), stable=True)[:m]


def fit_plane_v2(points, normals, weights, nofilter=False, filter_ratio=0.5):
    """primitive_forward_v2.py:716-752."""
    if not nofilter:
        index = _crop_nearest(points, int(points.shape[0] * filter_ratio))
        points, weights = points[index], weights[index]
    return O.fit_plane(points, None, weights)


def fit_sphere_v2(points, normals, weights):
    """primitive_forward_v2.py:772-796: the same algebra as stage 1 (src/primitive_forward.py:750-773)."""
    return O.fit_sphere(points, normals, weights)


def rodrigues_rot(P, n0, n1):
    """circle_fit_utils.py:11-28 (FP64)."""
    P = np.atleast_2d(np.asarray(P, np.float64))
    n0 = np.asarray(n0, np.float64) / np.linalg.norm(n0)
    n1 = np.asarray(n1, np.float64) / np.linalg.norm(n1)
    k = np.cross(n0, n1)
    k = k / np.linalg.norm(k)
    theta = np.arccos(np.dot(n0, n1))
    return P * np.cos(theta) + np.cross(k[None], P) * np.sin(theta) + k[None] * (P @ k)[:, None] * (1 - np.cos(theta))


def circle_segmentation(cloud):
    """circle_fit_utils.py:80-113: plane of the (already planar) projected points by SVD, rotation to z, algebraic
    2-D circle fit, centre rotated back.  Returns (centre (3,), radius)."""
    cloud = np.asarray(cloud)
    P_mean = cloud.mean(axis=0)
    P_centered = cloud - P_mean
    _, _, V = np.linalg.svd(P_centered, full_matrices=False)
    normal = V[2, :]
    P_xy = rodrigues_rot(P_centered, normal, [0, 0, 1])
    x, y = P_xy[:, 0], P_xy[:, 1]
    A = np.array([x, y, np.ones(len(x))]).T
    b = x ** 2 + y ** 2
    c = np.linalg.lstsq(A, b, rcond=None)[0]
    xc, yc = c[0] / 2, c[1] / 2
    r = np.sqrt(c[2] + xc ** 2 + yc ** 2)
    C = rodrigues_rot(np.array([xc, yc, 0]), [0, 0, 1], normal) + P_mean
    return C.flatten(), float(r)


def fit_cylinder_v2(points, normals, weights):
    """primitive_forward_v2.py:810-849: axis from the SVD of the weighted normals (nearest third of the points when
    there are more than 600), circle fit of the points projected along the axis."""
    points, normals, weights = points.float(), normals.float(), weights.float()
    wn = weights * normals
    if wn.shape[0] > 600:
        index = _crop_nearest(points, wn.shape[0] // 3)
        wn, points = wn[index], points[index]
    _, _, V = torch.svd(wn)
    a = V[:, -1].reshape((3, 1))
    a = a / (torch.norm(a, 2) + EPS)
    prj = points - ((points @ a).permute(1, 0) * a).permute(1, 0)
    center, radius = circle_segmentation(prj.numpy())
    return a, torch.from_numpy(center), radius


def fit_cone_v2(points, normals, weights):
    """primitive_forward_v2.py:851-891: nearest half of the points; apex from the unweighted least squares
    normals . c = normals . p; axis = plane fit of the POINTS, oriented from points[0] towards the apex, snapped to a
    coordinate axis when within 0.98; apex coordinates below 0.1 snapped to 0."""
    index = _crop_nearest(points, points.shape[0] // 2)
    points, normals, weights = points[index], normals[index], weights[index]
    N = points.shape[0]
    Y = torch.sum(normals * points, 1).reshape((N, 1))
    c = torch.linalg.lstsq(normals, Y).solution[:3].reshape((3, 1))
    a, _ = O.fit_plane(points, None, weights)
    if ((c.reshape((3,)) - points[0]) * a.reshape((3,))).sum() < 0:
        a = -a
    for i in range(3):
        if torch.abs(a[0, i]) >= 0.98:
            tmp = torch.zeros_like(a)
            tmp[0, i] = 1 if a[0, i] > 0 else -1
            a = tmp
        if torch.abs(c[i, 0]) <= 0.1:
            c[i, 0] = 0
    diff = torch.nn.functional.normalize(points - c.transpose(1, 0), p=2, dim=1) @ a.transpose(1, 0)
    diff = torch.clamp(torch.abs(diff), max=0.999)
    theta = torch.sum(weights * torch.acos(diff)) / (torch.sum(weights) + EPS)
    theta = torch.clamp(theta, min=1e-3, max=3.142 / 2 - 1e-3)
    return c, a, theta


def stage2_golden_cases(synth, seeds=(301, 302, 303), n=6000, n_patches=8):
    """The seeded segments oracle/make_golden_v2.py ran the reference on (tests regenerate them from the seeds):
    yields (key, type id, points (m,3) f32, normals (m,3) f32)."""
    for seed in seeds:
        pts, nrm, lab, typ, _ = synth.make_cloud(seed, n, n_patches=n_patches, normal_jitter=0.01)
        rng = np.random.default_rng(seed)
        pts = (pts + 2e-4 * rng.normal(size=pts.shape)).astype(np.float32)
        for s in range(int(lab.max()) + 1):
            m = lab == s
            yield f"c{seed}_s{s}", int(typ[m][0]), pts[m], nrm[m]


# ------------------------------------------------------------------------------------------------ wire format
def format_stage1(points, normals, inst, types, edges_softmax=None):
    """The text the stage-1 driver writes for stage 2 (generate_predictions_aug.py:424-437): returns
    {suffix: text} for _GT_points.txt ('%0.4f', ';'), _inst.txt, _type.txt ('%d'), _edge.txt ('%0.4f', ';')."""
    import io

    def dump(a, fmt, delim=" "):
        s = io.StringIO()
        np.savetxt(s, a, fmt=fmt, delimiter=delim)
        return s.getvalue()

    out = {"_GT_points.txt": dump(np.concatenate((points, normals), axis=-1), "%0.4f", ";"),
           "_inst.txt": dump(inst, "%d"), "_type.txt": dump(types, "%d")}
    if edges_softmax is not None:
        out["_edge.txt"] = dump(edges_softmax, "%0.4f", ";")
    return out
