"""Seeded synthetic ParseNet-shape inputs (no dataset is reachable offline).

Each cloud is a union of analytic patches (planes, spheres, cylinders, cones)
with exact unit normals, sampled to exactly N points, then normalised the way
the reference's loader does it: centre, divide by the largest axis extent,
rotate the smallest-variance PCA axis onto +x (reference
src/dataset_segments.py:376-379,400-417).  Primitive type ids follow the
reference's convention (1 plane, 3 cone, 4 cylinder, 5 sphere;
src/primitive_forward.py:1006-1021).

numpy only; shared by tests/, bench.py and oracle/make_golden.py.
"""
import numpy as np

PLANE, CONE, CYLINDER, SPHERE = 1, 3, 4, 5
EPS = float(np.finfo(np.float32).eps)


def _frame(axis):
    axis = axis / np.linalg.norm(axis)
    t = np.array([1.0, 0, 0]) if abs(axis[0]) < 0.9 else np.array([0, 1.0, 0])
    u = np.cross(axis, t)
    u /= np.linalg.norm(u)
    v = np.cross(axis, u)
    return axis, u, v


def sample_plane(rng, n):
    a, u, v = _frame(rng.normal(size=3))
    c = rng.uniform(-1, 1, 3)
    ext = rng.uniform(0.3, 0.8, 2)
    s = rng.uniform(-1, 1, (n, 2)) * ext
    p = c + s[:, :1] * u + s[:, 1:] * v
    nrm = np.repeat(a[None], n, 0)
    return p, nrm, dict(type=PLANE, axis=a, d=float(a @ c))


def sample_sphere(rng, n):
    c = rng.uniform(-1, 1, 3)
    r = rng.uniform(0.2, 0.6)
    # a cap covering a good fraction of the sphere so the fit is well conditioned
    a, u, v = _frame(rng.normal(size=3))
    cosmax = rng.uniform(-0.6, 0.2)
    z = rng.uniform(cosmax, 1.0, n)
    ph = rng.uniform(0, 2 * np.pi, n)
    s = np.sqrt(np.maximum(1 - z * z, 0))
    d = z[:, None] * a + (s * np.cos(ph))[:, None] * u + (s * np.sin(ph))[:, None] * v
    return c + r * d, d, dict(type=SPHERE, center=c, radius=r)


def sample_cylinder(rng, n):
    a, u, v = _frame(rng.normal(size=3))
    c = rng.uniform(-1, 1, 3)
    r = rng.uniform(0.1, 0.4)
    h = rng.uniform(0.4, 1.2)
    arc = rng.uniform(np.pi, 2 * np.pi)
    ph = rng.uniform(0, arc, n)
    t = rng.uniform(-h / 2, h / 2, n)
    d = np.cos(ph)[:, None] * u + np.sin(ph)[:, None] * v
    return c + r * d + t[:, None] * a, d, dict(type=CYLINDER, axis=a, center=c, radius=r)


def sample_cone(rng, n):
    a, u, v = _frame(rng.normal(size=3))
    apex = rng.uniform(-1, 1, 3)
    theta = rng.uniform(0.2, 1.2)
    l0, l1 = rng.uniform(0.15, 0.3), rng.uniform(0.6, 1.0)
    # area-uniform along the slant
    l = np.sqrt(rng.uniform(l0 * l0, l1 * l1, n))
    ph = rng.uniform(0, 2 * np.pi, n)
    rad = np.cos(ph)[:, None] * u + np.sin(ph)[:, None] * v
    p = apex + l[:, None] * (np.cos(theta) * a + np.sin(theta) * rad)
    nrm = np.cos(theta) * rad - np.sin(theta) * a  # outward normal; n . axis < 0
    return p, nrm, dict(type=CONE, apex=apex, axis=a, theta=theta)


_SAMPLERS = {PLANE: sample_plane, SPHERE: sample_sphere, CYLINDER: sample_cylinder, CONE: sample_cone}


def _pca_align(points, normals):
    """Rotate the smallest-variance axis to +x (reference src/dataset_segments.py:400-405)."""
    S, U = np.linalg.eig(points.T @ points)
    ev = np.real(U[:, np.argmin(np.real(S))])
    A, B = ev / np.linalg.norm(ev), np.array([1.0, 0, 0])
    cos = A @ B
    if abs(abs(cos) - 1) < 1e-9:
        return points, normals
    sin = np.linalg.norm(np.cross(B, A))
    v = B - cos * A
    v /= np.linalg.norm(v) + 1e-8
    w = np.cross(B, A)
    w /= np.linalg.norm(w) + 1e-8
    F = np.stack([A, v, w], 1)
    G = np.array([[cos, -sin, 0], [sin, cos, 0], [0, 0, 1]])
    R = F @ G @ np.linalg.inv(F)
    return (R @ points.T).T, (R @ normals.T).T


def make_cloud(seed, n_points=10000, n_patches=None, normal_jitter=0.0, min_pts=200):
    """One cloud: (points (N,3) f32, normals (N,3) f32, labels (N,) i64, types (N,) i64, patches)."""
    rng = np.random.default_rng(seed)
    if n_patches is None:
        n_patches = int(rng.integers(6, 21))
    n_patches = max(1, min(n_patches, n_points // min_pts))
    w = rng.uniform(0.5, 2.0, n_patches)
    cnt = np.maximum((w / w.sum() * n_points).astype(int), min_pts)
    while cnt.sum() > n_points:
        cnt[np.argmax(cnt)] -= 1
    cnt[np.argmax(cnt)] += n_points - cnt.sum()
    kinds = rng.choice([PLANE, SPHERE, CYLINDER, CONE], n_patches)
    P, Nn, L, T, patches = [], [], [], [], []
    for i, (k, c) in enumerate(zip(kinds, cnt)):
        p, nrm, info = _SAMPLERS[int(k)](rng, int(c))
        P.append(p); Nn.append(nrm); L.append(np.full(c, i)); T.append(np.full(c, int(k)))
        patches.append(info)
    P, Nn = np.concatenate(P), np.concatenate(Nn)
    L, T = np.concatenate(L), np.concatenate(T)
    perm = rng.permutation(n_points)
    P, Nn, L, T = P[perm], Nn[perm], L[perm], T[perm]
    if normal_jitter > 0:
        Nn = Nn + normal_jitter * rng.normal(size=Nn.shape)
        Nn /= np.linalg.norm(Nn, axis=1, keepdims=True)
    P = P - P.mean(0, keepdims=True)
    P = P / (np.max(P.max(0) - P.min(0)) + EPS)
    P, Nn = _pca_align(P, Nn)
    return (np.ascontiguousarray(P, dtype=np.float32), np.ascontiguousarray(Nn, dtype=np.float32),
            L.astype(np.int64), T.astype(np.int64), patches)


def make_batch(batch, n_points=10000, seed0=1234, **kw):
    """Batch of clouds with seeds seed0 + index. Returns arrays stacked on a new axis 0."""
    cl = [make_cloud(seed0 + b, n_points, **kw) for b in range(batch)]
    return tuple(np.stack([c[i] for c in cl]) for i in range(4))


def make_touching_instances(seed, n):
    """Instances that touch (stage-2 adjacency maps): a wavy sheet cut into Voronoi cells of 7 random sites (labels 0-6)
    plus one isolated blob (label 7: the 'lonely instance' branch of face_face_inter_map).  Returns points (n,3) f32,
    labels (n,) i64."""
    rng = np.random.default_rng(seed)
    uv = rng.uniform(-1, 1, (n - 200, 2))
    sheet = np.stack([uv[:, 0], uv[:, 1], 0.2 * np.sin(3 * uv[:, 0]) * np.cos(2 * uv[:, 1])], 1)
    sites = rng.uniform(-1, 1, (7, 2))
    lab = np.argmin(((uv[:, None] - sites[None]) ** 2).sum(-1), 1)
    blob = np.array([3.0, 3.0, 3.0]) + 0.05 * rng.normal(size=(200, 3))
    pts = np.concatenate([sheet, blob]).astype(np.float32)
    lab = np.concatenate([lab, np.full(200, 7)]).astype(np.int64)
    perm = rng.permutation(n)
    return pts[perm], lab[perm]


def make_metric_case(seed, n):
    """Labels for the evaluation metrics: a cloud of analytic patches with its ground-truth instance /
    type labels, and a 'prediction' that merges two instances, splits one, mislabels 3 % of the points and gets some
    per-point types wrong.  Returns points (n,3) f32, gt (n,), type_gt (n,), pred (n,), type_pred (n,) (all int64;
    pred labels are contiguous 0..C-1)."""
    pts, _, lab, typ, _ = make_cloud(seed, n, n_patches=9, min_pts=150)
    rng = np.random.default_rng(seed + 1000)
    pred = lab.copy()
    pred[pred == 1] = 0                                        # merge
    big = int(np.argmax(np.bincount(lab)))
    idx = np.flatnonzero(lab == big)
    pred[idx[: len(idx) // 3]] = int(lab.max()) + 1            # split
    flip = rng.choice(n, n * 3 // 100, replace=False)
    pred[flip] = rng.integers(0, int(pred.max()) + 1, len(flip))
    _, pred = np.unique(pred, return_inverse=True)             # contiguous ids
    type_pred = typ.copy()
    wrong = rng.choice(n, n // 10, replace=False)
    type_pred[wrong] = rng.choice([1, 3, 4, 5, 0, 8], len(wrong))
    return pts, lab.astype(np.int64), typ.astype(np.int64), pred.astype(np.int64), type_pred.astype(np.int64)


def make_embedding(labels, dim=128, sigma=0.01, seed=0):
    """Unit-norm embedding with one well separated mode per label:
    normalize(centroid[label] + sigma * randn)."""
    rng = np.random.default_rng(seed)
    n_lab = int(labels.max()) + 1
    cen = rng.normal(size=(n_lab, dim))
    cen /= np.linalg.norm(cen, axis=1, keepdims=True)
    e = cen[labels] + sigma * rng.normal(size=(labels.shape[0], dim))
    e /= np.linalg.norm(e, axis=1, keepdims=True)
    return e.astype(np.float32)


# state_dict layout of SEDNet with the inference driver's kwargs (generate_predictions_aug.py:142-154):
# name -> shape.  encoder.convK.1.* are the same modules as encoder.bnK.* (src/SEDNet.py:37-45).
STATE_SHAPES = {
    "encoder.bn1.weight": (64,), "encoder.bn1.bias": (64,),
    "encoder.bn2.weight": (64,), "encoder.bn2.bias": (64,),
    "encoder.bn3.weight": (128,), "encoder.bn3.bias": (128,),
    "encoder.bn4.weight": (256,), "encoder.bn4.bias": (256,),
    "encoder.bn5.weight": (1024,), "encoder.bn5.bias": (1024,),
    "encoder.conv1.0.weight": (64, 12, 1, 1),
    "encoder.conv1.1.weight": (64,), "encoder.conv1.1.bias": (64,),
    "encoder.conv2.0.weight": (64, 128, 1, 1),
    "encoder.conv2.1.weight": (64,), "encoder.conv2.1.bias": (64,),
    "encoder.conv3.0.weight": (128, 128, 1, 1),
    "encoder.conv3.1.weight": (128,), "encoder.conv3.1.bias": (128,),
    "encoder.mlp1.weight": (1024, 256, 1), "encoder.mlp1.bias": (1024,),
    "encoder.bnmlp1.weight": (1024,), "encoder.bnmlp1.bias": (1024,),
    "conv1.weight": (512, 1280, 1), "conv1.bias": (512,),
    "bn1.weight": (512,), "bn1.bias": (512,),
    "conv2.weight": (256, 512, 1), "conv2.bias": (256,),
    "bn2.weight": (256,), "bn2.bias": (256,),
    "edge_module.0.weight": (128, 256, 1), "edge_module.0.bias": (128,),
    "edge_module.1.weight": (128,), "edge_module.1.bias": (128,),
    "edge_module.2.weight": (2, 128, 1), "edge_module.2.bias": (2,),
    "asis.0.weight": (256, 256, 1), "asis.0.bias": (256,),
    "asis.1.weight": (256,), "asis.1.bias": (256,),
    "mlp_seg_prob1.weight": (256, 256, 1), "mlp_seg_prob1.bias": (256,),
    "mlp_seg_prob2.weight": (128, 256, 1), "mlp_seg_prob2.bias": (128,),
    "bn_seg_prob1.weight": (256,), "bn_seg_prob1.bias": (256,),
    "mlp_prim_prob1.weight": (256, 256, 1), "mlp_prim_prob1.bias": (256,),
    "mlp_prim_prob2.weight": (6, 256, 1), "mlp_prim_prob2.bias": (6,),
    "bn_prim_prob1.weight": (256,), "bn_prim_prob1.bias": (256,),
    "pos_enc.inv_freq": (128,),
    "prim_encoding.0.weight": (256, 8, 1), "prim_encoding.0.bias": (256,),
}
_GN_KEYS = ("bn", "edge_module.1", "asis.1")
_ALIASES = {"encoder.conv1.1": "encoder.bn1", "encoder.conv2.1": "encoder.bn2", "encoder.conv3.1": "encoder.bn3"}


def make_state_dict(seed=0, randomize_gn=False):
    """Random-init weights as numpy f32 arrays (no checkpoints are reachable offline).

    Conv weights/biases ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)) like torch's default init;
    GroupNorm affine = (1, 0), or with ``randomize_gn`` weight ~ N(1, 0.5) (some negative,
    exercising the min branch of the fused EdgeConv) and bias ~ N(0, 0.5).  numpy RNG so the
    same seed gives the same weights on every torch version."""
    rng = np.random.default_rng(seed)
    sd = {}
    for name, shape in STATE_SHAPES.items():
        stem, leaf = name.rsplit(".", 1)
        if stem in _ALIASES:
            continue
        is_gn = any(stem.split(".")[-1].startswith("bn") or stem.endswith(g) for g in _GN_KEYS[1:]) \
            or stem.split(".")[-1].startswith("bn")
        if name == "pos_enc.inv_freq":
            sd[name] = (1.0 / (10000 ** (np.arange(0, 256, 2, dtype=np.float32) / 256))).astype(np.float32)
        elif is_gn:
            if randomize_gn:
                sd[name] = (rng.normal(1.0 if leaf == "weight" else 0.0, 0.5, shape)).astype(np.float32)
            else:
                sd[name] = (np.ones(shape) if leaf == "weight" else np.zeros(shape)).astype(np.float32)
        else:
            wshape = STATE_SHAPES[stem + ".weight"]
            bound = 1.0 / np.sqrt(wshape[1])
            sd[name] = rng.uniform(-bound, bound, shape).astype(np.float32)
    for alias, src in _ALIASES.items():
        sd[alias + ".weight"] = sd[src + ".weight"]
        sd[alias + ".bias"] = sd[src + ".bias"]
    return {k: sd[k] for k in STATE_SHAPES}
