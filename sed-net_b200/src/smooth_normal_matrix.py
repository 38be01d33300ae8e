"""src/smooth_normal_matrix.py of the reference with its names and signatures: compute_entropy (:95-154), knn_idx (:31-39),
construction_affinity_matrix_normal (:42-92) and hpnet_process (:157-233), both branches.

compute_entropy runs on the pairwise kernel of libsednet_b200.so (no CHUNK x CHUNK x K temporaries).  The branch of
hpnet_process that BUILDS the spectral vectors (:190-196) never forms the reference's dense (N, N) matrices: the farthest-k
index table comes from the kNN kernel (sed_far_idx; the reference's topk keeps the k largest distances), the affinity matrix
lives in factored form (sed_affinity_normal_build: k weights per row + D^-1/2; AffinityOperator) and torch.lobpcg's
iteration is restated here around its block product A X (sed_affinity_matmul: N k m multiply-adds instead of N^2 m).

torch.lobpcg (PyTorch, un-vendored and un-pinned by the reference) is restated from its published algorithm -- method
"ortho" of torch/_lobpcg.py: Rayleigh-Ritz over [X, P, W] with SVQB orthonormalisation [DuerschEtal2018,
StathopoulosWu2002] -- including its quirks (istep 0 forms the residual with E = 0; niter counts the initial step).  The
tiny dense algebra of the iteration (eigh / qr / cholesky of matrices up to 36 x 36, N x 36 block updates) stays on
torch.linalg, the reference's own library; the only N x N operator is the kernel's.  The reference starts the iteration
from an unseeded torch.randn block and stops after 10 steps on a spectrum whose top is nearly degenerate (one eigenvalue
close to 1 per connected patch of parallel normals), so its vectors are not reproducible run to run; `X=` fixes the start
(tests feed the same block to torch.lobpcg on the reference's dense matrix)."""
import os

import torch

from . import _lib


def compute_entropy(features, CHUNK=2000):
    """features (1,N,K) -> 0-d tensor E (src/smooth_normal_matrix.py:95-154)."""
    assert features.shape[0] == 1
    feat = _lib.require_cuda(features[0], name="features")
    N, K = feat.shape
    out = torch.empty(3, dtype=torch.float32, device=feat.device)
    _lib.call("sed_compute_entropy", _lib.ptr(feat), N, K, int(CHUNK), _lib.ptr(out), _lib.stream())
    return out[0]


def knn_idx(x, k):
    """src/smooth_normal_matrix.py:31-39: x (B,N,3) -> (B,N,k) int64, the k largest entries of every row of the squared
    distance matrix, largest first (torch.topk's default: the FARTHEST points; kept as the reference computes it)."""
    x = _lib.require_cuda(x, name="x")
    B, N, _ = x.shape
    idx = torch.empty((B, N, k), dtype=torch.int32, device=x.device)
    _lib.call("sed_far_idx", _lib.ptr(x), B, N, int(k), _lib.ptr(idx), _lib.stream())
    return idx.to(torch.int64)


class AffinityOperator:
    """The (B, N, N) matrix of construction_affinity_matrix_normal in factored form: idx, w (B,N,k), dinv (B,N).
    ``matmul(b, X)`` is its block product for cloud b; ``to_dense()`` rebuilds the reference's tensor (small N only)."""

    def __init__(self, idx, w, dinv):
        self.idx, self.w, self.dinv = idx, w, dinv
        self.shape = (idx.shape[0], idx.shape[1], idx.shape[1])
        self.device = idx.device
        self._ws = {}

    def matmul(self, b, X):
        N, k = self.idx.shape[1], self.idx.shape[2]
        X = _lib.require_cuda(X, name="X")
        if b not in self._ws:
            ws = torch.empty(_lib.load().sed_affinity_workspace_bytes(N, k), dtype=torch.uint8, device=self.device)
            _lib.call("sed_affinity_prepare", _lib.ptr(self.idx[b]), _lib.ptr(self.w[b]), N, k, _lib.ptr(ws), _lib.stream())
            self._ws[b] = ws
        Y = torch.empty_like(X)
        _lib.call("sed_affinity_matmul", _lib.ptr(self.idx[b]), _lib.ptr(self.w[b]), _lib.ptr(self.dinv[b]),
                  _lib.ptr(self._ws[b]), _lib.ptr(X), N, k, X.shape[1], _lib.ptr(Y), _lib.stream())
        return Y

    def to_dense(self):
        B, N, k = self.idx.shape
        a = torch.full((B, N, N), 1e-12, dtype=torch.float32, device=self.device)
        nz = self.w != 0
        rows = torch.arange(N, device=self.device).view(1, N, 1).expand(B, N, k)
        bb = torch.arange(B, device=self.device).view(B, 1, 1).expand(B, N, k)
        a[bb[nz], rows[nz], self.idx[nz].long()] = self.w[nz]
        m = a * self.dinv[:, :, None] * self.dinv[:, None, :]
        return (m + m.transpose(1, 2)) / 2


def construction_affinity_matrix_normal(inputs_xyz, N_gt, sigma=0.1, knn=50):
    """src/smooth_normal_matrix.py:42-92: inputs_xyz, N_gt (normals) (B,N,3) -> the affinity matrix, as an AffinityOperator
    (the dense (B,N,N) tensor of the reference is ``.to_dense()``)."""
    xyz = _lib.require_cuda(inputs_xyz, name="inputs_xyz")
    nrm = _lib.require_cuda(N_gt, name="N_gt")
    B, N, _ = nrm.shape
    idx = torch.empty((B, N, knn), dtype=torch.int32, device=xyz.device)
    _lib.call("sed_far_idx", _lib.ptr(xyz), B, N, int(knn), _lib.ptr(idx), _lib.stream())
    w = torch.empty((B, N, knn), dtype=torch.float32, device=xyz.device)
    dinv = torch.empty((B, N), dtype=torch.float32, device=xyz.device)
    _lib.call("sed_affinity_normal_build", _lib.ptr(nrm), _lib.ptr(idx), B, N, int(knn), float(sigma), _lib.ptr(w),
              _lib.ptr(dinv), _lib.stream())
    return AffinityOperator(idx, w, dinv)


# ---------------------------------------------------------------------------------------------- torch.lobpcg, method "ortho"
def _symeig(M, largest=False):
    E, Z = torch.linalg.eigh(M, UPLO="U")
    return (torch.flip(E, dims=(-1,)), torch.flip(Z, dims=(-1,))) if largest else (E, Z)


def _svqb(U, tau):
    """torch/_lobpcg.py LOBPCG._get_svqb with B = I, drop = False."""
    if U.numel() == 0:
        return U
    UBU = U.mT @ U
    d = UBU.diagonal(0, -2, -1)
    nz = torch.where(d.abs() != 0.0)[0]
    if nz.numel() < d.numel():                       # exact zero columns are dropped
        U = U[:, nz]
        if U.numel() == 0:
            return U
        UBU = U.mT @ U
        d = UBU.diagonal(0, -2, -1)
    d_col = (d ** -0.5).reshape(-1, 1)
    E, Z = _symeig((UBU * d_col) * d_col.mT)
    t = tau * E.abs().max()
    E = torch.where(E < t, t, E)
    return (U * d_col.mT) @ (Z * E ** -0.5)


def _ortho(U, V, tol, i_max=3, j_max=3):
    """LOBPCG._get_ortho with B = I, ortho_use_drop = False: U orthonormal with columns orthogonal to V."""
    BV_norm = torch.norm(V)
    VBU = V.mT @ U
    for _ in range(i_max):
        U = U - V @ VBU
        for _ in range(j_max):
            U = _svqb(U, tol)
            if U.numel() == 0:
                return U
            UBU = U.mT @ U
            U_norm = torch.norm(U)
            R = UBU - torch.eye(UBU.shape[-1], device=U.device, dtype=U.dtype)
            if float(torch.norm(R)) * float(U_norm * U_norm) ** -1 < tol:
                break
        VBU = V.mT @ U
        if float(torch.norm(VBU)) * float(BV_norm * torch.norm(U)) ** -1 < tol:
            break
    return U


def lobpcg_top(matmul, N, k, niter, X):
    """torch.lobpcg(A, k=k, niter=niter, X=X)[0:2] for a symmetric operator given by its block product `matmul` (largest
    eigenpairs, B = None, iK = None, method "ortho", tol = sqrt(float32 eps)): E (k,), X (N,k)."""
    n = X.shape[-1]
    tol = 1.2e-07 ** 0.5
    X = X.clone()
    E = torch.zeros(n, dtype=X.dtype, device=X.device)
    S = torch.zeros((N, 3 * n), dtype=X.dtype, device=X.device)
    X_norm = float(torch.norm(X))
    A_norm = float(torch.norm(matmul(X))) / X_norm
    B_norm = 1.0
    nc = ns = 0

    def converged_count(R, prev):
        rerr = torch.norm(R, 2, (0,)) / (torch.norm(X, 2, (0,)) * (A_norm + E.abs() * B_norm))
        count = 0
        for c in (rerr < tol).tolist():
            if not c:
                break
            count += 1
        return max(count, prev)

    for istep in range(niter):                       # niter counts the initial step (iterations_left bookkeeping)
        if istep == 0:
            SBS = X.mT @ X                           # _get_rayleigh_ritz_transform
            d_row = SBS.diagonal(0, -2, -1) ** -0.5
            Rc = torch.linalg.cholesky((SBS * d_row) * d_row.reshape(-1, 1), upper=True)
            Ri = torch.linalg.solve_triangular(Rc, d_row.diag_embed(), upper=True, left=False)
            M = Ri.mT @ (X.mT @ matmul(X)) @ Ri
            _, Z = _symeig(M, True)
            X = X @ (Ri @ Z)
            R = matmul(X) - X * E                    # E is still zero here, as in torch
            nc = converged_count(R, 0)
            S[:, :n] = X
            W = _ortho(R, X, tol)
            ns = n + W.shape[-1]
            S[:, n:ns] = W
        else:
            S_ = S[:, nc:ns]
            E_, Z = _symeig(S_.mT @ matmul(S_.contiguous()), True)
            X[:, nc:] = S_ @ Z[:, :n - nc]
            E[nc:] = E_[:n - nc]
            P = S_ @ (Z[:, n - nc:] @ torch.linalg.qr(Z[:n - nc, n - nc:].mT).Q)
            np_ = P.shape[-1]
            R = matmul(X) - X * E
            nc = converged_count(R, nc)
            S[:, :n] = X
            S[:, n:n + np_] = P
            W = _ortho(R[:, nc:], S[:, :n + np_], tol)
            ns = n + np_ + W.shape[-1]
            S[:, n + np_:ns] = W
        if nc >= k:
            break
    return E[:k], X[:, :k]


def hpnet_process(affinity_feat, inputs_xyz, normals, id=None, types=None, edges=None, normal_smooth_w=0.5, CHUNK=2000,
                  gpu='cuda:0', drop_rest_idx=None, v=None, ent=None, X=None):
    """src/smooth_normal_matrix.py:157-233: affinity_feat (B,N,K) without L2, types (B,N,6) log-probabilities, edges
    (B,N,2) logits -> (B,N,K+12[+6[+2]]) weighted concatenation [features, spectral vectors, type/edge probabilities].
    Extra keywords: v, ent -- spectral vectors (B,N,12) and their entropy handed in; X -- (B,N,12) start block of the
    eigen-iteration (the reference draws it with an unseeded torch.randn)."""
    weight_ent, parts = [], []
    weight_ent.append(1.7 - float(compute_entropy(affinity_feat, CHUNK=CHUNK)))                    # :172-174
    parts.append(affinity_feat)
    edge_topk, normal_sigma, edge_knn = 12, 0.1, 50                                                # :179-182
    fn = "src/normal_smooth_cache/Us_{}_{}_{}.pt".format(id, normal_sigma, edge_knn)
    fn_ent = "src/normal_smooth_cache/WUs_{}_{}_{}.pt".format(id, normal_sigma, edge_knn)
    if v is None:
        if id is not None and os.path.exists(fn) and os.path.exists(fn_ent):                       # :186-189
            v = torch.load(fn).to(affinity_feat.device)
            ent = torch.load(fn_ent)
        else:                                                                                      # :190-196
            op = construction_affinity_matrix_normal(inputs_xyz, normals, sigma=normal_sigma, knn=edge_knn)
            B, N = op.shape[0], op.shape[1]
            vs = []
            for b in range(B):
                X0 = torch.randn((N, edge_topk), dtype=torch.float32, device=op.device) if X is None else X[b].to(op.device)
                vs.append(lobpcg_top(lambda Y, b=b: op.matmul(b, Y), N, edge_topk, 10, X0)[1])
            v = torch.stack(vs)
            v = v / (torch.norm(v, dim=-1, keepdim=True) + 1e-16)
            ent = compute_entropy(v, CHUNK=CHUNK)
            if os.path.isdir(os.path.dirname(fn)):          # the reference writes its cache unconditionally (:193,196)
                torch.save(v, fn)
                torch.save(ent, fn_ent)
    if ent is None:
        ent = compute_entropy(v, CHUNK=CHUNK)                                                      # :195
    if drop_rest_idx is not None:
        v = v[:, drop_rest_idx, :]
    weight_ent.append(normal_smooth_w - float(ent))                                                # :202-205
    parts.append(v)
    if types is not None:                                                                          # :208-216
        types = torch.exp(types)
        if edges is not None:
            types = torch.cat((types, torch.softmax(edges, dim=-1)), dim=-1)
        weight_ent.append(0.25 - float(compute_entropy(types, CHUNK=CHUNK)))
        parts.append(types)
    return torch.cat([p * w for p, w in zip(parts, weight_ent)], dim=-1)                           # :221-233
