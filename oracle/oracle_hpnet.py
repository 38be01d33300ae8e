"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

CPU restatement of the reference's hpnet_process:
    compute_entropy                              src/smooth_normal_matrix.py:95-154
    hpnet_process with a cache hit               :157-233 (the spectral vectors v and their entropy are loaded from
                                                 src/normal_smooth_cache/, :184-189)
    knn_idx, construction_affinity_matrix_normal :31-39, :42-92 (dense (N, N) tensors, as the reference builds them)
    spectral_vectors                             :190-192 (torch.lobpcg on that matrix, k = 12, niter = 10, + row
                                                 normalisation); the start block is an argument: the reference draws it
                                                 with an unseeded torch.randn, so its own result is not reproducible
Pinned by oracle/make_golden_hpnet.py against the unmodified reference (tests/golden/hpnet.npz, hpnet_spectral.npz).
Third-party arithmetic: torch.lobpcg (PyTorch, not vendored or pinned by the reference) -- executed as is here."""
import numpy as np
import torch


def compute_entropy(features, CHUNK=2000):
    """features (1,N,K) -> 0-d tensor E (:95-154).  Only rows [0, 5 * CHUNK) enter the pairwise sums."""
    feat = features[0]
    N, K = feat.shape
    ITER = 5
    blocks = [feat[i * CHUNK:(i + 1) * CHUNK] for i in range(ITER)]
    mx, mn = [], []
    for a in blocks:
        for b in blocks:
            d = (a[:, None, :] - b[None, :, :]).view(-1, K)
            mx.append(torch.max(d, dim=0)[0][None]); mn.append(torch.min(d, dim=0)[0][None])
    interval = torch.max(torch.cat(mx, 0), 0)[0] - torch.min(torch.cat(mn, 0), 0)[0]
    average_dst = 0
    for a in blocks:
        for b in blocks:
            average_dst += torch.sum(torch.norm((a[:, None, :] - b[None, :, :]) / interval, dim=2))
    average_dst /= (N * N)
    alpha = -np.log(0.5) / average_dst
    E = 0
    for a in blocks:
        for b in blocks:
            s = torch.exp(-alpha * torch.norm((a[:, None, :] - b[None, :, :]) / interval, dim=2))
            E += torch.sum(-s * torch.log(s + 1e-7) - (1 - s) * torch.log(1 - s + 1e-7))
    return E / (N * N)


def hpnet_combine(affinity_feat, v, ent_v, types=None, edges=None, normal_smooth_w=0.5, CHUNK=2000):
    """hpnet_process on a cache hit (:157-233): (1,N,K) features, (1,N,12) cached spectral vectors and their cached
    entropy, (1,N,6) type log-probabilities, (1,N,2) edge logits -> (1,N,K+12[+6[+2]]) weighted concatenation."""
    parts = [affinity_feat * (1.7 - float(compute_entropy(affinity_feat, CHUNK))), v * (normal_smooth_w - float(ent_v))]
    if types is not None:
        t = torch.exp(types)
        if edges is not None:
            t = torch.cat((t, torch.softmax(edges, dim=-1)), dim=-1)
        parts.append(t * (0.25 - float(compute_entropy(t, CHUNK))))
    return torch.cat(parts, dim=-1)


def hpnet_case(seed, n, k=128):
    """Seeded inputs: clustered features (not normalised), unit spectral vectors, type log-probabilities, edge logits."""
    g = torch.Generator().manual_seed(seed)
    lab = torch.randint(0, 7, (n,), generator=g)
    feat = (torch.randn((7, k), generator=g)[lab] + 0.3 * torch.randn((n, k), generator=g))[None]
    v = torch.randn((1, n, 12), generator=g)
    v = v / (torch.norm(v, dim=-1, keepdim=True) + 1e-16)
    types = torch.log_softmax(2.0 * torch.randn((1, n, 6), generator=g), -1)
    edges = torch.randn((1, n, 2), generator=g)
    return feat, v, types, edges


def knn_idx(x, k):
    """src/smooth_normal_matrix.py:9-39: x (B,N,3) -> (B,N,k) indices of the k LARGEST squared distances (torch.topk's
    default largest=True: the farthest points)."""
    dist = -2 * torch.matmul(x, x.permute(0, 2, 1))                                  # :25
    dist += torch.sum(x ** 2, -1).view(x.shape[0], x.shape[1], 1)                    # :26
    dist += torch.sum(x ** 2, -1).view(x.shape[0], 1, x.shape[1])                    # :27
    return dist.topk(k=k, dim=-1)[1]                                                 # :37


def construction_affinity_matrix_normal(inputs_xyz, N_gt, sigma=0.1, knn=50):
    """src/smooth_normal_matrix.py:42-92 -> dense (B,N,N)."""
    B, N, _ = N_gt.shape
    normal = N_gt.transpose(1, 2).contiguous()                                       # :53
    A = torch.zeros(B, N, N).float()                                                 # :54
    nnid = knn_idx(inputs_xyz, knn)                                                  # :67
    k = nnid.shape[-1]
    n_sub = torch.gather(normal, -1, nnid.view(B, 1, -1).repeat(1, 3, 1)).view(B, 3, -1, k)        # :72
    dst = torch.acos((normal.unsqueeze(-1) * n_sub).sum(1).clamp(-0.99, 0.99))       # :74
    dst = torch.exp(-dst ** 2 / (2 * sigma * sigma))                                 # :76
    A = A.scatter_add(-1, nnid, dst)                                                 # :78
    A[A == 0] = 1e-12                                                                # :79-82
    D = torch.diag_embed(1.0 / A.sum(-1).sqrt())                                     # :84-85
    A = torch.matmul(torch.matmul(D, A), D)                                          # :86
    mask = (A > 0).float()                                                           # :88
    return (A + A.permute(0, 2, 1)) / (mask + mask.permute(0, 2, 1)).clamp(1, 2)     # :89-90


def spectral_vectors(inputs_xyz, normals, X0, sigma=0.1, knn=50, topk=12, niter=10):
    """src/smooth_normal_matrix.py:190-192 with the start block X0 (B,N,topk) given: (B,N,topk) row-normalised Ritz vectors
    and the Ritz values."""
    A = construction_affinity_matrix_normal(inputs_xyz, normals, sigma=sigma, knn=knn)
    E, v = torch.lobpcg(A, k=topk, niter=niter, X=X0)
    return v / (torch.norm(v, dim=-1, keepdim=True) + 1e-16), E
