// The operator behind hpnet_process's spectral ("normal smooth") embedding, reference src/smooth_normal_matrix.py:31-92,
// 190-196, without the dense (N, N) matrices the reference builds (400 MB each at N = 10 000: the distance matrix, the
// affinity matrix, two diag_embed matrices and their products).
//
//   knn_idx (:31-39)                 topk of the squared-distance matrix with torch's default largest=True: the k FARTHEST
//                                    points of every point (kept as the reference computes it) -> knn.cu, metric M_FAR.
//   construction_affinity_matrix_normal (:42-92)
//        w_ij  = exp(-acos(clamp(n_i . n_j, -0.99, 0.99))^2 / (2 sigma^2))      at the k scattered entries of row i
//        a_ij  = w_ij where scattered and w_ij != 0, 1e-12 everywhere else       (the `== 0` background mask, :76-79)
//        D_i   = sum_j a_ij,     M_ij = a_ij / sqrt(D_i D_j),     A = (M + M^T) / 2   (every entry of M is > 0, so the
//                                                                  (mask + mask^T).clamp(1, 2) divisor of :88-90 is 2)
//   Factored form kept here:  A = D^-1/2 [ bg 1 1^T + (C + C^T) / 2 ] D^-1/2  with  C sparse, c_ij = w_ij - bg on the scattered
//   non-zero entries (<= k per row), bg = 1e-12.  A block product Y = A X (the only thing torch.lobpcg needs from A) is
//        z = D^-1/2 X,   t = sum_j z_j,   Y_i = d_i [ bg t + 1/2 sum_{j in out(i)} c_ij z_j + 1/2 sum_{j in in(i)} c_ji z_j ]
//   -- N k m multiply-adds instead of N^2 m, and nothing of size N x N exists.  in(i), the transposed adjacency, is built
//   once per cloud by a counting sort whose rows are then ordered by source index, so the products are deterministic.
#include "internal.h"

namespace sed {

constexpr float kAffBg = 1e-12f;

// (B,N,3) -> (B,3,N)
__global__ void xyz_to_cm_kernel(const float* __restrict__ xyz, int N, float* __restrict__ cm) {
    const int b = blockIdx.y, n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float* p = xyz + ((long long)b * N + n) * 3;
    float* o = cm + (long long)b * 3 * N + n;
    o[0] = p[0]; o[N] = p[1]; o[2LL * N] = p[2];
}

// One warp per row: the k scattered weights and D_i^-1/2.
__global__ void __launch_bounds__(256) aff_weights_kernel(const float* __restrict__ nrm, const int* __restrict__ idx, int N, int k,
                                                          float inv_2s2, float* __restrict__ w, float* __restrict__ dinv) {
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= N) return;
    const float* nb = nrm + (long long)b * N * 3;
    const float nx = nb[3 * i], ny = nb[3 * i + 1], nz = nb[3 * i + 2];
    double sum = 0.0;
    int nnz = 0;
    for (int j = lane; j < k; j += 32) {
        const long long e = ((long long)b * N + i) * k + j;
        const int t = idx[e];
        // (normal.unsqueeze(-1) * n_sub).sum(1): x, y, z products added in that order (:69)
        float dot = __fadd_rn(__fadd_rn(__fmul_rn(nx, nb[3 * t]), __fmul_rn(ny, nb[3 * t + 1])), __fmul_rn(nz, nb[3 * t + 2]));
        dot = fminf(fmaxf(dot, -0.99f), 0.99f);
        const float a = acosf(dot);
        const float v = expf(-__fmul_rn(a, a) * inv_2s2);       // exp(-dst^2 / (2 sigma^2)) (:71)
        w[e] = v;
        if (v != 0.f) { sum += (double)v; ++nnz; }
    }
    sum = warp_sum_d(sum);
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) nnz += __shfl_xor_sync(0xffffffffu, nnz, m);
    if (lane == 0) {
        const float D = (float)(sum + (double)(N - nnz) * (double)kAffBg);      // affinity_matrix.sum(-1) (:81)
        dinv[(long long)b * N + i] = __fdiv_rn(1.0f, sqrtf(D));                   // 1.0 / D.sqrt() (:82)
    }
}

// ---- transposed adjacency (one cloud): count, scan, fill, sort rows by source
__global__ void aff_count_kernel(const int* __restrict__ idx, const float* __restrict__ w, int N, int k, int* __restrict__ cnt) {
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= (long long)N * k) return;
    if (w[e] != 0.f) atomicAdd(cnt + idx[e], 1);
}
// exclusive scan of cnt[0..N) into start[0..N], single CTA of 1024 threads; also copies start into cursor
__global__ void __launch_bounds__(1024) aff_scan_kernel(const int* __restrict__ cnt, int N, int* __restrict__ start, int* __restrict__ cursor) {
    __shared__ int part[1024];
    const int per = (N + 1023) / 1024, lo = threadIdx.x * per, hi = min(N, lo + per);
    int s = 0;
    for (int i = lo; i < hi; ++i) s += cnt[i];
    part[threadIdx.x] = s;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        const int v = threadIdx.x >= off ? part[threadIdx.x - off] : 0;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    int run = part[threadIdx.x] - s;
    for (int i = lo; i < hi; ++i) { start[i] = run; cursor[i] = run; run += cnt[i]; }
    if (threadIdx.x == 1023) start[N] = part[1023];
}
__global__ void aff_fill_kernel(const int* __restrict__ idx, const float* __restrict__ w, int N, int k, int* __restrict__ cursor,
                                int* __restrict__ src, float* __restrict__ val) {
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= (long long)N * k) return;
    const float v = w[e];
    if (v == 0.f) return;
    const int pos = atomicAdd(cursor + idx[e], 1);
    src[pos] = (int)(e / k);
    val[pos] = v;
}
// one CTA per row: order the row's (source, value) pairs by source (sources are distinct: a row scatters to k distinct
// columns).  Rank sort: every element counts the smaller sources of its row -- O(L^2 / 256) per row; the farthest-point
// tables concentrate on a few extreme points, whose in-lists can hold thousands of entries, hence a whole CTA per row.
__global__ void __launch_bounds__(256) aff_sort_kernel(const int* __restrict__ start, int N, const int* __restrict__ src,
                                                       const float* __restrict__ val, int* __restrict__ src_o, float* __restrict__ val_o) {
    const int i = blockIdx.x;
    const int lo = start[i], hi = start[i + 1];
    for (int a = lo + threadIdx.x; a < hi; a += 256) {
        const int s = src[a];
        int r = 0;
        for (int c = lo; c < hi; ++c) r += (src[c] < s) ? 1 : 0;
        src_o[lo + r] = s;
        val_o[lo + r] = val[a];
    }
}

// z = d (.) X and the per-CTA column sums of z (FP64 partials, fixed order)
constexpr int AFF_ROWS = 256;   // rows per CTA of the scale kernel
__global__ void __launch_bounds__(256) aff_scale_kernel(const float* __restrict__ X, const float* __restrict__ dinv, int N, int m,
                                                        float* __restrict__ Z, double* __restrict__ part) {
    __shared__ double sh[8][64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r0 = blockIdx.x * AFF_ROWS;
    double s0 = 0.0, s1 = 0.0;
    for (int r = r0 + warp; r < min(N, r0 + AFF_ROWS); r += 8) {
        const float d = dinv[r];
        if (lane < m) { const float z = __fmul_rn(d, X[(long long)r * m + lane]); Z[(long long)r * m + lane] = z; s0 += (double)z; }
        if (lane + 32 < m) { const float z = __fmul_rn(d, X[(long long)r * m + lane + 32]); Z[(long long)r * m + lane + 32] = z; s1 += (double)z; }
    }
    sh[warp][lane] = s0; sh[warp][lane + 32] = s1;
    __syncthreads();
    if (threadIdx.x < 64) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += sh[w][threadIdx.x];
        part[(long long)blockIdx.x * 64 + threadIdx.x] = t;
    }
}
__global__ void aff_colsum_kernel(const double* __restrict__ part, int nblk, float* __restrict__ t) {
    const int c = threadIdx.x;   // 64 threads
    double s = 0.0;
    for (int b = 0; b < nblk; ++b) s += part[(long long)b * 64 + c];
    t[c] = (float)s;
}
// one warp per row: Y_i = d_i [ bg t + 1/2 sum_out c_ij z_j + 1/2 sum_in c_ji z_j ]
__global__ void __launch_bounds__(256) aff_apply_kernel(const int* __restrict__ idx, const float* __restrict__ w, const float* __restrict__ dinv,
                                                        const int* __restrict__ start, const int* __restrict__ src, const float* __restrict__ val,
                                                        const float* __restrict__ Z, const float* __restrict__ t, int N, int k, int m,
                                                        float* __restrict__ Y) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= N) return;
    const bool c0 = lane < m, c1 = lane + 32 < m;
    float a0 = 0.f, a1 = 0.f;
    for (int j = 0; j < k; ++j) {
        const float v = w[(long long)i * k + j];
        if (v == 0.f) continue;                                   // warp-uniform
        const float c = v - kAffBg;
        const float* zr = Z + (long long)idx[(long long)i * k + j] * m;
        if (c0) a0 = fmaf(c, zr[lane], a0);
        if (c1) a1 = fmaf(c, zr[lane + 32], a1);
    }
    for (int e = start[i]; e < start[i + 1]; ++e) {
        const float c = val[e] - kAffBg;
        const float* zr = Z + (long long)src[e] * m;
        if (c0) a0 = fmaf(c, zr[lane], a0);
        if (c1) a1 = fmaf(c, zr[lane + 32], a1);
    }
    const float d = dinv[i];
    if (c0) Y[(long long)i * m + lane] = d * fmaf(0.5f, a0, kAffBg * t[lane]);
    if (c1) Y[(long long)i * m + lane + 32] = d * fmaf(0.5f, a1, kAffBg * t[lane + 32]);
}

struct AffWs {
    int *cnt, *start, *cursor, *src_tmp, *src;
    float *val_tmp, *val, *Z, *t;
    double* part;
};
static void carve_aff(Arena& A, int N, int k, AffWs& w) {
    const int64_t E = (int64_t)N * k;
    w.cnt = A.take<int>(N + 1);
    w.start = A.take<int>(N + 1);
    w.cursor = A.take<int>(N + 1);
    w.src_tmp = A.take<int>(E);
    w.src = A.take<int>(E);
    w.val_tmp = A.take<float>(E);
    w.val = A.take<float>(E);
    w.Z = A.take<float>((int64_t)N * 64);
    w.t = A.take<float>(64);
    w.part = A.take<double>((int64_t)((N + AFF_ROWS - 1) / AFF_ROWS) * 64);
}

}  // namespace sed

using namespace sed;

extern "C" {

int sed_far_idx(const float* xyz, int B, int N, int k, int* idx, sed_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!xyz || !idx || B <= 0 || N <= 0 || k <= 0 || k > N || k > 256) return SED_ERR_ARG;
    ensure_pool_config();
    float* cm = nullptr;
    SED_CUDA(cudaMallocAsync((void**)&cm, (size_t)B * 3 * N * sizeof(float), st));
    xyz_to_cm_kernel<<<dim3((N + 255) / 256, B), 256, 0, st>>>(xyz, N, cm);
    ++g_sed_launches;
    const int rc = knn_far(cm, 3LL * N, B, 3, N, k, idx, 0, st);
    cudaFreeAsync(cm, st);
    return rc;
}

int sed_affinity_normal_build(const float* normals, const int* idx, int B, int N, int k, float sigma, float* w, float* dinv,
                              sed_stream_t stream) {
    if (!normals || !idx || !w || !dinv || B <= 0 || N <= 0 || k <= 0 || !(sigma > 0.f)) return SED_ERR_ARG;
    aff_weights_kernel<<<dim3((N + 7) / 8, B), 256, 0, (cudaStream_t)stream>>>(normals, idx, N, k, 1.0f / (2.0f * sigma * sigma), w, dinv);
    SED_CHECK_LAUNCH();
    return SED_OK;
}

int64_t sed_affinity_workspace_bytes(int N, int k) {
    Arena A(nullptr, 0);
    AffWs w;
    carve_aff(A, N, k, w);
    return A.off;
}

int sed_affinity_prepare(const int* idx, const float* w, int N, int k, void* workspace, sed_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!idx || !w || !workspace || N <= 0 || k <= 0) return SED_ERR_ARG;
    Arena A(workspace, sed_affinity_workspace_bytes(N, k));
    AffWs a;
    carve_aff(A, N, k, a);
    const long long E = (long long)N * k;
    SED_CUDA(cudaMemsetAsync(a.cnt, 0, (size_t)(N + 1) * sizeof(int), st));
    aff_count_kernel<<<(unsigned)((E + 255) / 256), 256, 0, st>>>(idx, w, N, k, a.cnt);
    aff_scan_kernel<<<1, 1024, 0, st>>>(a.cnt, N, a.start, a.cursor);
    aff_fill_kernel<<<(unsigned)((E + 255) / 256), 256, 0, st>>>(idx, w, N, k, a.cursor, a.src_tmp, a.val_tmp);
    aff_sort_kernel<<<N, 256, 0, st>>>(a.start, N, a.src_tmp, a.val_tmp, a.src, a.val);
    g_sed_launches += 4;
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SED_OK : SED_ERR_CUDA_BASE - (int)e;
}

int sed_affinity_matmul(const int* idx, const float* w, const float* dinv, void* workspace, const float* X, int N, int k, int m,
                        float* Y, sed_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!idx || !w || !dinv || !workspace || !X || !Y || N <= 0 || k <= 0 || m <= 0 || m > 64) return SED_ERR_ARG;
    Arena A(workspace, sed_affinity_workspace_bytes(N, k));
    AffWs a;
    carve_aff(A, N, k, a);
    const int nblk = (N + AFF_ROWS - 1) / AFF_ROWS;
    aff_scale_kernel<<<nblk, 256, 0, st>>>(X, dinv, N, m, a.Z, a.part);
    aff_colsum_kernel<<<1, 64, 0, st>>>(a.part, nblk, a.t);
    aff_apply_kernel<<<(N + 7) / 8, 256, 0, st>>>(idx, w, dinv, a.start, a.src, a.val, a.Z, a.t, N, k, m, Y);
    g_sed_launches += 3;
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SED_OK : SED_ERR_CUDA_BASE - (int)e;
}

}  // extern "C"
