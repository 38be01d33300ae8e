// Per-shape evaluation of the reference driver on the device:
//   SIOU_matched_segments[_usecd], mean_IOU_primitive_segment[_usecd], relaxed_iou_fast, compute_type_miou_abc
//       src/segment_utils.py:140-243, 300-357, 359-495, 609-627
//   chamfer_distance                                                                  src/utils.py:273-296
//
// The reference builds (N x 50) one-hot matrices, multiplies them, and then walks the matched segment pairs on the host
// with one boolean mask pass over the N points per pair (plus an n_r x n_c distance matrix per pair for the chamfer
// recall).  Every quantity those loops produce is a function of a few small integer tables -- the K x K confusion
// matrix of (predicted, ground-truth) labels, the per-segment histograms of the point types, the first point of every
// ground-truth segment -- which one pass over the points fills in shared memory; the chamfer terms of ALL matched pairs
// come from one masked N x N nearest-point pass.  The Hungarian assignment itself (50 x 50) stays on the host, as in
// the reference.
#include "internal.h"

namespace sed {

constexpr int MT_KMAX = 64, MT_TMAX = 16;

__global__ void __launch_bounds__(1024) segment_tables_kernel(const long long* __restrict__ pred, const long long* __restrict__ gt,
                                                              const long long* __restrict__ tpred, const long long* __restrict__ tgt,
                                                              int N, int K, int T, int* __restrict__ confusion,
                                                              int* __restrict__ npred, int* __restrict__ ngt,
                                                              int* __restrict__ pred_types, int* __restrict__ gt_types,
                                                              int* __restrict__ gt_first) {
    __shared__ int conf[MT_KMAX * MT_KMAX];
    __shared__ int pt[MT_KMAX * MT_TMAX], gtt[MT_KMAX * MT_TMAX];
    __shared__ int np_[MT_KMAX], ng_[MT_KMAX], first[MT_KMAX];
    const int b = blockIdx.x;
    for (int i = threadIdx.x; i < K * K; i += blockDim.x) conf[i] = 0;
    for (int i = threadIdx.x; i < K * T; i += blockDim.x) { pt[i] = 0; gtt[i] = 0; }
    for (int i = threadIdx.x; i < K; i += blockDim.x) { np_[i] = 0; ng_[i] = 0; first[i] = N; }
    __syncthreads();
    const long long o = (long long)b * N;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const long long r = pred[o + i], c = gt[o + i];
        const bool rv = r >= 0 && r < K, cv = c >= 0 && c < K;
        if (rv) atomicAdd(&np_[r], 1);
        if (cv) { atomicAdd(&ng_[c], 1); atomicMin(&first[c], i); }
        if (rv && cv) atomicAdd(&conf[r * K + c], 1);
        if (rv && tpred) { const long long t = tpred[o + i]; if (t >= 0 && t < T) atomicAdd(&pt[r * T + t], 1); }
        if (cv && tgt) { const long long t = tgt[o + i]; if (t >= 0 && t < T) atomicAdd(&gtt[c * T + t], 1); }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K * K; i += blockDim.x) confusion[(long long)b * K * K + i] = conf[i];
    for (int i = threadIdx.x; i < K * T; i += blockDim.x) {
        pred_types[(long long)b * K * T + i] = pt[i];
        gt_types[(long long)b * K * T + i] = gtt[i];
    }
    for (int i = threadIdx.x; i < K; i += blockDim.x) {
        npred[(long long)b * K + i] = np_[i];
        ngt[(long long)b * K + i] = ng_[i];
        gt_first[(long long)b * K + i] = first[i];
    }
}

// sum_n [type[n] == l] * w[n, k] (primitive_type_segment_torch, src/segment_utils.py:509-517): one CTA per cloud column block
__global__ void __launch_bounds__(256) type_vote_weighted_kernel(const long long* __restrict__ types, const float* __restrict__ w,
                                                                 int N, int K, int T, float* __restrict__ out) {
    __shared__ double acc[MT_TMAX][32];
    const int k = blockIdx.x * 32 + (threadIdx.x & 31), sub = threadIdx.x >> 5;   // 8 point-strided sub-sums per column
    double loc[MT_TMAX];
#pragma unroll
    for (int l = 0; l < MT_TMAX; ++l) loc[l] = 0.0;
    if (k < K)
        for (int n = sub; n < N; n += 8) {
            const long long t = types[n];
            const double v = (double)w[(long long)n * K + k];
#pragma unroll
            for (int l = 0; l < MT_TMAX; ++l) loc[l] += (t == l) ? v : 0.0;
        }
    for (int l = threadIdx.x; l < MT_TMAX * 32; l += 256) acc[l / 32][l % 32] = 0.0;
    __syncthreads();
    for (int s = 0; s < 8; ++s) {       // fixed order: deterministic
        if (sub == s && k < K)
#pragma unroll
            for (int l = 0; l < MT_TMAX; ++l) acc[l][threadIdx.x & 31] += loc[l];
        __syncthreads();
    }
    if (sub == 0 && k < K)
        for (int l = 0; l < T; ++l) out[(long long)l * K + k] = (float)acc[l][threadIdx.x & 31];
}

constexpr int CH_THREADS = 128, CH_TILE = 1024;

__device__ __forceinline__ float sq_dist_rn(float ax, float ay, float az, float bx, float by, float bz) {
    // torch.sum((a - b) ** 2, -1) in FP32: no FMA contraction (src/utils.py:289-290)
    const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// out[i] = min_j |q_i - r_j|^2 (the row / column minima of src/utils.py:294)
__global__ void __launch_bounds__(CH_THREADS) nearest_sq_kernel(const float* __restrict__ Q, int n, const float* __restrict__ R,
                                                                int m, float* __restrict__ out) {
    __shared__ float4 tile[CH_TILE];
    const int b = blockIdx.y, i = blockIdx.x * CH_THREADS + threadIdx.x;
    const float* q = Q + (long long)b * n * 3;
    const float* r = R + (long long)b * m * 3;
    const bool live = i < n;
    const float x = live ? q[3 * i] : 0.f, y = live ? q[3 * i + 1] : 0.f, z = live ? q[3 * i + 2] : 0.f;
    float best = INFINITY;
    for (int j0 = 0; j0 < m; j0 += CH_TILE) {
        const int cnt = min(CH_TILE, m - j0);
        __syncthreads();
        for (int t = threadIdx.x; t < cnt; t += CH_THREADS) tile[t] = make_float4(r[3 * (j0 + t)], r[3 * (j0 + t) + 1], r[3 * (j0 + t) + 2], 0.f);
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < cnt; ++j) best = fminf(best, sq_dist_rn(x, y, z, tile[j].x, tile[j].y, tile[j].z));
    }
    if (live) out[(long long)b * n + i] = best;
}

// For every point i: min_pred[i] = min over the points j of the ground-truth segment matched to i's predicted segment,
// min_gt[i] = min over the points j of the predicted segment matched to i's ground-truth segment (+inf when unmatched).
__global__ void __launch_bounds__(CH_THREADS) matched_chamfer_kernel(const float* __restrict__ pts, const long long* __restrict__ pred,
                                                                     const long long* __restrict__ gt, const int* __restrict__ pred2gt,
                                                                     const int* __restrict__ gt2pred, int N, int K,
                                                                     float* __restrict__ min_pred, float* __restrict__ min_gt) {
    __shared__ float4 tile[CH_TILE];
    const int b = blockIdx.y, i = blockIdx.x * CH_THREADS + threadIdx.x;
    const float* P = pts + (long long)b * N * 3;
    const long long* pl = pred + (long long)b * N;
    const long long* gl = gt + (long long)b * N;
    const bool live = i < N;
    const float x = live ? P[3 * i] : 0.f, y = live ? P[3 * i + 1] : 0.f, z = live ? P[3 * i + 2] : 0.f;
    int want_gt = -1, want_pred = -1;
    if (live) {
        const long long r = pl[i], c = gl[i];
        if (r >= 0 && r < K) want_gt = pred2gt[(long long)b * K + r];
        if (c >= 0 && c < K) want_pred = gt2pred[(long long)b * K + c];
    }
    float b1 = INFINITY, b2 = INFINITY;
    for (int j0 = 0; j0 < N; j0 += CH_TILE) {
        const int cnt = min(CH_TILE, N - j0);
        __syncthreads();
        for (int t = threadIdx.x; t < cnt; t += CH_THREADS) {
            const long long r = pl[j0 + t], c = gl[j0 + t];
            const unsigned lab = (unsigned)((r >= 0 && r < K) ? (int)r : 0xff) | ((unsigned)((c >= 0 && c < K) ? (int)c : 0xff) << 8);
            tile[t] = make_float4(P[3 * (j0 + t)], P[3 * (j0 + t) + 1], P[3 * (j0 + t) + 2], __uint_as_float(lab));
        }
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < cnt; ++j) {
            const float4 c = tile[j];
            const unsigned lab = __float_as_uint(c.w);
            const float d = sq_dist_rn(x, y, z, c.x, c.y, c.z);
            if ((int)(lab >> 8) == want_gt) b1 = fminf(b1, d);
            if ((int)(lab & 0xffu) == want_pred) b2 = fminf(b2, d);
        }
    }
    if (live) { min_pred[(long long)b * N + i] = b1; min_gt[(long long)b * N + i] = b2; }
}

}  // namespace sed

using namespace sed;

extern "C" {

int sed_segment_tables(const int64_t* pred, const int64_t* gt, const int64_t* type_pred, const int64_t* type_gt, int B, int N,
                       int K, int T, int* confusion, int* npred, int* ngt, int* pred_types, int* gt_types, int* gt_first,
                       sed_stream_t stream) {
    if (!pred || !gt || !confusion || !npred || !ngt || !pred_types || !gt_types || !gt_first) return SED_ERR_ARG;
    if (B <= 0 || N <= 0 || K <= 0 || T <= 0) return SED_ERR_ARG;
    if (K > MT_KMAX || T > MT_TMAX) return SED_ERR_UNSUPPORTED;
    segment_tables_kernel<<<B, 1024, 0, (cudaStream_t)stream>>>((const long long*)pred, (const long long*)gt,
                                                               (const long long*)type_pred, (const long long*)type_gt, N, K, T,
                                                               confusion, npred, ngt, pred_types, gt_types, gt_first);
    SED_CHECK_LAUNCH();
    return SED_OK;
}

int sed_type_vote_weighted(const int64_t* types, const float* weights, int N, int K, int T, float* out, sed_stream_t stream) {
    if (!types || !weights || !out || N <= 0 || K <= 0 || T <= 0) return SED_ERR_ARG;
    if (T > MT_TMAX) return SED_ERR_UNSUPPORTED;
    type_vote_weighted_kernel<<<(K + 31) / 32, 256, 0, (cudaStream_t)stream>>>((const long long*)types, weights, N, K, T, out);
    SED_CHECK_LAUNCH();
    return SED_OK;
}

int sed_chamfer_min(const float* a, const float* b, int B, int n, int m, float* min_a, float* min_b, sed_stream_t stream) {
    if (!a || !b || !min_a || !min_b || B <= 0 || n <= 0 || m <= 0) return SED_ERR_ARG;
    nearest_sq_kernel<<<dim3((n + CH_THREADS - 1) / CH_THREADS, B), CH_THREADS, 0, (cudaStream_t)stream>>>(a, n, b, m, min_a);
    SED_CHECK_LAUNCH();
    nearest_sq_kernel<<<dim3((m + CH_THREADS - 1) / CH_THREADS, B), CH_THREADS, 0, (cudaStream_t)stream>>>(b, m, a, n, min_b);
    SED_CHECK_LAUNCH();
    return SED_OK;
}

int sed_matched_chamfer(const float* points, const int64_t* pred, const int64_t* gt, const int* pred2gt, const int* gt2pred, int B,
                        int N, int K, float* min_pred, float* min_gt, sed_stream_t stream) {
    if (!points || !pred || !gt || !pred2gt || !gt2pred || !min_pred || !min_gt || B <= 0 || N <= 0 || K <= 0) return SED_ERR_ARG;
    if (K > 254) return SED_ERR_UNSUPPORTED;
    matched_chamfer_kernel<<<dim3((N + CH_THREADS - 1) / CH_THREADS, B), CH_THREADS, 0, (cudaStream_t)stream>>>(
        points, (const long long*)pred, (const long long*)gt, pred2gt, gt2pred, N, K, min_pred, min_gt);
    SED_CHECK_LAUNCH();
    return SED_OK;
}

}  // extern "C"
