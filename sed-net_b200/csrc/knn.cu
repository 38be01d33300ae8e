// Fused pairwise-distance + streaming top-k ("kNN graph") kernels.
//
// Replaces reference src/PointNet.py:62-87 (knn), :90-137 (knn_points_normals) and the
// N x N Gram + topk temporaries of src/mean_shift.py:130-135 (bandwidth) and :146-149
// (nms membership).  The N x N matrix is never materialised: a CTA owns TQ query rows,
// streams candidate tiles of 128 points through shared memory, evaluates the metric with
// register-tiled FP32 FFMA in the reference's operation order, and keeps the running top-k of
// each row in shared memory (threshold filter + lazy bitonic merge).
#include <stdlib.h>
#include <string.h>

#include "internal.h"

namespace sed {

enum { M_L2 = 0, M_PN = 1, M_COS = 2, M_FAR = 3 };
enum { L_CHANNEL_MAJOR = 0, L_ROW_MAJOR = 1 };

constexpr int KNN_THREADS = 256;
constexpr int TC = 128;  // candidates per tile
constexpr int CK = 32;   // channels per staged chunk

struct KnnParams {
    const float* xq;   // queries
    const float* xc;   // candidates
    long long q_bstride, c_bstride;  // elements between consecutive clouds
    int ldq, ldc;      // channel-major: elements between channels; row-major: elements between rows
    int C, Nq, Nc, k;
    float W;           // normal_metric_W (M_PN)
    void* out_idx;     // (B, Nq, k) int64 or int32, nearest first (may be null)
    int idx64;
    float* out_kth;    // (B, Nq) score of rank k-1 (may be null)
    const int* nc_ptr; // optional per-cloud candidate count (<= Nc); a count <= 0 writes index 0
};

template <typename KeyT> __device__ __forceinline__ KeyT make_key(float score, int j);
template <> __device__ __forceinline__ unsigned long long make_key<unsigned long long>(float score, int j) {
    // higher score first; among equal scores the lower index first
    return ((unsigned long long)f2ord(score) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)j);
}
template <> __device__ __forceinline__ uint32_t make_key<uint32_t>(float score, int j) { return f2ord(score); }

// Bitonic sort (descending) of n (power of two) keys in shared memory by one warp.
template <typename KeyT>
__device__ __forceinline__ void warp_bitonic_desc(KeyT* a, int n, int lane) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = lane; t < (n >> 1); t += 32) {
                int i = 2 * t - (t & (j - 1));
                int l = i + j;
                bool up = ((i & k) == 0);
                KeyT x = a[i], y = a[l];
                if ((x < y) == up) { a[i] = y; a[l] = x; }
            }
            __syncwarp();
        }
    }
}

// Loads one CK x TC chunk of candidates (or queries) into registers. Returns values in v[16];
// thread->element mapping depends on the layout (see store_chunk).
template <int LAYOUT>
__device__ __forceinline__ void load_chunk(float (&v)[16], const float* __restrict__ xb, int ld, int C, int N,
                                           int c0, int j0, int tid) {
    if (LAYOUT == L_CHANNEL_MAJOR) {
        const int j = j0 + (tid & 127);
        const int cb = c0 + (tid >> 7);
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            int c = cb + 2 * r;
            v[r] = (c < C && j < N) ? __ldg(xb + (long long)c * ld + j) : 0.f;
        }
    } else {
        const int j = j0 + (tid & 127);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            int c = c0 + 4 * ((tid >> 7) + 2 * r);
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < C && j < N) t = __ldg(reinterpret_cast<const float4*>(xb + (long long)j * ld + c));
            v[4 * r + 0] = t.x; v[4 * r + 1] = t.y; v[4 * r + 2] = t.z; v[4 * r + 3] = t.w;
        }
    }
}

// Stores the staged chunk to shared memory as [c][ncols]; accumulates this thread's share of the squared
// norm of column (tid & 127) over channels < normC.
template <int LAYOUT>
__device__ __forceinline__ void store_chunk(const float (&v)[16], float* __restrict__ dst, int ncols, int c0,
                                            int normC, float& nrm, int tid) {
    const int j = tid & 127;
    if (j >= ncols) return;
    if (LAYOUT == L_CHANNEL_MAJOR) {
        const int cb = tid >> 7;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            int c = cb + 2 * r;
            dst[c * ncols + j] = v[r];
            if (c0 + c < normC) nrm = fmaf(v[r], v[r], nrm);
        }
    } else {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            int c = 4 * ((tid >> 7) + 2 * r);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                dst[(c + e) * ncols + j] = v[4 * r + e];
                if (c0 + c + e < normC) nrm = fmaf(v[4 * r + e], v[4 * r + e], nrm);
            }
        }
    }
}

template <int QPT>
__device__ __forceinline__ void load_qfrag(float (&a)[QPT], const float* __restrict__ src) {
    if (QPT == 4) {
        float4 t = *reinterpret_cast<const float4*>(src);
        a[0] = t.x; a[1] = t.y; a[2] = t.z; a[3] = t.w;
    } else {
        float2 t = *reinterpret_cast<const float2*>(src);
        a[0] = t.x; a[1] = t.y;
    }
}

template <int METRIC>
__device__ __forceinline__ float metric_score(float dot, float dot2, float xxq, float xxc, float W) {
    if (METRIC == M_L2) {
        // src/PointNet.py:76-78: inner = -2 x.x' ; pd = -xx_j - inner - xx_i
        float inner = -2.0f * dot;
        return __fsub_rn(__fsub_rn(-xxc, inner), xxq);
    } else if (METRIC == M_PN) {
        // src/PointNet.py:112-120,128
        float pd = __fadd_rn(__fsub_rn(xxc, 2.0f * dot), xxq);
        float nd = __fsub_rn(2.0f, 2.0f * dot2);
        return -__fmul_rn(pd, __fadd_rn(1.0f, __fmul_rn(nd, W)));
    } else if (METRIC == M_FAR) {
        // src/smooth_normal_matrix.py:25-28,35-37: dist = -2 x.y ; dist += |x|^2 ; dist += |y|^2 ; topk keeps the LARGEST
        return __fadd_rn(__fadd_rn(-2.0f * dot, xxq), xxc);
    } else {
        // src/mean_shift.py:130,146: dist = 2 - 2 x.y ; score = -dist
        return -__fsub_rn(2.0f, 2.0f * dot);
    }
}

// TQ: query rows per CTA (64 or 32). TOTAL: power-of-two capacity of list + overflow buffer per row.
// KMAXL: sorted-list capacity (>= k). TOP1: keep only the best candidate per row in registers.
template <int METRIC, int LAYOUT, typename KeyT, int TQ, int KMAXL, int TOTAL, bool TOP1>
__global__ void __launch_bounds__(KNN_THREADS, 1) knn_kernel(KnnParams p) {
    constexpr int QPT = TQ / 16;        // query rows per thread
    constexpr int CAP = TOTAL - KMAXL;  // overflow buffer capacity
    static_assert(TOP1 || CAP >= TC, "overflow buffer must take one full tile");
    extern __shared__ __align__(16) unsigned char smem_raw[];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ty = tid >> 4, tx = tid & 15;
    const int b = blockIdx.y;
    const int q0 = blockIdx.x * TQ;
    const int C = p.C;
    const int Cp = ((C + CK - 1) / CK) * CK;
    const int nchunk = Cp / CK;
    const int normC = (METRIC == M_L2 || METRIC == M_FAR) ? C : (METRIC == M_PN ? 3 : 0);

    float* Qs = reinterpret_cast<float*>(smem_raw);  // [Cp][TQ]
    float* Xs = Qs + Cp * TQ;                         // [2][CK][TC]
    float* xxq = Xs + 2 * CK * TC;                    // [TQ]
    float* xxp = xxq + TQ;                            // [2][TC] partial candidate norms
    int* cnt = reinterpret_cast<int*>(xxp + 2 * TC);  // [TQ]
    KeyT* thr = reinterpret_cast<KeyT*>(cnt + TQ);    // [TQ]
    KeyT* keys = thr + TQ;                            // [TQ][TOTAL]

    const float* xq = p.xq + (long long)b * p.q_bstride;
    const float* xc = p.xc + (long long)b * p.c_bstride;

    // ---- stage the query tile (all channels) and its norms -------------------------------------------
    {
        float qn = 0.f;
        for (int ch = 0; ch < nchunk; ++ch) {
            float v[16];
            // reuse the candidate loader with a 128-wide window; only the first TQ columns are kept
            load_chunk<LAYOUT>(v, xq, p.ldq, C, p.Nq, ch * CK, q0, tid);
            store_chunk<LAYOUT>(v, Qs + ch * CK * TQ, TQ, ch * CK, normC, qn, tid);
        }
        if ((tid & 127) < TQ) xxp[(tid >> 7) * TC + (tid & 127)] = qn;
        if (!TOP1) {
            for (int i = tid; i < TQ * TOTAL; i += KNN_THREADS) keys[i] = 0;
            if (tid < TQ) { cnt[tid] = 0; thr[tid] = 0; }
        }
        __syncthreads();
        if (tid < TQ) xxq[tid] = xxp[tid] + xxp[TC + tid];
        __syncthreads();
    }

    const int Nc = p.nc_ptr ? min(max(p.nc_ptr[b], 0), p.Nc) : p.Nc;
    const int ntiles = (Nc + TC - 1) / TC;
    const int nsteps = ntiles * nchunk;
    float stage[16];
    float cn = 0.f;
    load_chunk<LAYOUT>(stage, xc, p.ldc, C, Nc, 0, 0, tid);
    store_chunk<LAYOUT>(stage, Xs, TC, 0, normC, cn, tid);
    __syncthreads();

    float acc[QPT][8];
    float acc2[QPT][8];
    unsigned long long best[QPT];
#pragma unroll
    for (int i = 0; i < QPT; ++i) best[i] = 0ull;

    for (int step = 0; step < nsteps; ++step) {
        const int tile = step / nchunk, ch = step - tile * nchunk;
        const int buf = step & 1;
        const bool has_next = (step + 1 < nsteps);
        if (ch == 0) {
#pragma unroll
            for (int i = 0; i < QPT; ++i)
#pragma unroll
                for (int e = 0; e < 8; ++e) { acc[i][e] = 0.f; acc2[i][e] = 0.f; }
        }
        if (has_next) {
            const int nt = (step + 1) / nchunk, nc = (step + 1) - nt * nchunk;
            load_chunk<LAYOUT>(stage, xc, p.ldc, C, Nc, nc * CK, nt * TC, tid);
        }
        // ---- register-tiled FFMA over this chunk ----
        {
            const float* xs = Xs + buf * CK * TC;
            const float* qs = Qs + ch * CK * TQ + ty * QPT;
            if (METRIC == M_PN) {
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    float a[QPT];
                    load_qfrag<QPT>(a, qs + c * TQ);
                    float4 b0 = *reinterpret_cast<const float4*>(xs + c * TC + tx * 4);
                    float4 b1 = *reinterpret_cast<const float4*>(xs + c * TC + 64 + tx * 4);
                    float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                    for (int i = 0; i < QPT; ++i)
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            if (c < 3) acc[i][e] = fmaf(a[i], bb[e], acc[i][e]);
                            else acc2[i][e] = fmaf(a[i], bb[e], acc2[i][e]);
                        }
                }
            } else {
                const int cmax = min(CK, C - ch * CK);
#pragma unroll 8
                for (int c = 0; c < cmax; ++c) {
                    float a[QPT];
                    load_qfrag<QPT>(a, qs + c * TQ);
                    float4 b0 = *reinterpret_cast<const float4*>(xs + c * TC + tx * 4);
                    float4 b1 = *reinterpret_cast<const float4*>(xs + c * TC + 64 + tx * 4);
                    float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                    for (int i = 0; i < QPT; ++i)
#pragma unroll
                        for (int e = 0; e < 8; ++e) acc[i][e] = fmaf(a[i], bb[e], acc[i][e]);
                }
            }
        }
        const bool tile_done = (ch == nchunk - 1);
        if (tile_done) {
            // candidate norms of this tile are complete (cn covers all chunks of the tile)
            xxp[(tid >> 7) * TC + (tid & 127)] = cn;
            cn = 0.f;
        }
        if (has_next) {
            const int nt = (step + 1) / nchunk, nc = (step + 1) - nt * nchunk;
            store_chunk<LAYOUT>(stage, Xs + (buf ^ 1) * CK * TC, TC, nc * CK, normC, cn, tid);
        }
        __syncthreads();
        if (!tile_done) continue;

        // ---- tile epilogue: score, filter against the row threshold, append ----
        const int j0 = tile * TC;
#pragma unroll
        for (int i = 0; i < QPT; ++i) {
            const int ql = ty * QPT + i;
            const bool qok = (q0 + ql) < p.Nq;
            const float xq_i = xxq[ql];
            KeyT th = 0;
            if (!TOP1) th = thr[ql];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int jl = (e < 4) ? (tx * 4 + e) : (64 + tx * 4 + (e - 4));
                const int j = j0 + jl;
                if (j < Nc && qok) {
                    float xc_j = xxp[jl] + xxp[TC + jl];
                    float s = metric_score<METRIC>(acc[i][e], acc2[i][e], xq_i, xc_j, p.W);
                    if (TOP1) {
                        unsigned long long kk = make_key<unsigned long long>(s, j);
                        best[i] = kk > best[i] ? kk : best[i];
                    } else {
                        KeyT kk = make_key<KeyT>(s, j);
                        if (kk > th) {
                            int pos = atomicAdd(&cnt[ql], 1);
                            keys[ql * TOTAL + KMAXL + pos] = kk;
                        }
                    }
                }
            }
        }
        __syncthreads();
        if (TOP1) continue;
        // ---- lazy merge: only rows whose buffer could overflow on the next tile (or at the end) ----
        const bool last = (tile == ntiles - 1);
        for (int ql = warp; ql < TQ; ql += KNN_THREADS / 32) {
            const int c = cnt[ql];
            if (c > CAP - TC || (last && c > 0)) {
                KeyT* row = keys + ql * TOTAL;
                warp_bitonic_desc<KeyT>(row, TOTAL, lane);
                KeyT nth = row[p.k - 1];
                __syncwarp();
                for (int t = KMAXL + lane; t < TOTAL; t += 32) row[t] = 0;
                if (lane == 0) { cnt[ql] = 0; thr[ql] = nth; }
                __syncwarp();
            }
        }
        __syncthreads();
    }

    // ---- write results ----
    if (TOP1) {
#pragma unroll
        for (int i = 0; i < QPT; ++i) {
            unsigned long long v = best[i];
#pragma unroll
            for (int m = 8; m > 0; m >>= 1) {
                unsigned long long o = shfl_xor_u64(v, m);
                v = o > v ? o : v;
            }
            const int q = q0 + ty * QPT + i;
            if (tx == 0 && q < p.Nq) {
                int j = v ? (int)(0xFFFFFFFFu - (uint32_t)(v & 0xFFFFFFFFull)) : 0;
                if (p.idx64) reinterpret_cast<long long*>(p.out_idx)[(long long)b * p.Nq + q] = j;
                else reinterpret_cast<int*>(p.out_idx)[(long long)b * p.Nq + q] = j;
                if (p.out_kth) p.out_kth[(long long)b * p.Nq + q] = ord2f((uint32_t)(v >> 32));
            }
        }
        return;
    }
    const int k = p.k;
    if (p.out_idx != nullptr && sizeof(KeyT) == 8) {
        for (int e = tid; e < TQ * k; e += KNN_THREADS) {
            const int ql = e / k, r = e - ql * k;
            const int q = q0 + ql;
            if (q < p.Nq) {
                unsigned long long kk = (unsigned long long)keys[ql * TOTAL + r];
                int j = (int)(0xFFFFFFFFu - (uint32_t)(kk & 0xFFFFFFFFull));
                long long o = ((long long)b * p.Nq + q) * k + r;
                if (p.idx64) reinterpret_cast<long long*>(p.out_idx)[o] = j;
                else reinterpret_cast<int*>(p.out_idx)[o] = j;
            }
        }
    }
    if (p.out_kth != nullptr) {
        for (int ql = tid; ql < TQ; ql += KNN_THREADS) {
            const int q = q0 + ql;
            if (q < p.Nq) {
                KeyT kk = keys[ql * TOTAL + k - 1];
                uint32_t o = (sizeof(KeyT) == 8) ? (uint32_t)((unsigned long long)kk >> 32) : (uint32_t)kk;
                p.out_kth[(long long)b * p.Nq + q] = ord2f(o);
            }
        }
    }
}

template <int METRIC, int LAYOUT, typename KeyT, int TQ, int KMAXL, int TOTAL, bool TOP1>
static int launch_knn(const KnnParams& p, int B, cudaStream_t st) {
    const int Cp = ((p.C + CK - 1) / CK) * CK;
    size_t smem = (size_t)(Cp * TQ + 2 * CK * TC + TQ + 2 * TC) * sizeof(float) + TQ * sizeof(int) + TQ * sizeof(KeyT);
    if (!TOP1) smem += (size_t)TQ * TOTAL * sizeof(KeyT);
    smem = (smem + 15) & ~(size_t)15;
    auto kern = knn_kernel<METRIC, LAYOUT, KeyT, TQ, KMAXL, TOTAL, TOP1>;
    if (smem > 227 * 1024) return SED_ERR_UNSUPPORTED;
    SED_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((p.Nq + TQ - 1) / TQ, B);
    kern<<<grid, KNN_THREADS, smem, st>>>(p);
    SED_CHECK_LAUNCH();
    return SED_OK;
}

template <int METRIC, int LAYOUT>
static int dispatch_knn_idx(const KnnParams& p, int B, cudaStream_t st) {
    typedef unsigned long long K64;
    if (p.k <= 32) return launch_knn<METRIC, LAYOUT, K64, 64, 32, 256, false>(p, B, st);
    if (p.k <= 64) return launch_knn<METRIC, LAYOUT, K64, 64, 64, 256, false>(p, B, st);
    if (p.k <= 128) return launch_knn<METRIC, LAYOUT, K64, 32, 128, 512, false>(p, B, st);
    if (p.k <= 256) return launch_knn<METRIC, LAYOUT, K64, 32, 256, 512, false>(p, B, st);
    return SED_ERR_UNSUPPORTED;
}

// mean of sqrt(max(-score_kth, 1e-6)) over rows: src/mean_shift.py:135-137. One CTA per cloud, fixed order.
__global__ void bandwidth_mean_kernel(const float* __restrict__ kth, int N, float* __restrict__ bw_out, float min_bw) {
    __shared__ double sh[32];
    const int b = blockIdx.x;
    double s = 0.0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        float d = -kth[(long long)b * N + i];
        s += (double)sqrtf(fmaxf(d, 1e-6f));
    }
    s = warp_sum_d(s);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
        float bw = (float)(t / (double)N);
        bw_out[b] = fmaxf(bw, min_bw);
    }
}

}  // namespace sed

namespace sed {

// tensor-core implementations (select_tc.cu); SED_ERR_UNSUPPORTED when the shape is outside their range
int knn_tc(const float* x, long long bstride, int B, int C, int N, int k, int pn, float W, void* idx, int idx64,
           cudaStream_t st, int sorted);
int cos_select_tc(const float* Q, const float* Cand, int B, int Nq, int Nc, const int* nc_ptr, int d, int K,
                  float* kth_out, void* idx_out, int idx64, cudaStream_t st);

// SEDNET_B200_SELECT=ffma forces the CUDA-core kernels of this file (A/B comparisons); default: tensor cores
static bool use_tc() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("SEDNET_B200_SELECT");
        v = (e && !strcmp(e, "ffma")) ? 0 : 1;
    }
    return v == 1;
}

int knn_l2(const float* x, long long bstride, int B, int C, int N, int k, void* idx, int idx64, cudaStream_t st, int sorted) {
    if (!x || !idx || B <= 0 || C <= 0 || C > 256 || N < k || k <= 0) return SED_ERR_ARG;
    if (use_tc()) {
        const int rc = knn_tc(x, bstride, B, C, N, k, 0, 0.f, idx, idx64, st, sorted);
        if (rc != SED_ERR_UNSUPPORTED) return rc;
    }
    KnnParams p{};
    p.xq = p.xc = x; p.q_bstride = p.c_bstride = bstride; p.ldq = p.ldc = N;
    p.C = C; p.Nq = p.Nc = N; p.k = k; p.W = 0.f; p.out_idx = idx; p.idx64 = idx64;
    return dispatch_knn_idx<M_L2, L_CHANNEL_MAJOR>(p, B, st);
}

int knn_pn(const float* x6, long long bstride, int B, int N, int k, float W, void* idx, int idx64, cudaStream_t st, int sorted) {
    if (!x6 || !idx || B <= 0 || N < k || k <= 0) return SED_ERR_ARG;
    if (use_tc()) {
        const int rc = knn_tc(x6, bstride, B, 6, N, k, 1, W, idx, idx64, st, sorted);
        if (rc != SED_ERR_UNSUPPORTED) return rc;
    }
    KnnParams p{};
    p.xq = p.xc = x6; p.q_bstride = p.c_bstride = bstride; p.ldq = p.ldc = N;
    p.C = 6; p.Nq = p.Nc = N; p.k = k; p.W = W; p.out_idx = idx; p.idx64 = idx64;
    return dispatch_knn_idx<M_PN, L_CHANNEL_MAJOR>(p, B, st);
}

// The k FARTHEST points of every point (largest squared distance first): the reference's knn_idx, whose topk keeps the
// largest entries of the distance matrix (src/smooth_normal_matrix.py:31-39).  x channel-major (B,C,N), k <= 256.
int knn_far(const float* x, long long bstride, int B, int C, int N, int k, void* idx, int idx64, cudaStream_t st) {
    if (!x || !idx || B <= 0 || C <= 0 || C > 256 || N < k || k <= 0) return SED_ERR_ARG;
    KnnParams p{};
    p.xq = p.xc = x; p.q_bstride = p.c_bstride = bstride; p.ldq = p.ldc = N;
    p.C = C; p.Nq = p.Nc = N; p.k = k; p.W = 0.f; p.out_idx = idx; p.idx64 = idx64;
    return dispatch_knn_idx<M_FAR, L_CHANNEL_MAJOR>(p, B, st);
}

// Nearest candidate (cosine distance 2 - 2 q.c, lowest index on ties) for every query row.
// Q (B,Nq,d), Cand (B,Nc,d) row-major; nc_ptr optional per-cloud candidate count; out (B,Nq) int32/int64.
int nearest_cos(const float* Q, const float* Cand, int B, int Nq, int Nc, const int* nc_ptr, int d, void* out,
                int idx64, cudaStream_t st) {
    if (!Q || !Cand || !out || B <= 0 || d <= 0 || d > 256 || (d & 3) || Nq <= 0 || Nc <= 0) return SED_ERR_ARG;
    if (use_tc()) {
        const int rc = cos_select_tc(Q, Cand, B, Nq, Nc, nc_ptr, d, 1, nullptr, out, idx64, st);
        if (rc != SED_ERR_UNSUPPORTED) return rc;
    }
    KnnParams p{};
    p.xq = Q; p.xc = Cand; p.q_bstride = (long long)Nq * d; p.c_bstride = (long long)Nc * d; p.ldq = p.ldc = d;
    p.C = d; p.Nq = Nq; p.Nc = Nc; p.k = 1; p.out_idx = out; p.idx64 = idx64; p.nc_ptr = nc_ptr;
    return launch_knn<M_COS, L_ROW_MAJOR, unsigned long long, 64, 32, 256, true>(p, B, st);
}

}  // namespace sed

using namespace sed;

extern "C" {

int sed_knn_l2(const float* x, int B, int C, int N, int k, void* idx, int idx64, sed_stream_t stream) {
    return knn_l2(x, (long long)C * N, B, C, N, k, idx, idx64, (cudaStream_t)stream);
}

int sed_knn_pn(const float* x6, int B, int N, int k, float W, void* idx, int idx64, sed_stream_t stream) {
    return knn_pn(x6, 6LL * N, B, N, k, W, idx, idx64, (cudaStream_t)stream);
}

int sed_ms_bandwidth(const float* X, int B, int N, int d, int K, float min_bw, float* kth_ws, float* bw,
                     sed_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!X || !kth_ws || !bw || B <= 0 || d <= 0 || d > 256 || (d & 3) || K <= 0 || N < K) return SED_ERR_ARG;
    KnnParams p{};
    p.xq = p.xc = X; p.q_bstride = p.c_bstride = (long long)N * d; p.ldq = p.ldc = d;
    p.C = d; p.Nq = p.Nc = N; p.k = K; p.W = 0.f; p.out_idx = nullptr; p.idx64 = 0; p.out_kth = kth_ws;
    int rc = use_tc() ? cos_select_tc(X, X, B, N, N, nullptr, d, K, kth_ws, nullptr, 0, st) : SED_ERR_UNSUPPORTED;
    if (rc == SED_OK) {
        bandwidth_mean_kernel<<<B, 1024, 0, st>>>(kth_ws, N, bw, min_bw);
        SED_CHECK_LAUNCH();
        return SED_OK;
    }
    if (rc != SED_ERR_UNSUPPORTED) return rc;
    if (K <= 64) rc = launch_knn<M_COS, L_ROW_MAJOR, uint32_t, 64, 64, 256, false>(p, B, st);
    else if (K <= 256) rc = launch_knn<M_COS, L_ROW_MAJOR, uint32_t, 64, 256, 512, false>(p, B, st);
    else if (K <= 512) rc = launch_knn<M_COS, L_ROW_MAJOR, uint32_t, 32, 512, 1024, false>(p, B, st);
    else return SED_ERR_UNSUPPORTED;
    if (rc != SED_OK) return rc;
    bandwidth_mean_kernel<<<B, 1024, 0, st>>>(kth_ws, N, bw, min_bw);
    SED_CHECK_LAUNCH();
    return SED_OK;
}

}  // extern "C"
