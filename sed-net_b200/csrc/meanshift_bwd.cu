// Backward of one mean-shift iteration (training path: src/segment_loss.py:50-56 calls mean_shift(..., nms=False) inside the
// triplet loss and lets autograd differentiate src/mean_shift.py:45-79), FP32 on the CUDA cores, sm_100a.
//
// Forward of iteration t (Gaussian kernel; the 1 / sum K factor of the reference cancels in the normalisation, and so does
// its gradient: d/dD of normalize((K X) D) is zero):
//     S = Q X^T,   P = exp(max((S - 1) / b^2, -75)),   O = P X,   Q' = O / |O|
// Given G = dL/dQ' the chain rule gives
//     dO = (G - Q' (Q' . G)) / |O|                      (rows)
//     dP = dO X^T,   dS = P o dP / b^2                  (zero where the exponent was clamped)
//     dQ = dS X,     dX = P^T dO + dS^T Q               (+ dQ of iteration 0 at the end: the iteration starts from Q = X)
// Nothing N x N is stored (the reference's autograd keeps dist, K and their products per iteration: ~6 x 400 MB at
// N = 10 000): like a flash-attention backward, P and dS are recomputed tile by tile in three passes,
//     rows A :  per 64-row block, over all keys:     S, P, O += P X           -> |O|, Q', dO
//     rows B :  per 64-row block, over all keys:     S, P, dP, dS, dQ += dS X
//     cols   :  per 64-key block, over all rows:     S^T, P^T, dP^T, dS^T, dX += P^T dO + dS^T Q
// i.e. 9 N^2 d multiply-adds per iteration.  Tiles are rows x 128 floats in shared memory (row pitch 132): a CTA keeps a block
// of 16 R rows (keys) resident and streams 64-row tiles of the other side; a thread owns an R x 4 patch of the score tile (rows
// ty + 16 r, columns tx + 16 q: conflict-free 16-byte reads) and an R x 8 patch of the accumulators.  R is chosen per call so
// that the CTAs fill whole waves of the 148 SMs (N = 10 000: R = 5, 125 CTAs, one wave; 64-row blocks would be 157 = two).  b (the bandwidth) is a constant of the graph, as in the reference (computed under no_grad).
#include <type_traits>

#include "common.cuh"
#include "internal.h"

namespace sed {

constexpr int MB_T = 64;                 // rows / keys per STREAMED tile; a CTA's resident block has 16 R rows, R = 2 ... 8
constexpr int MB_D = 128;                // channels (narrower rows are zero-padded by the loader)
constexpr int MB_P = MB_D + 4;           // row pitch of a 64 x 128 tile
constexpr int MB_WP = MB_T + 4;          // row pitch of a 64 x 64 weight tile
constexpr int MB_THREADS = 256;
constexpr int MB_TILE = MB_T * MB_P;     // floats
constexpr int MB_WTILE = MB_T * MB_WP;

struct MsbParams {
    const float* Q;      // (B,N,d) positions entering the iteration
    const float* X;      // (B,N,d) keys
    const float* G;      // (B,N,d) dL/dQ'
    const float* bw;     // (B)
    float* dO;           // (B,N,128) workspace
    float* dQ;           // (B,N,d) out
    float* dX;           // (B,N,d) accumulated
    int N, d;
};

// rows [row0, row0 + rows) of a (N, d) matrix -> tile[rows][132]; rows >= N and columns >= d read as zero
__device__ __forceinline__ void msb_load(float* tile, const float* __restrict__ src, int row0, int N, int d, int rows = MB_T) {
    for (int e = threadIdx.x; e < rows * (MB_D / 4); e += MB_THREADS) {
        const int r = e >> 5, c = (e & 31) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + r < N && c < d) v = __ldg(reinterpret_cast<const float4*>(src + (long long)(row0 + r) * d + c));
        *reinterpret_cast<float4*>(tile + r * MB_P + c) = v;
    }
}

// acc[r][q] = A[ty + 16 r] . Bm[tx + 16 q]   over the 128 channels (RA rows of A per thread, 4 of Bm)
template <int RA>
__device__ __forceinline__ void msb_dot(const float* __restrict__ A, const float* __restrict__ Bm, int ty, int tx, float (&acc)[RA][4]) {
#pragma unroll
    for (int r = 0; r < RA; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[r][q] = 0.f;
#pragma unroll 2
    for (int c = 0; c < MB_D; c += 4) {
        float4 a[RA], b[4];
#pragma unroll
        for (int r = 0; r < RA; ++r) a[r] = *reinterpret_cast<const float4*>(A + (ty + 16 * r) * MB_P + c);
#pragma unroll
        for (int q = 0; q < 4; ++q) b[q] = *reinterpret_cast<const float4*>(Bm + (tx + 16 * q) * MB_P + c);
#pragma unroll
        for (int r = 0; r < RA; ++r)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                acc[r][q] = fmaf(a[r].x, b[q].x, acc[r][q]);
                acc[r][q] = fmaf(a[r].y, b[q].y, acc[r][q]);
                acc[r][q] = fmaf(a[r].z, b[q].z, acc[r][q]);
                acc[r][q] = fmaf(a[r].w, b[q].w, acc[r][q]);
            }
    }
}

// acc[r][0..3] += sum_j W[ty + 16 r][j] T[j][4 tx ..],  acc[r][4..7] += ... T[j][64 + 4 tx ..]
template <int RA>
__device__ __forceinline__ void msb_update(const float* __restrict__ W, const float* __restrict__ T, int ty, int tx, float (&acc)[RA][8]) {
#pragma unroll 2
    for (int j = 0; j < MB_T; j += 4) {
        float4 w[RA];
#pragma unroll
        for (int r = 0; r < RA; ++r) w[r] = *reinterpret_cast<const float4*>(W + (ty + 16 * r) * MB_WP + j);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const float4 t0 = *reinterpret_cast<const float4*>(T + (j + jj) * MB_P + 4 * tx);
            const float4 t1 = *reinterpret_cast<const float4*>(T + (j + jj) * MB_P + 64 + 4 * tx);
#pragma unroll
            for (int r = 0; r < RA; ++r) {
                const float wv = jj == 0 ? w[r].x : jj == 1 ? w[r].y : jj == 2 ? w[r].z : w[r].w;
                acc[r][0] = fmaf(wv, t0.x, acc[r][0]); acc[r][1] = fmaf(wv, t0.y, acc[r][1]);
                acc[r][2] = fmaf(wv, t0.z, acc[r][2]); acc[r][3] = fmaf(wv, t0.w, acc[r][3]);
                acc[r][4] = fmaf(wv, t1.x, acc[r][4]); acc[r][5] = fmaf(wv, t1.y, acc[r][5]);
                acc[r][6] = fmaf(wv, t1.z, acc[r][6]); acc[r][7] = fmaf(wv, t1.w, acc[r][7]);
            }
        }
    }
}

// P and, when asked, the factor of dS: exponent clamped below at -75 as guard_exp does (src/guard.py:7-9), zero gradient there
__device__ __forceinline__ float msb_weight(float s, float inv_b2, bool& live) {
    const float e = (s - 1.0f) * inv_b2;
    live = e >= -75.0f;
    return __expf(fmaxf(e, -75.0f));
}

// ---------------------------------------------------------------------------------------------- rows, pass A
template <int R>
__global__ void __launch_bounds__(MB_THREADS) msb_rows_a_kernel(MsbParams p) {
    constexpr int BT = 16 * R;            // rows of this CTA
    extern __shared__ __align__(16) float sm[];
    float* Qs = sm;                       // [BT][132]
    float* Xs = Qs + BT * MB_P;           // [64][132]
    float* Ps = Xs + MB_TILE;             // [BT][68]
    float* red = Ps + BT * MB_WP;         // [BT][16]
    const int b = blockIdx.y, row0 = blockIdx.x * BT;
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    const long long off = (long long)b * p.N * p.d;
    const float bwv = p.bw[b], inv_b2 = 1.0f / (bwv * bwv);
    msb_load(Qs, p.Q + off, row0, p.N, p.d, BT);
    float O[R][8];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) O[r][c] = 0.f;
    for (int k0 = 0; k0 < p.N; k0 += MB_T) {
        __syncthreads();                                      // previous tile's readers are done (and Qs is loaded)
        msb_load(Xs, p.X + off, k0, p.N, p.d);
        __syncthreads();
        float s[R][4];
        msb_dot(Qs, Xs, ty, tx, s);
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                bool live;
                const float w = msb_weight(s[r][q], inv_b2, live);
                Ps[(ty + 16 * r) * MB_WP + tx + 16 * q] = (k0 + tx + 16 * q < p.N) ? w : 0.f;
            }
        __syncthreads();
        msb_update(Ps, Xs, ty, tx, O);
    }
    // |O| and Q' . G per row: sums over the 16 threads that share a row
    float g[R][8];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int row = row0 + ty + 16 * r;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int c = 64 * h + 4 * tx;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < p.N && c < p.d) v = __ldg(reinterpret_cast<const float4*>(p.G + off + (long long)row * p.d + c));
            g[r][4 * h] = v.x; g[r][4 * h + 1] = v.y; g[r][4 * h + 2] = v.z; g[r][4 * h + 3] = v.w;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < R; ++r) {
        float ss = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) ss = fmaf(O[r][c], O[r][c], ss);
        red[(ty + 16 * r) * 16 + tx] = ss;
    }
    __syncthreads();
    float rn[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) ss += red[(ty + 16 * r) * 16 + i];
        rn[r] = ss > 0.f ? rsqrtf(ss) : 0.f;                  // a dead row (every weight underflowed) passes no gradient
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < R; ++r) {
        float dt = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) dt = fmaf(O[r][c] * rn[r], g[r][c], dt);
        red[(ty + 16 * r) * 16 + tx] = dt;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < R; ++r) {
        float dt = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) dt += red[(ty + 16 * r) * 16 + i];
        const int row = row0 + ty + 16 * r;
        if (row < p.N) {
            float* o = p.dO + ((long long)b * p.N + row) * MB_D;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float4 v;
                v.x = (g[r][4 * h] - O[r][4 * h] * rn[r] * dt) * rn[r];
                v.y = (g[r][4 * h + 1] - O[r][4 * h + 1] * rn[r] * dt) * rn[r];
                v.z = (g[r][4 * h + 2] - O[r][4 * h + 2] * rn[r] * dt) * rn[r];
                v.w = (g[r][4 * h + 3] - O[r][4 * h + 3] * rn[r] * dt) * rn[r];
                *reinterpret_cast<float4*>(o + 64 * h + 4 * tx) = v;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------- rows, pass B
template <int R>
__global__ void __launch_bounds__(MB_THREADS) msb_rows_b_kernel(MsbParams p) {
    constexpr int BT = 16 * R;
    extern __shared__ __align__(16) float sm[];
    float* Qs = sm;                       // [BT][132]
    float* Ds = Qs + BT * MB_P;           // dO rows of this block
    float* Xs = Ds + BT * MB_P;           // [64][132]
    float* Ws = Xs + MB_TILE;             // dS tile [BT][68]
    const int b = blockIdx.y, row0 = blockIdx.x * BT;
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    const long long off = (long long)b * p.N * p.d;
    const float bwv = p.bw[b], inv_b2 = 1.0f / (bwv * bwv);
    msb_load(Qs, p.Q + off, row0, p.N, p.d, BT);
    msb_load(Ds, p.dO + (long long)b * p.N * MB_D, row0, p.N, MB_D, BT);
    float dQ[R][8];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) dQ[r][c] = 0.f;
    for (int k0 = 0; k0 < p.N; k0 += MB_T) {
        __syncthreads();
        msb_load(Xs, p.X + off, k0, p.N, p.d);
        __syncthreads();
        float s[R][4], dp[R][4];
        msb_dot(Qs, Xs, ty, tx, s);
        msb_dot(Ds, Xs, ty, tx, dp);
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                bool live;
                const float w = msb_weight(s[r][q], inv_b2, live);
                Ws[(ty + 16 * r) * MB_WP + tx + 16 * q] = (live && k0 + tx + 16 * q < p.N) ? w * dp[r][q] * inv_b2 : 0.f;
            }
        __syncthreads();
        msb_update(Ws, Xs, ty, tx, dQ);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int row = row0 + ty + 16 * r;
        if (row >= p.N) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int c = 64 * h + 4 * tx;
            if (c < p.d)
                *reinterpret_cast<float4*>(p.dQ + off + (long long)row * p.d + c) =
                    make_float4(dQ[r][4 * h], dQ[r][4 * h + 1], dQ[r][4 * h + 2], dQ[r][4 * h + 3]);
        }
    }
}

// ---------------------------------------------------------------------------------------------- columns (keys)
template <int R>
__global__ void __launch_bounds__(MB_THREADS) msb_cols_kernel(MsbParams p) {
    constexpr int BT = 16 * R;            // keys of this CTA
    extern __shared__ __align__(16) float sm[];
    float* Xs = sm;                       // this block's keys [BT][132]
    float* Qs = Xs + BT * MB_P;           // streamed rows [64][132]
    float* Ds = Qs + MB_TILE;             // their dO
    float* Pt = Ds + MB_TILE;             // P^T  [key][row]  [BT][68]
    float* St = Pt + BT * MB_WP;          // dS^T [key][row]
    const int b = blockIdx.y, key0 = blockIdx.x * BT;
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    const long long off = (long long)b * p.N * p.d;
    const float bwv = p.bw[b], inv_b2 = 1.0f / (bwv * bwv);
    msb_load(Xs, p.X + off, key0, p.N, p.d, BT);
    float dX[R][8];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) dX[r][c] = 0.f;
    for (int r0 = 0; r0 < p.N; r0 += MB_T) {
        __syncthreads();
        msb_load(Qs, p.Q + off, r0, p.N, p.d);
        msb_load(Ds, p.dO + (long long)b * p.N * MB_D, r0, p.N, MB_D);
        __syncthreads();
        float s[R][4], dp[R][4];
        msb_dot(Xs, Qs, ty, tx, s);                          // s[r][q] = X[key ty + 16 r] . Q[row tx + 16 q]
        msb_dot(Xs, Ds, ty, tx, dp);
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                bool live;
                const float w = msb_weight(s[r][q], inv_b2, live);
                const bool in = (r0 + tx + 16 * q < p.N);    // rows beyond the cloud carry no weight (their dO tile is zero anyway)
                Pt[(ty + 16 * r) * MB_WP + tx + 16 * q] = in ? w : 0.f;
                St[(ty + 16 * r) * MB_WP + tx + 16 * q] = (in && live) ? w * dp[r][q] * inv_b2 : 0.f;
            }
        __syncthreads();
        msb_update(Pt, Ds, ty, tx, dX);
        msb_update(St, Qs, ty, tx, dX);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int key = key0 + ty + 16 * r;
        if (key >= p.N) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int c = 64 * h + 4 * tx;
            if (c < p.d) {
                float4* o = reinterpret_cast<float4*>(p.dX + off + (long long)key * p.d + c);
                float4 v = *o;
                v.x += dX[r][4 * h]; v.y += dX[r][4 * h + 1]; v.z += dX[r][4 * h + 2]; v.w += dX[r][4 * h + 3];
                *o = v;
            }
        }
    }
}

}  // namespace sed

using namespace sed;

extern "C" {

int64_t sed_ms_shift_backward_workspace_bytes(int B, int N) { return (int64_t)B * N * MB_D * (int64_t)sizeof(float); }

int sed_ms_shift_backward_step(const float* Q, const float* X, const float* G, const float* bw, int B, int N, int d,
                               float* dQ, float* dX_accum, void* workspace, sed_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!Q || !X || !G || !bw || !dQ || !dX_accum || !workspace || B <= 0 || N <= 0 || d <= 0) return SED_ERR_ARG;
    if (d > MB_D || (d & 3)) return SED_ERR_UNSUPPORTED;
    MsbParams p{Q, X, G, bw, (float*)workspace, dQ, dX_accum, N, d};
    // rows (keys) per CTA = 16 R: the R that needs the fewest CTA-rows of work in whole waves (one CTA per SM).  At
    // N = 10 000 the natural 64 gives 157 CTAs = two waves on 148 SMs, 80 gives 125 = one.
    int best = 4;
    long long best_cost = -1;
    for (int R : {2, 3, 4, 5, 6, 8}) {
        const long long ctas = (long long)((N + 16 * R - 1) / (16 * R)) * B;
        const long long cost = ((ctas + kNumSMs - 1) / kNumSMs) * R;
        if (best_cost < 0 || cost < best_cost) { best = R; best_cost = cost; }
    }
    int rc = SED_OK;
    auto go = [&](auto tag) -> int {
        constexpr int R = decltype(tag)::value;
        constexpr int BT = 16 * R;
        const dim3 grid((N + BT - 1) / BT, B);
        const size_t sm_a = (size_t)(BT * MB_P + MB_TILE + BT * MB_WP + BT * 16) * sizeof(float);
        const size_t sm_b = (size_t)(2 * BT * MB_P + MB_TILE + BT * MB_WP) * sizeof(float);
        const size_t sm_c = (size_t)(BT * MB_P + 2 * MB_TILE + 2 * BT * MB_WP) * sizeof(float);
        static_assert((size_t)(8 * 16 * MB_P + 2 * MB_TILE + 2 * 8 * 16 * MB_WP) * sizeof(float) <= 227 * 1024, "shared memory budget");
        SED_CUDA(cudaFuncSetAttribute(msb_rows_a_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_a));
        SED_CUDA(cudaFuncSetAttribute(msb_rows_b_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_b));
        SED_CUDA(cudaFuncSetAttribute(msb_cols_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_c));
        msb_rows_a_kernel<R><<<grid, MB_THREADS, sm_a, st>>>(p);
        SED_CHECK_LAUNCH();
        msb_rows_b_kernel<R><<<grid, MB_THREADS, sm_b, st>>>(p);
        SED_CHECK_LAUNCH();
        msb_cols_kernel<R><<<grid, MB_THREADS, sm_c, st>>>(p);
        SED_CHECK_LAUNCH();
        return SED_OK;
    };
    switch (best) {
        case 2: rc = go(std::integral_constant<int, 2>{}); break;
        case 3: rc = go(std::integral_constant<int, 3>{}); break;
        case 5: rc = go(std::integral_constant<int, 5>{}); break;
        case 6: rc = go(std::integral_constant<int, 6>{}); break;
        case 8: rc = go(std::integral_constant<int, 8>{}); break;
        default: rc = go(std::integral_constant<int, 4>{}); break;
    }
    return rc;
}

}  // extern "C"
