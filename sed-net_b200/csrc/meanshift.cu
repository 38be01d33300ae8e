// Mean-shift clustering on the unit hypersphere (reference src/mean_shift.py:19-179, src/guard.py).
//
//   * ms_shift_ffma_kernel   one fixed-point iteration new_X <- normalize(K X / sum K), K = exp(-(2 - 2 q.x)/b^2/2),
//                            as a flash-style pass: a CTA owns 64 query rows, streams the N keys through shared
//                            memory, never materialises the N x N kernel matrix (FP32 FFMA, reference op order);
//   * nms_*                  src/mean_shift.py:139-179 without the host round trip (np.unique -> device histogram
//                            + ordered compaction).
// The tcgen05 (tensor-core) shift kernel lives in meanshift_tc.cu.
#include "internal.h"

namespace sed {

// ------------------------------------------------------------------------------------------------ normalise
// X[b, n, :] = emb[b, :, n] / max(||emb[b, :, n]||, 1e-12)   (F.normalize, generate_predictions_aug.py:379-380)
__global__ void __launch_bounds__(256) normalize_transpose_kernel(const float* __restrict__ emb, int d, int N,
                                                                  float* __restrict__ X) {
    extern __shared__ float tile[];  // [d][33]
    const int b = blockIdx.y, n0 = blockIdx.x * 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* eb = emb + (long long)b * d * N;
    for (int c = warp; c < d; c += 8) tile[c * 33 + lane] = (n0 + lane < N) ? eb[(long long)c * N + n0 + lane] : 0.f;
    __syncthreads();
    for (int r = warp; r < 32; r += 8) {
        const int n = n0 + r;
        if (n >= N) continue;
        float ss = 0.f;
        for (int c = lane; c < d; c += 32) { float v = tile[c * 33 + r]; ss = fmaf(v, v, ss); }
        ss = warp_sum_f(ss);
        const float den = fmaxf(sqrtf(ss), 1e-12f);
        for (int c = lane; c < d; c += 32) X[((long long)b * N + n) * d + c] = tile[c * 33 + r] / den;
    }
}

// XT[b, c, n] = X[b, n, c]  (channel-major copy of the keys, constant over the iterations)
__global__ void __launch_bounds__(256) transpose_rows_kernel(const float* __restrict__ X, int N, int d,
                                                             float* __restrict__ XT) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = warp; r < 32; r += 8)
        tile[r][lane] = (n0 + r < N && c0 + lane < d) ? X[((long long)b * N + n0 + r) * d + c0 + lane] : 0.f;
    __syncthreads();
    for (int r = warp; r < 32; r += 8)
        if (c0 + r < d && n0 + lane < N) XT[((long long)b * d + c0 + r) * N + n0 + lane] = tile[lane][r];
}

// ------------------------------------------------------------------------------------------------ shift (FFMA)
constexpr int MS_TQ = 64;   // query rows per CTA
constexpr int MS_TK = 64;   // keys per tile
constexpr int MS_DMAX = 256;   // widest embedding of the FFMA path (rows padded to 128 or 256 channels)
constexpr int MS_THREADS = 256;

struct ShiftParams {
    const float* X;    // (B,N,d) keys / values
    const float* XT;   // (B,d,N) channel-major copy of X
    const float* Q;    // (B,N,d) current positions
    const float* bw;   // (B)
    float* out;        // (B,N,d)
    int N, d, kernel_type;
};

__device__ __forceinline__ float ms_kernel_weight(float dot, float b2, int kernel_type) {
    // src/mean_shift.py:60-68: dist = 2 - 2 q.x ; gaussian exp(clamp(-dist/b^2/2, -75, 75)) ; epa relu(3/4 (1 - dist/b^2))
    const float dist = __fsub_rn(2.0f, __fmul_rn(2.0f, dot));
    if (kernel_type == 0) {
        float t = __fmul_rn(__fdiv_rn(-dist, b2), 0.5f);
        t = fminf(fmaxf(t, -75.f), 75.f);
        return expf(t);
    }
    return fmaxf(__fmul_rn(0.75f, __fsub_rn(1.0f, __fdiv_rn(dist, b2))), 0.f);
}

// D: padded channel count (128, or 256 for embeddings wider than 128: the hpnet spectral embedding has 148 columns)
template <int D>
__global__ void __launch_bounds__(MS_THREADS, 1) ms_shift_ffma_kernel(ShiftParams p) {
    constexpr int G = D / 64;   // float4 channel groups per thread: channels g * 64 + tx * 4 + [0,4)
    extern __shared__ __align__(16) float sm[];
    float* Qt = sm;                         // [D][64]     Qt[c][q]
    float* Xt = Qt + D * MS_TQ;          // [D][64]     Xt[c][key]
    float* Xr = Xt + D * MS_TK;          // [64][D]     Xr[key][c]
    float* Ps = Xr + MS_TK * D;          // [64][64]    Ps[key][q], q-quads swizzled by key

    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int b = blockIdx.y, q0 = blockIdx.x * MS_TQ;
    const int N = p.N, d = p.d;
    const float* X = p.X + (long long)b * N * d;
    const float* XT = p.XT + (long long)b * d * N;
    const float* Q = p.Q + (long long)b * N * d;
    const float bwv = p.bw[b];
    const float b2 = __fmul_rn(bwv, bwv);

    // ---- stage the query tile transposed: thread -> (row, channel quad)
    for (int e = tid; e < MS_TQ * (D / 4); e += MS_THREADS) {
        const int r = e & 63, c4 = e >> 6;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q0 + r < N && c4 * 4 < d) v = __ldg(reinterpret_cast<const float4*>(Q + (long long)(q0 + r) * d + c4 * 4));
        Qt[(c4 * 4 + 0) * MS_TQ + r] = v.x;
        Qt[(c4 * 4 + 1) * MS_TQ + r] = v.y;
        Qt[(c4 * 4 + 2) * MS_TQ + r] = v.z;
        Qt[(c4 * 4 + 3) * MS_TQ + r] = v.w;
    }

    float o[4][4 * G];
    float l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        l[i] = 0.f;
#pragma unroll
        for (int e = 0; e < 4 * G; ++e) o[i][e] = 0.f;
    }

    const int ntiles = (N + MS_TK - 1) / MS_TK;
    for (int t = 0; t < ntiles; ++t) {
        const int j0 = t * MS_TK;
        __syncthreads();  // previous tile fully consumed (and Qt visible on the first pass)
        // ---- load the key tile in both layouts (coalesced from X and XT)
        for (int e = tid; e < MS_TK * (D / 4); e += MS_THREADS) {
            const int r = e / (D / 4), c4 = e % (D / 4);   // row-major: D / 4 quads per row
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j0 + r < N && c4 * 4 < d) v = __ldg(reinterpret_cast<const float4*>(X + (long long)(j0 + r) * d + c4 * 4));
            *reinterpret_cast<float4*>(Xr + r * D + c4 * 4) = v;
        }
        for (int e = tid; e < D * (MS_TK / 4); e += MS_THREADS) {
            const int c = e >> 4, k4 = e & 15;   // channel-major: 16 quads per channel row
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < d) {
                const float* src = XT + (long long)c * N + j0 + k4 * 4;
                if (j0 + k4 * 4 + 3 < N && ((N & 3) == 0)) v = __ldg(reinterpret_cast<const float4*>(src));
                else {
                    if (j0 + k4 * 4 + 0 < N) v.x = __ldg(src + 0);
                    if (j0 + k4 * 4 + 1 < N) v.y = __ldg(src + 1);
                    if (j0 + k4 * 4 + 2 < N) v.z = __ldg(src + 2);
                    if (j0 + k4 * 4 + 3 < N) v.w = __ldg(src + 3);
                }
            }
            *reinterpret_cast<float4*>(Xt + c * MS_TK + k4 * 4) = v;
        }
        __syncthreads();

        // ---- S = Q . X^T for this thread's 4 queries x 4 keys
        float s[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int e = 0; e < 4; ++e) s[i][e] = 0.f;
#pragma unroll 8
        for (int c = 0; c < D; ++c) {
            const float4 a = *reinterpret_cast<const float4*>(Qt + c * MS_TQ + ty * 4);
            const float4 k4 = *reinterpret_cast<const float4*>(Xt + c * MS_TK + tx * 4);
            const float aa[4] = {a.x, a.y, a.z, a.w};
            const float kk[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int e = 0; e < 4; ++e) s[i][e] = fmaf(aa[i], kk[e], s[i][e]);
        }
        // ---- kernel weights; row sums; P^T to shared memory
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int key = tx * 4 + e;
            const bool ok = (j0 + key) < N;
            float pv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                pv[i] = ok ? ms_kernel_weight(s[i][e], b2, p.kernel_type) : 0.f;
                l[i] += pv[i];
            }
            const int slot = ty ^ ((key >> 2) & 15);
            *reinterpret_cast<float4*>(Ps + key * MS_TQ + slot * 4) = make_float4(pv[0], pv[1], pv[2], pv[3]);
        }
        __syncthreads();
        // ---- O += P . X : 4 queries x 4 G channels per thread
#pragma unroll 4
        for (int key = 0; key < MS_TK; ++key) {
            const int slot = ty ^ ((key >> 2) & 15);
            const float4 pq = *reinterpret_cast<const float4*>(Ps + key * MS_TQ + slot * 4);
            const float pp[4] = {pq.x, pq.y, pq.z, pq.w};
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const float4 xg = *reinterpret_cast<const float4*>(Xr + key * D + g * 64 + tx * 4);
                const float xx[4] = {xg.x, xg.y, xg.z, xg.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int e = 0; e < 4; ++e) o[i][g * 4 + e] = fmaf(pp[i], xx[e], o[i][g * 4 + e]);
            }
        }
    }

    // ---- epilogue: D = 1 / sum K ; new = q + ((K X) D - q) ; new /= ||new||   (src/mean_shift.py:70-77)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float sum = l[i];
#pragma unroll
        for (int m = 8; m > 0; m >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, m);
        const float D = __fdiv_rn(1.0f, sum);
        const int ql = ty * 4 + i;
        float z[4 * G];
        float nn = 0.f;
#pragma unroll
        for (int e = 0; e < 4 * G; ++e) {
            const int c = (e >> 2) * 64 + tx * 4 + (e & 3);
            const float qv = Qt[c * MS_TQ + ql];
            const float M = __fsub_rn(__fmul_rn(o[i][e], D), qv);
            z[e] = __fadd_rn(qv, M);
            nn = fmaf(z[e], z[e], nn);
        }
#pragma unroll
        for (int m = 8; m > 0; m >>= 1) nn += __shfl_xor_sync(0xffffffffu, nn, m);
        const float nrm = sqrtf(nn);
        const int q = q0 + ql;
        if (q < N) {
            float* dst = p.out + ((long long)b * N + q) * d;
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const int c = g * 64 + tx * 4;
                if (c < d)
                    *reinterpret_cast<float4*>(dst + c) = make_float4(z[g * 4] / nrm, z[g * 4 + 1] / nrm, z[g * 4 + 2] / nrm, z[g * 4 + 3] / nrm);
            }
        }
    }
}

int ms_shift_ffma(const float* X, const float* XT, const float* Q, const float* bw, int B, int N, int d,
                  int kernel_type, float* out, cudaStream_t st) {
    ShiftParams p{X, XT, Q, bw, out, N, d, kernel_type};
    const int D = d <= 128 ? 128 : (d <= 192 ? 192 : 256);
    const size_t smem = (size_t)(D * MS_TQ + D * MS_TK + MS_TK * D + MS_TK * MS_TQ) * sizeof(float);
    dim3 grid((N + MS_TQ - 1) / MS_TQ, B);
    if (D == 128) {
        SED_CUDA(cudaFuncSetAttribute(ms_shift_ffma_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ms_shift_ffma_kernel<128><<<grid, MS_THREADS, smem, st>>>(p);
    } else if (D == 192) {   // the 148-column hpnet embedding
        SED_CUDA(cudaFuncSetAttribute(ms_shift_ffma_kernel<192>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ms_shift_ffma_kernel<192><<<grid, MS_THREADS, smem, st>>>(p);
    } else {
        SED_CUDA(cudaFuncSetAttribute(ms_shift_ffma_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ms_shift_ffma_kernel<256><<<grid, MS_THREADS, smem, st>>>(p);
    }
    SED_CHECK_LAUNCH();
    return SED_OK;
}

// ------------------------------------------------------------------------------------------------ nms
// Block-wide exclusive scan of one int per thread (1024 threads). Returns the exclusive prefix; total in *total.
__device__ __forceinline__ int block_excl_scan_1024(int v, int* sh /*[33]*/, int* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) sh[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = sh[lane];
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        sh[lane] = winc - w;
        if (lane == 31) sh[32] = winc;
    }
    __syncthreads();
    const int res = sh[warp] + inc - v;
    *total = sh[32];
    __syncthreads();
    return res;
}

__global__ void histogram_kernel(const int* __restrict__ memb, int N, int* __restrict__ counts) {
    const int b = blockIdx.y, n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < N) atomicAdd(&counts[(long long)b * N + memb[(long long)b * N + n]], 1);
}

// Ascending compaction of {i : flag[b,i] > 0} into list[b, :], count into n_out[b]. One CTA of 1024 per cloud.
// With cap > 0, a count above cap stores -1 (and the list is truncated).
__global__ void __launch_bounds__(1024) compact_kernel(const int* __restrict__ flag, int N, int* __restrict__ list,
                                                       int list_stride, int cap, int* __restrict__ n_out) {
    __shared__ int sh[33];
    const int b = blockIdx.x;
    int base = 0;
    for (int i0 = 0; i0 < N; i0 += 1024) {
        const int i = i0 + threadIdx.x;
        const int f = (i < N && flag[(long long)b * N + i] > 0) ? 1 : 0;
        int total;
        const int pos = base + block_excl_scan_1024(f, sh, &total);
        if (f && (cap <= 0 || pos < cap)) list[(long long)b * list_stride + pos] = i;
        base += total;
    }
    if (threadIdx.x == 0) n_out[b] = (cap > 0 && base > cap) ? -1 : base;
}

// src/mean_shift.py:164-171: for every occupied centre u, the neighbour (dist < b, note b not b^2) with the most
// members; first index on ties, index 0 if every product is zero.  One warp per u; flags the winners.
__global__ void __launch_bounds__(256) nms_vote_kernel(const float* __restrict__ centers, const float* __restrict__ bw,
                                                       const int* __restrict__ uniq, const int* __restrict__ n_uniq,
                                                       const int* __restrict__ counts, int N, int d,
                                                       int* __restrict__ winner_flag) {
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const int U = n_uniq[b];
    const int ui = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (ui >= U) return;
    const float* Cb = centers + (long long)b * N * d;
    const int* ub = uniq + (long long)b * N;
    const int* cb = counts + (long long)b * N;
    const int u = ub[ui];
    const float bwv = bw[b];
    float4 cu[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};   // d <= 256: two quads per lane
#pragma unroll
    for (int h = 0; h < 2; ++h)
        if (h * 128 + lane * 4 < d) cu[h] = __ldg(reinterpret_cast<const float4*>(Cb + (long long)u * d + h * 128 + lane * 4));
    int best_v = 0, best_j = 0;
    for (int t = 0; t < U; ++t) {
        const int j = ub[t];
        float dot = 0.f;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float4 cj = make_float4(0.f, 0.f, 0.f, 0.f);
            if (h * 128 + lane * 4 < d) cj = __ldg(reinterpret_cast<const float4*>(Cb + (long long)j * d + h * 128 + lane * 4));
            dot += fmaf(cu[h].x, cj.x, fmaf(cu[h].y, cj.y, fmaf(cu[h].z, cj.z, cu[h].w * cj.w)));
        }
        dot = warp_sum_f(dot);
        const float dist = __fsub_rn(2.0f, __fmul_rn(2.0f, dot));
        const int v = (dist < bwv) ? cb[j] : 0;
        if (v > best_v) { best_v = v; best_j = j; }
    }
    if (lane == 0) winner_flag[(long long)b * N + best_j] = 1;
}

__global__ void gather_rows_kernel(const float* __restrict__ src, const int* __restrict__ ids,
                                   const int* __restrict__ n_ids, int N, int d, int max_rows, float* __restrict__ dst) {
    const int b = blockIdx.y, r = blockIdx.x;
    const int n = n_ids[b];
    float* o = dst + ((long long)b * max_rows + r) * d;
    if (r < n) {
        const float* s = src + ((long long)b * N + ids[(long long)b * max_rows + r]) * d;
        for (int c = threadIdx.x; c < d; c += blockDim.x) o[c] = s[c];
    } else {
        for (int c = threadIdx.x; c < d; c += blockDim.x) o[c] = 0.f;
    }
}

// number of distinct labels in use per cloud (= torch.unique(labels).shape[0])
__global__ void __launch_bounds__(1024) count_labels_kernel(const long long* __restrict__ labels, int N, int max_c,
                                                            int* __restrict__ n_labels) {
    extern __shared__ int used[];
    const int b = blockIdx.x;
    for (int i = threadIdx.x; i < max_c; i += blockDim.x) used[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const long long l = labels[(long long)b * N + i];
        if (l >= 0 && l < max_c) used[l] = 1;
    }
    __syncthreads();
    int c = 0;
    for (int i = threadIdx.x; i < max_c; i += blockDim.x) c += used[i];
    __shared__ int tot;
    if (threadIdx.x == 0) tot = 0;
    __syncthreads();
    if (c) atomicAdd(&tot, c);
    __syncthreads();
    if (threadIdx.x == 0) n_labels[b] = tot;
}

struct NmsWs {
    int *memb, *counts, *uniq, *n_uniq, *wflag;
};
static void carve_nms(Arena& A, int B, int N, NmsWs& w) {
    w.memb = A.take<int>((int64_t)B * N);
    w.counts = A.take<int>((int64_t)B * N);
    w.uniq = A.take<int>((int64_t)B * N);
    w.n_uniq = A.take<int>(B);
    w.wflag = A.take<int>((int64_t)B * N);
}

}  // namespace sed

using namespace sed;

namespace sed {
int ms_shift_tc(const float* X, const float* bw, int B, int N, int d, int iterations, int kernel_type, int prec_mode,
                float* out, float* tmp, cudaStream_t st, const float* Q0);
}

extern "C" {

int sed_normalize_transpose(const float* emb, int B, int d, int N, float* X, sed_stream_t stream) {
    if (!emb || !X || B <= 0 || d <= 0 || N <= 0) return SED_ERR_ARG;
    const size_t smem = (size_t)d * 33 * sizeof(float);
    if (smem > 48 * 1024) return SED_ERR_UNSUPPORTED;
    dim3 grid((N + 31) / 32, B);
    normalize_transpose_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(emb, d, N, X);
    SED_CHECK_LAUNCH();
    return SED_OK;
}

int sed_ms_shift(const float* X, const float* bw, int B, int N, int d, int iterations, int kernel_type, int prec_mode,
                 float* out, float* tmp, sed_stream_t stream) {
    return sed_ms_shift_from(X, nullptr, bw, B, N, d, iterations, kernel_type, prec_mode, out, tmp, stream);
}

int sed_ms_shift_from(const float* X, const float* Q0, const float* bw, int B, int N, int d, int iterations, int kernel_type,
                      int prec_mode, float* out, float* tmp, sed_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!X || !bw || !out || !tmp || B <= 0 || N <= 0 || d <= 0 || d > MS_DMAX || (d & 3) || iterations < 0)
        return SED_ERR_ARG;
    if (kernel_type != 0 && kernel_type != 1) return SED_ERR_ARG;
    if (Q0 == X) Q0 = nullptr;
    if (iterations == 0) {
        SED_CUDA(cudaMemcpyAsync(out, Q0 ? Q0 : X, (size_t)B * N * d * sizeof(float), cudaMemcpyDeviceToDevice, st));
        return SED_OK;
    }
    if (prec_mode >= 1 && prec_mode <= 4) {
        const int rc = ms_shift_tc(X, bw, B, N, d, iterations, kernel_type, prec_mode, out, tmp, st, Q0);
        if (rc != SED_ERR_UNSUPPORTED) return rc;
        prec_mode = 0;   // shape outside the tensor-core kernel's range: the FP32 FFMA kernel
    }
    if (prec_mode != 0) return SED_ERR_ARG;
    // XT lives in the second half of tmp's allocation?  No: tmp is exactly (B,N,d); the channel-major copy of X is
    // allocated from the stream-ordered pool (freed after the last iteration is enqueued).
    ensure_pool_config();
    float* XT = nullptr;
    SED_CUDA(cudaMallocAsync((void**)&XT, (size_t)B * N * d * sizeof(float), st));
    dim3 tg((N + 31) / 32, (d + 31) / 32, B);
    transpose_rows_kernel<<<tg, 256, 0, st>>>(X, N, d, XT);
    ++g_sed_launches;
    // ping-pong so that the last iteration lands in `out`
    const float* cur = Q0 ? Q0 : X;
    int rc = SED_OK;
    for (int it = 0; it < iterations && rc == SED_OK; ++it) {
        float* dst = ((iterations - 1 - it) & 1) ? tmp : out;
        rc = ms_shift_ffma(X, XT, cur, bw, B, N, d, kernel_type, dst, st);
        cur = dst;
    }
    cudaFreeAsync(XT, st);
    return rc;
}

int64_t sed_ms_nms_workspace_bytes(int B, int N) {
    Arena A(nullptr, 0);
    NmsWs w;
    carve_nms(A, B, N, w);
    return A.off;
}

int sed_ms_nms(const float* centers, const float* X, const float* bw, int B, int N, int d, int max_centers,
               int64_t* labels, int* center_ids, int* n_centers, int* n_labels, float* centers_out, void* workspace,
               sed_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!centers || !X || !bw || !labels || !center_ids || !n_centers || !centers_out || !workspace) return SED_ERR_ARG;
    if (B <= 0 || N <= 0 || d <= 0 || d > MS_DMAX || (d & 3) || max_centers <= 0) return SED_ERR_ARG;
    Arena A(workspace, sed_ms_nms_workspace_bytes(B, N));
    NmsWs w;
    carve_nms(A, B, N, w);
    // membership[j] = argmin_i dist(centers_i, X_j)   (:146-149)
    SED_TRY(nearest_cos(X, centers, B, N, N, nullptr, d, w.memb, 0, st));
    // member counts per centre (np.unique ... return_counts, :152-161)
    SED_CUDA(cudaMemsetAsync(w.counts, 0, (size_t)B * N * sizeof(int), st));
    SED_CUDA(cudaMemsetAsync(w.wflag, 0, (size_t)B * N * sizeof(int), st));
    histogram_kernel<<<dim3((N + 255) / 256, B), 256, 0, st>>>(w.memb, N, w.counts);
    SED_CHECK_LAUNCH();
    compact_kernel<<<B, 1024, 0, st>>>(w.counts, N, w.uniq, N, 0, w.n_uniq);
    SED_CHECK_LAUNCH();
    // winners (:164-171) and their ordered unique list
    nms_vote_kernel<<<dim3((N + 7) / 8, B), 256, 0, st>>>(centers, bw, w.uniq, w.n_uniq, w.counts, N, d, w.wflag);
    SED_CHECK_LAUNCH();
    compact_kernel<<<B, 1024, 0, st>>>(w.wflag, N, center_ids, max_centers, max_centers, n_centers);
    SED_CHECK_LAUNCH();
    gather_rows_kernel<<<dim3(max_centers, B), 128, 0, st>>>(centers, center_ids, n_centers, N, d, max_centers, centers_out);
    SED_CHECK_LAUNCH();
    // labels = argmax_c centers[ids] . X   (:177-178)
    SED_TRY(nearest_cos(X, centers_out, B, N, max_centers, n_centers, d, labels, 1, st));
    if (n_labels) {
        count_labels_kernel<<<B, 1024, max_centers * sizeof(int), st>>>((const long long*)labels, N, max_centers, n_labels);
        SED_CHECK_LAUNCH();
    }
    return SED_OK;
}

}  // extern "C"
