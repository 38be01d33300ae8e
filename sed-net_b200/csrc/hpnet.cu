// compute_entropy of the reference's HPNet-style embedding weighting (src/smooth_normal_matrix.py:95-154):
//     interval_k = max_{i,j} (f_ik - f_jk) - min_{i,j} (f_ik - f_jk)              over the first M = 5 * CHUNK points
//     dst_ij     = || (f_i - f_j) / interval ||_2
//     alpha      = -log(0.5) / (sum_ij dst_ij / N^2)
//     E          = sum_ij [ -s log(s + eps) - (1 - s) log(1 - s + eps) ] / N^2 ,   s = exp(-alpha dst_ij),  eps = 1e-7
// The reference materialises (CHUNK x CHUNK x K) difference tensors three times over 25 chunk pairs (512 MB each at
// CHUNK = 1000, K = 128).  Here one CTA owns a 64 x 64 tile of pairs, the two row blocks stream through shared memory
// in 32-channel slices, every thread keeps a 4 x 4 block of squared distances in registers (direct-form FP32
// differences, as the reference), and the tile sums leave the CTA as FP64 partials that a second kernel adds in a fixed
// order.  Nothing of size M x M is written.  The max / min of the pairwise differences per channel are
// (max_i f_ik - min_i f_ik) and its negative exactly (rounding is monotone), so the interval is 2 (max - min).
#include "internal.h"

namespace sed {

constexpr int CE_T = 64, CE_KC = 32, CE_THREADS = 256;

__global__ void __launch_bounds__(256) ce_interval_kernel(const float* __restrict__ f, int M, int K, float* __restrict__ inv_interval) {
    __shared__ float smx[256], smn[256];
    const int k = blockIdx.x;
    float mx = -INFINITY, mn = INFINITY;
    for (int i = threadIdx.x; i < M; i += 256) {
        const float v = f[(long long)i * K + k];
        mx = fmaxf(mx, v); mn = fminf(mn, v);
    }
    smx[threadIdx.x] = mx; smn[threadIdx.x] = mn;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            smx[threadIdx.x] = fmaxf(smx[threadIdx.x], smx[threadIdx.x + s]);
            smn[threadIdx.x] = fminf(smn[threadIdx.x], smn[threadIdx.x + s]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float a = __fsub_rn(smx[0], smn[0]);            // max of the differences; the min is -a
        inv_interval[k] = __fdiv_rn(1.0f, __fsub_rn(a, -a));
    }
}

// MODE 0: sum of dst;  MODE 1: sum of the binary entropies of exp(-alpha dst), alpha read from scal[1]
template <int MODE>
__global__ void __launch_bounds__(CE_THREADS) ce_pairs_kernel(const float* __restrict__ f, int M, int K,
                                                              const float* __restrict__ inv_interval,
                                                              const float* __restrict__ scal, double* __restrict__ partial) {
    __shared__ float A[CE_KC][CE_T + 4], Bt[CE_KC][CE_T + 4], inv[CE_KC];
    __shared__ double red[CE_THREADS / 32];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int i0 = blockIdx.y * CE_T, j0 = blockIdx.x * CE_T;
    float d2[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) d2[a][b] = 0.f;
    for (int k0 = 0; k0 < K; k0 += CE_KC) {
        __syncthreads();
        for (int e = tid; e < CE_KC * CE_T; e += CE_THREADS) {
            const int r = e / CE_KC, c = e % CE_KC;           // consecutive threads: consecutive channels of one row
            const bool kv = k0 + c < K;
            A[c][r] = (kv && i0 + r < M) ? f[(long long)(i0 + r) * K + k0 + c] : 0.f;
            Bt[c][r] = (kv && j0 + r < M) ? f[(long long)(j0 + r) * K + k0 + c] : 0.f;
        }
        if (tid < CE_KC) inv[tid] = (k0 + tid < K) ? inv_interval[k0 + tid] : 0.f;
        __syncthreads();
        const int kc = min(CE_KC, K - k0);
        for (int c = 0; c < kc; ++c) {
            const float4 a4 = *reinterpret_cast<const float4*>(&A[c][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&Bt[c][tx * 4]);
            const float w = inv[c];
            const float aa[4] = {a4.x, a4.y, a4.z, a4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const float t = __fmul_rn(__fsub_rn(aa[a], bb[b]), w);
                    d2[a][b] = fmaf(t, t, d2[a][b]);
                }
        }
    }
    const float alpha = MODE == 1 ? scal[1] : 0.f;
    double acc = 0.0;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            if (i0 + ty * 4 + a >= M || j0 + tx * 4 + b >= M) continue;
            const float dst = sqrtf(d2[a][b]);
            if (MODE == 0) {
                acc += (double)dst;
            } else {
                const float s = expf(-alpha * dst);
                const float ent = -s * logf(s + 1e-7f) - (1.0f - s) * logf(1.0f - s + 1e-7f);
                acc += (double)ent;
            }
        }
    acc = warp_sum_d(acc);
    if ((tid & 31) == 0) red[tid >> 5] = acc;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < CE_THREADS / 32; ++w) t += red[w];
        partial[(long long)blockIdx.y * gridDim.x + blockIdx.x] = t;
    }
}

// fixed-order sum of the tile partials; writes scal[0] = sum / N^2 and (MODE 0) scal[1] = alpha = -log(0.5) / scal[0]
__global__ void __launch_bounds__(1024) ce_finish_kernel(const double* __restrict__ partial, int n, double nn, int mode,
                                                         float* __restrict__ scal, float* __restrict__ out) {
    __shared__ double sh[1024];
    double t = 0.0;
    for (int i = threadIdx.x; i < n; i += 1024) t += partial[i];
    sh[threadIdx.x] = t;
    __syncthreads();
    for (int s = 512; s > 0; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float v = (float)(sh[0] / nn);
        if (mode == 0) {
            scal[0] = v;
            scal[1] = __fdiv_rn(0.6931471805599453f, v);      // -np.log(0.5) / average_dst in FP32 (:132-134)
            if (out) { out[1] = v; out[2] = scal[1]; }
        } else if (out) {
            out[0] = v;
        }
    }
}

}  // namespace sed

using namespace sed;

extern "C" {

int sed_compute_entropy(const float* features, int N, int K, int chunk, float* out3, sed_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!features || !out3 || N <= 0 || K <= 0 || chunk <= 0) return SED_ERR_ARG;
    const int M = (int)std::min<long long>(N, 5LL * chunk);   // ITER = 5 chunks of CHUNK rows (:106-110)
    const int T = (M + CE_T - 1) / CE_T;
    ensure_pool_config();
    char* ws = nullptr;
    const size_t off_part = ((size_t)K * 4 + 8 + 255) / 256 * 256;
    SED_CUDA(cudaMallocAsync((void**)&ws, off_part + (size_t)T * T * sizeof(double), st));
    float* inv = reinterpret_cast<float*>(ws);
    float* scal = inv + K;
    double* partial = reinterpret_cast<double*>(ws + off_part);
    const double nn = (double)N * (double)N;
    ce_interval_kernel<<<K, 256, 0, st>>>(features, M, K, inv);
    ce_pairs_kernel<0><<<dim3(T, T), CE_THREADS, 0, st>>>(features, M, K, inv, scal, partial);
    ce_finish_kernel<<<1, 1024, 0, st>>>(partial, T * T, nn, 0, scal, out3);
    ce_pairs_kernel<1><<<dim3(T, T), CE_THREADS, 0, st>>>(features, M, K, inv, scal, partial);
    ce_finish_kernel<<<1, 1024, 0, st>>>(partial, T * T, nn, 1, scal, out3);
    const cudaError_t e = cudaGetLastError();
    g_sed_launches += 5;
    cudaFreeAsync(ws, st);
    return e == cudaSuccess ? SED_OK : SED_ERR_CUDA_BASE - (int)e;
}

}  // extern "C"
