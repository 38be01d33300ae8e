// Stage-2 neighbourhood maps of the reference's fitting pipeline:
//   three_nn                Fitting_patches_and_edges/pointnet2/_ext_src/src/interpolate_gpu.cu:14-66
//   get_edges_between_insts Fitting_patches_and_edges/proj_2_edge_utils.py:45-60
//   face_face_inter_map     Fitting_patches_and_edges/proj_2_edge_utils.py:63-110
//
// The reference's three_nn runs ONE thread block per cloud, every thread walking all m known points in global memory.
// Here a CTA owns 64 query points, the known points stream through shared memory in 2048-point tiles (coalesced float4
// loads, broadcast reads), and the grid covers the queries of every cloud: 157 CTAs per 10 000-point cloud, one wave
// on 148 SMs.  Same arithmetic (FP32 direct-form squared distance), same update rule (strict '<' in index order, so
// equal distances keep the lowest index).
#include "internal.h"

namespace sed {

constexpr int NN_THREADS = 64;
constexpr int NN_TILE = 2048;

__global__ void __launch_bounds__(NN_THREADS) three_nn_kernel(const float* __restrict__ unknown, const float* __restrict__ known,
                                                              int n, int m, float* __restrict__ dist2, int* __restrict__ idx) {
    __shared__ float tile[NN_TILE * 3];
    const int b = blockIdx.y, j = blockIdx.x * NN_THREADS + threadIdx.x;
    const float* U = unknown + (long long)b * n * 3;
    const float* K = known + (long long)b * m * 3;
    const bool live = j < n;
    const float ux = live ? U[3 * j] : 0.f, uy = live ? U[3 * j + 1] : 0.f, uz = live ? U[3 * j + 2] : 0.f;
    float best1 = INFINITY, best2 = INFINITY, best3 = INFINITY;   // the reference starts from 1e40 (double): same order
    int i1 = 0, i2 = 0, i3 = 0;
    for (int k0 = 0; k0 < m; k0 += NN_TILE) {
        const int cnt = min(NN_TILE, m - k0);
        __syncthreads();
        for (int t = threadIdx.x; t < cnt * 3; t += NN_THREADS) tile[t] = K[(long long)k0 * 3 + t];
        __syncthreads();
#pragma unroll 4
        for (int k = 0; k < cnt; ++k) {
            const float x = tile[3 * k], y = tile[3 * k + 1], z = tile[3 * k + 2];
            const float d = (ux - x) * (ux - x) + (uy - y) * (uy - y) + (uz - z) * (uz - z);
            if (d < best3) {
                if (d < best1) {
                    best3 = best2; i3 = i2; best2 = best1; i2 = i1; best1 = d; i1 = k0 + k;
                } else if (d < best2) {
                    best3 = best2; i3 = i2; best2 = d; i2 = k0 + k;
                } else {
                    best3 = d; i3 = k0 + k;
                }
            }
        }
    }
    if (live) {
        float* D = dist2 + ((long long)b * n + j) * 3;
        int* I = idx + ((long long)b * n + j) * 3;
        // fewer than three known points: the reference leaves 1e40 -> +inf after the float store
        D[0] = best1; D[1] = best2; D[2] = best3;
        I[0] = i1; I[1] = i2; I[2] = i3;
    }
}

// one_nn / two_nn instance differs (proj_2_edge_utils.py:49-60)
__global__ void inst_edges_kernel(const int* __restrict__ idx3, const long long* __restrict__ insts, int n, int strict,
                                  unsigned char* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long a = insts[i];
    const bool one = insts[idx3[3 * i + 1]] != a;
    const bool two = insts[idx3[3 * i + 2]] != a;
    out[i] = (strict ? (one && two) : one) ? 1 : 0;
}

constexpr int FF_DIM = 30;   // the reference's fixed 30 x 30 map (:79)

__global__ void __launch_bounds__(1024) face_face_kernel(const float* __restrict__ pts, const long long* __restrict__ insts,
                                                         const int* __restrict__ idx3, const long long* __restrict__ ids,
                                                         int n_ids, int n, int thresh, unsigned char* __restrict__ mat) {
    __shared__ int cnt[FF_DIM * FF_DIM];
    __shared__ unsigned char valid[FF_DIM], rowany[FF_DIM];
    __shared__ int first_pt[FF_DIM];
    __shared__ unsigned long long best;
    for (int i = threadIdx.x; i < FF_DIM * FF_DIM; i += blockDim.x) cnt[i] = 0;
    if (threadIdx.x < FF_DIM) { valid[threadIdx.x] = 0; rowany[threadIdx.x] = 0; first_pt[threadIdx.x] = 0x7fffffff; }
    __syncthreads();
    for (int i = threadIdx.x; i < n_ids; i += blockDim.x)
        if (ids[i] >= 0 && ids[i] < FF_DIM) valid[ids[i]] = 1;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const long long a = insts[i];
        if (a < 0 || a >= FF_DIM || !valid[a]) continue;
        atomicMin(&first_pt[a], i);
#pragma unroll
        for (int t = 1; t <= 2; ++t) {
            const long long bb = insts[idx3[3 * i + t]];
            if (bb != a && bb >= 0 && bb < FF_DIM) atomicAdd(&cnt[a * FF_DIM + (int)bb], 1);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < FF_DIM * FF_DIM; i += blockDim.x) {
        const unsigned char v = cnt[i] >= thresh ? 1 : 0;
        mat[i] = v;
        if (v) rowany[i / FF_DIM] = 1;
    }
    __syncthreads();
    // an instance with no neighbour: its first point's nearest point of another instance names the neighbour (:95-109)
    for (int a = 0; a < FF_DIM; ++a) {
        if (!valid[a] || rowany[a] || first_pt[a] == 0x7fffffff) continue;   // uniform across the block
        if (threadIdx.x == 0) best = ~0ull;
        __syncthreads();
        const int f = first_pt[a];
        const float px = pts[3 * f], py = pts[3 * f + 1], pz = pts[3 * f + 2];
        unsigned long long mine = ~0ull;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            if (insts[i] == a) continue;
            const float dx = __fsub_rn(pts[3 * i], px), dy = __fsub_rn(pts[3 * i + 1], py), dz = __fsub_rn(pts[3 * i + 2], pz);
            const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            const unsigned long long e = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)i;
            mine = e < mine ? e : mine;
        }
        atomicMin(&best, mine);
        __syncthreads();
        if (threadIdx.x == 0 && best != ~0ull) {
            const long long nb = insts[(int)(best & 0xffffffffull)];
            if (nb >= 0 && nb < FF_DIM) mat[a * FF_DIM + (int)nb] = 1;
        }
        __syncthreads();
    }
}

}  // namespace sed

using namespace sed;

extern "C" {

int sed_three_nn(const float* unknown, const float* known, int B, int n, int m, float* dist2, int* idx, sed_stream_t stream) {
    if (!unknown || !known || !dist2 || !idx || B <= 0 || n <= 0 || m <= 0) return SED_ERR_ARG;
    three_nn_kernel<<<dim3((n + NN_THREADS - 1) / NN_THREADS, B), NN_THREADS, 0, (cudaStream_t)stream>>>(unknown, known, n, m,
                                                                                                        dist2, idx);
    SED_CHECK_LAUNCH();
    return SED_OK;
}

int sed_inst_edges(const int* idx3, const int64_t* insts, int n, int strict, uint8_t* out, sed_stream_t stream) {
    if (!idx3 || !insts || !out || n <= 0) return SED_ERR_ARG;
    inst_edges_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(idx3, (const long long*)insts, n, strict, out);
    SED_CHECK_LAUNCH();
    return SED_OK;
}

int sed_face_face_map(const float* points, const int64_t* insts, const int* idx3, const int64_t* primitive_ids, int n_ids,
                      int n, int nn_num_thresh, uint8_t* mat, sed_stream_t stream) {
    if (!points || !insts || !idx3 || !primitive_ids || !mat || n <= 0 || n_ids < 0) return SED_ERR_ARG;
    face_face_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(points, (const long long*)insts, idx3,
                                                           (const long long*)primitive_ids, n_ids, n, nn_num_thresh, mat);
    SED_CHECK_LAUNCH();
    return SED_OK;
}

}  // extern "C"
