// Stage-2 neighbourhood maps of the reference's fitting pipeline:
//   three_nn                Fitting_patches_and_edges/pointnet2/_ext_src/src/interpolate_gpu.cu:14-66
//   get_edges_between_insts Fitting_patches_and_edges/proj_2_edge_utils.py:45-60
//   face_face_inter_map     Fitting_patches_and_edges/proj_2_edge_utils.py:63-110
//
// The reference's three_nn runs ONE thread block per cloud, every thread walking all m known points in global memory.
// Here a CTA owns 128 query points (two per thread: one LDS.128 of a known point serves both), the known points stream
// through shared memory in 1024-point float4 tiles (broadcast reads), and when the query tiles alone would not fill
// the 148 SMs the known range is split across CTAs and the per-split triples merged by a second small kernel.  Same arithmetic (FP32 direct-form squared distance), same update rule (strict '<' in index order, so
// equal distances keep the lowest index).
#include "internal.h"

namespace sed {

constexpr int NN_THREADS = 128;
constexpr int NN_QPT = 2;                 // queries per thread: every shared-memory read of a known point serves two queries
constexpr int NN_QCTA = NN_THREADS * NN_QPT;
constexpr int NN_TILE = 1024;             // known points per shared-memory tile (float4: one LDS.128 per point)

// strict-'<' insertion of (d, k) into an ascending triple: interpolate_gpu.cu:38-54
__device__ __forceinline__ void nn_insert(float d, int k, float& b1, float& b2, float& b3, int& i1, int& i2, int& i3) {
    if (d < b1) {
        b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = k;
    } else if (d < b2) {
        b3 = b2; i3 = i2; b2 = d; i2 = k;
    } else if (d < b3) {
        b3 = d; i3 = k;
    }
}

// Grid (query tiles, known-range splits, clouds).  With gridDim.y > 1 the triples of each split go to a workspace and
// three_nn_merge_kernel combines them in split order (earlier splits hold the lower indices, so ties keep the lowest).
__global__ void __launch_bounds__(NN_THREADS) three_nn_kernel(const float* __restrict__ unknown, const float* __restrict__ known,
                                                              int n, int m, int chunk, float* __restrict__ dist2,
                                                              int* __restrict__ idx) {
    __shared__ float4 tile[NN_TILE];
    const int b = blockIdx.z, split = blockIdx.y, nsplit = gridDim.y;
    const float* U = unknown + (long long)b * n * 3;
    const float* K = known + (long long)b * m * 3;
    const int kbeg = split * chunk, kend = min(m, kbeg + chunk);
    int q[NN_QPT];
    float ux[NN_QPT], uy[NN_QPT], uz[NN_QPT], b1[NN_QPT], b2[NN_QPT], b3[NN_QPT];
    int i1[NN_QPT], i2[NN_QPT], i3[NN_QPT];
#pragma unroll
    for (int r = 0; r < NN_QPT; ++r) {
        q[r] = blockIdx.x * NN_QCTA + r * NN_THREADS + threadIdx.x;
        const bool live = q[r] < n;
        ux[r] = live ? U[3 * q[r]] : 0.f; uy[r] = live ? U[3 * q[r] + 1] : 0.f; uz[r] = live ? U[3 * q[r] + 2] : 0.f;
        b1[r] = b2[r] = b3[r] = INFINITY;   // the reference starts from 1e40 (double): same order
        i1[r] = i2[r] = i3[r] = 0;
    }
    for (int k0 = kbeg; k0 < kend; k0 += NN_TILE) {
        const int cnt = min(NN_TILE, kend - k0);
        __syncthreads();
        for (int t = threadIdx.x; t < ((cnt + 3) & ~3); t += NN_THREADS) {
            const float* src = K + (long long)(k0 + t) * 3;
            tile[t] = t < cnt ? make_float4(src[0], src[1], src[2], 0.f) : make_float4(INFINITY, INFINITY, INFINITY, 0.f);
        }
        __syncthreads();
        // four known points at a time: their 4 x NN_QPT distances are independent (latency hidden by ILP), and ONE
        // branch decides whether any of them can enter a triple; the rare insertions then run in index order
        for (int k = 0; k < cnt; k += 4) {
            float d[4][NN_QPT];
            bool any = false;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float4 c = tile[k + e];       // the tile is padded with +inf points up to a multiple of 4
#pragma unroll
                for (int r = 0; r < NN_QPT; ++r) {
                    d[e][r] = (ux[r] - c.x) * (ux[r] - c.x) + (uy[r] - c.y) * (uy[r] - c.y) + (uz[r] - c.z) * (uz[r] - c.z);
                    any |= d[e][r] < b3[r];
                }
            }
            if (any) {
#pragma unroll
                for (int e = 0; e < 4; ++e)
#pragma unroll
                    for (int r = 0; r < NN_QPT; ++r) nn_insert(d[e][r], k0 + k + e, b1[r], b2[r], b3[r], i1[r], i2[r], i3[r]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < NN_QPT; ++r) {
        if (q[r] >= n) continue;
        // fewer than three known points: the reference leaves 1e40 -> +inf after the float store
        const long long o = (((long long)b * nsplit + split) * n + q[r]) * 3;
        dist2[o] = b1[r]; dist2[o + 1] = b2[r]; dist2[o + 2] = b3[r];
        idx[o] = i1[r]; idx[o + 1] = i2[r]; idx[o + 2] = i3[r];
    }
}

__global__ void three_nn_merge_kernel(const float* __restrict__ pd, const int* __restrict__ pi, int n, int nsplit,
                                      float* __restrict__ dist2, int* __restrict__ idx) {
    const int b = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    float b1 = INFINITY, b2 = INFINITY, b3 = INFINITY;
    int i1 = 0, i2 = 0, i3 = 0;
    for (int s = 0; s < nsplit; ++s) {
        const long long o = (((long long)b * nsplit + s) * n + j) * 3;
#pragma unroll
        for (int e = 0; e < 3; ++e) nn_insert(pd[o + e], pi[o + e], b1, b2, b3, i1, i2, i3);
    }
    const long long o = ((long long)b * n + j) * 3;
    dist2[o] = b1; dist2[o + 1] = b2; dist2[o + 2] = b3;
    idx[o] = i1; idx[o + 1] = i2; idx[o + 2] = i3;
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
    cudaStream_t st = (cudaStream_t)stream;
    if (!unknown || !known || !dist2 || !idx || B <= 0 || n <= 0 || m <= 0) return SED_ERR_ARG;
    const int qtiles = (n + NN_QCTA - 1) / NN_QCTA;
    // enough CTAs for ~4 per SM: split the known range when one cloud's query tiles do not fill the GPU
    int nsplit = (4 * kNumSMs + qtiles * B - 1) / (qtiles * B);
    nsplit = max(1, min(nsplit, min(8, (m + NN_TILE - 1) / NN_TILE)));
    const int chunk = ((m + nsplit - 1) / nsplit + NN_TILE - 1) / NN_TILE * NN_TILE;
    nsplit = (m + chunk - 1) / chunk;
    if (nsplit == 1) {
        three_nn_kernel<<<dim3(qtiles, 1, B), NN_THREADS, 0, st>>>(unknown, known, n, m, chunk, dist2, idx);
        SED_CHECK_LAUNCH();
        return SED_OK;
    }
    ensure_pool_config();
    char* ws = nullptr;
    const size_t part = (size_t)B * nsplit * n * 3;
    SED_CUDA(cudaMallocAsync((void**)&ws, part * (sizeof(float) + sizeof(int)), st));
    float* pd = reinterpret_cast<float*>(ws);
    int* pi = reinterpret_cast<int*>(ws + part * sizeof(float));
    three_nn_kernel<<<dim3(qtiles, nsplit, B), NN_THREADS, 0, st>>>(unknown, known, n, m, chunk, pd, pi);
    three_nn_merge_kernel<<<dim3((n + 255) / 256, B), 256, 0, st>>>(pd, pi, n, nsplit, dist2, idx);
    const cudaError_t e = cudaGetLastError();
    g_sed_launches += 2;
    cudaFreeAsync(ws, st);
    return e == cudaSuccess ? SED_OK : SED_ERR_CUDA_BASE - (int)e;
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
