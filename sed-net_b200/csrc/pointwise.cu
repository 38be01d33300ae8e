// Per-point network kernels of the SEDNet forward (reference src/SEDNet.py:78-98, :292-342):
//   * pw_gemm      1x1 convolution over channel-major activations with the previous layer's GroupNorm
//                  affine + activation applied while loading, bias in the epilogue, and GroupNorm partial
//                  statistics / per-channel max+min pooled in the epilogue (no normalised tensor is ever
//                  written: "normalise in the consumer's prologue");
//   * edge_reduce  EdgeConv as gather-reduce: W.[x_j - x_i ; x_i] = W1.x_j + (W2 - W1).x_i = U_j + V_i, so one
//                  per-point GEMM produces U|V and this kernel takes, per point and channel, max/min over the
//                  k neighbours plus the GroupNorm sums.  GroupNorm-affine + LeakyReLU is monotone per channel,
//                  so max_k f(y) = f(max_k y) or f(min_k y) depending on the sign of gamma*rstd;
//   * gn_finalize / edge_finalize / pool_finalize / gemv_bias / head_combine / log_softmax: the small glue.
#include <stdlib.h>
#include <string.h>

#include "internal.h"

namespace sed {

constexpr int PW_THREADS = 256;
constexpr int PW_BM = 64;    // output channels per CTA
constexpr int PW_BN = 128;   // points per CTA
constexpr int PW_KC = 16;    // input channels per staged chunk
constexpr int PW_AS = PW_BM + 4;  // padded row stride of the weight tile

struct PwParams {
    const float* X; long long x_bstride; int ldx;   // X[b*x_bstride + c*ldx + n]
    const float* Wt; int ldw;                        // Wt[co*ldw + c]
    const float* bias; long long bias_bstride;       // nullable; per-cloud when stride != 0
    const float* in_a; const float* in_s; int in_act;  // (B,Cin) affine on load (nullable); 0 none 1 relu 2 leaky(0.2)
    float* Y; long long y_bstride; int ldy; int y_point_major;  // nullable
    double* stats;   // (B, ntiles_n, ceil(Cout/32), 2) partial sum / sum of squares (nullable)
    float* mm;       // (B, ntiles_n, Cout, 2) partial max / min over points (nullable)
    int Cin, Cout, N;
};

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == 1) return fmaxf(v, 0.f);
    if (act == 2) return v >= 0.f ? v : 0.2f * v;
    return v;
}

__global__ void __launch_bounds__(PW_THREADS, 2) pw_gemm_kernel(PwParams p) {
    __shared__ __align__(16) float As[2][PW_KC][PW_AS];
    __shared__ __align__(16) float Bs[2][PW_KC][PW_BN];
    __shared__ double red[8][2];

    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int n0 = blockIdx.x * PW_BN, co0 = blockIdx.y * PW_BM, b = blockIdx.z;
    const float* X = p.X + (long long)b * p.x_bstride;
    const float* ia = p.in_a ? p.in_a + (long long)b * p.Cin : nullptr;
    const float* is = p.in_s ? p.in_s + (long long)b * p.Cin : nullptr;
    const int nk = (p.Cin + PW_KC - 1) / PW_KC;

    float ra[4], rb[8];
    auto load = [&](int kc) {
        const int k0 = kc * PW_KC;
        {   // weights: kk fastest (coalesced along Cin)
            const int kk = tid & 15;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int co = (tid >> 4) + 16 * r;
                ra[r] = (k0 + kk < p.Cin && co0 + co < p.Cout) ? __ldg(p.Wt + (long long)(co0 + co) * p.ldw + k0 + kk) : 0.f;
            }
        }
        {   // activations: n fastest
            const int n = n0 + (tid & 127);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int c = k0 + (tid >> 7) + 2 * r;
                float v = 0.f;
                if (c < p.Cin && n < p.N) {
                    v = __ldg(X + (long long)c * p.ldx + n);
                    if (ia) v = apply_act(fmaf(__ldg(ia + c), v, __ldg(is + c)), p.in_act);
                }
                rb[r] = v;
            }
        }
    };
    auto store = [&](int buf) {
        const int kk = tid & 15;
#pragma unroll
        for (int r = 0; r < 4; ++r) As[buf][kk][(tid >> 4) + 16 * r] = ra[r];
#pragma unroll
        for (int r = 0; r < 8; ++r) Bs[buf][(tid >> 7) + 2 * r][tid & 127] = rb[r];
    };

    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[i][e] = 0.f;

    load(0);
    store(0);
    __syncthreads();
    for (int kc = 0; kc < nk; ++kc) {
        const int buf = kc & 1;
        if (kc + 1 < nk) load(kc + 1);
#pragma unroll
        for (int kk = 0; kk < PW_KC; ++kk) {
            float4 a = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
            float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
            float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
            const float aa[4] = {a.x, a.y, a.z, a.w};
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[i][e] = fmaf(aa[i], bb[e], acc[i][e]);
        }
        if (kc + 1 < nk) store(buf ^ 1);
        __syncthreads();
    }

    // ---- epilogue ----
    const float* bias = p.bias ? p.bias + (long long)b * p.bias_bstride : nullptr;
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int co = co0 + ty * 4 + i;
        const float bv = (bias && co < p.Cout) ? __ldg(bias + co) : 0.f;
        float mx = -INFINITY, mn = INFINITY;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int n = n0 + ((e < 4) ? (tx * 4 + e) : (64 + tx * 4 + e - 4));
            const float y = acc[i][e] + bv;
            acc[i][e] = y;
            if (n < p.N && co < p.Cout) {
                s1 += y; s2 = fmaf(y, y, s2);
                mx = fmaxf(mx, y); mn = fminf(mn, y);
            }
        }
        if (p.mm) {
#pragma unroll
            for (int m = 8; m > 0; m >>= 1) {
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, m));
                mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, m));
            }
            if (tx == 0 && co < p.Cout) {
                float* o = p.mm + (((long long)b * gridDim.x + blockIdx.x) * p.Cout + co) * 2;
                o[0] = mx; o[1] = mn;
            }
        }
    }
    if (p.Y) {
        float* Y = p.Y + (long long)b * p.y_bstride;
        if (p.y_point_major) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int n = n0 + ((e < 4) ? (tx * 4 + e) : (64 + tx * 4 + e - 4));
                if (n >= p.N) continue;
                const int co = co0 + ty * 4;
                float* o = Y + (long long)n * p.ldy + co;
                if (co + 3 < p.Cout && ((p.ldy & 3) == 0)) {
                    *reinterpret_cast<float4*>(o) = make_float4(acc[0][e], acc[1][e], acc[2][e], acc[3][e]);
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) if (co + i < p.Cout) o[i] = acc[i][e];
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int co = co0 + ty * 4 + i;
                if (co >= p.Cout) continue;
                float* o = Y + (long long)co * p.ldy;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int n = n0 + h * 64 + tx * 4;
                    if (n + 3 < p.N && ((p.ldy & 3) == 0)) {
                        *reinterpret_cast<float4*>(o + n) = make_float4(acc[i][4 * h], acc[i][4 * h + 1], acc[i][4 * h + 2], acc[i][4 * h + 3]);
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e) if (n + e < p.N) o[n + e] = acc[i][4 * h + e];
                    }
                }
            }
        }
    }
    if (p.stats) {
        // a warp holds ty = 2w, 2w+1 -> one 32-channel block (w / 4)
        double d1 = warp_sum_d((double)s1), d2 = warp_sum_d((double)s2);
        const int warp = tid >> 5;
        if ((tid & 31) == 0) { red[warp][0] = d1; red[warp][1] = d2; }
        __syncthreads();
        if (tid < 2) {
            const int blk = blockIdx.y * 2 + tid;
            const int nblk = (p.Cout + 31) / 32;
            if (blk < nblk) {
                double t1 = 0.0, t2 = 0.0;
                for (int w = 0; w < 4; ++w) { t1 += red[tid * 4 + w][0]; t2 += red[tid * 4 + w][1]; }
                double* o = p.stats + (((long long)b * gridDim.x + blockIdx.x) * nblk + blk) * 2;
                o[0] = t1; o[1] = t2;
            }
        }
    }
}

// ---- GroupNorm finalize: partial sums -> per-(cloud, channel) scale a = gamma*rstd and shift s = beta - a*mean ----
// partial layout (B, P, NBLK, 2); group g owns blocks [g*bpg, (g+1)*bpg).  One CTA per cloud.
__global__ void gn_finalize_kernel(const double* __restrict__ part, int P, int NBLK, int bpg, double count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, int C, int G,
                                   float eps, float* __restrict__ a_out, float* __restrict__ s_out) {
    extern __shared__ double sh[];  // [G][2] mean, rstd
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    // one warp per group: lanes stride over the (tile, block) partials, fixed-shape shuffle tree (deterministic)
    for (int g = warp; g < G; g += nw) {
        double t1 = 0.0, t2 = 0.0;
        for (int e = lane; e < P * bpg; e += 32) {
            const int q = e / bpg, k = e - q * bpg;
            const double* o = part + (((long long)b * P + q) * NBLK + g * bpg + k) * 2;
            t1 += o[0]; t2 += o[1];
        }
        t1 = warp_sum_d(t1); t2 = warp_sum_d(t2);
        if (lane == 0) {
            const double mean = t1 / count;
            double var = t2 / count - mean * mean;
            var = var > 0.0 ? var : 0.0;
            sh[2 * g] = mean;
            sh[2 * g + 1] = 1.0 / sqrt(var + (double)eps);
        }
    }
    __syncthreads();
    const int gs = C / G;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int g = c / gs;
        const double a = (double)gamma[c] * sh[2 * g + 1];
        a_out[(long long)b * C + c] = (float)a;
        s_out[(long long)b * C + c] = (float)((double)beta[c] - a * sh[2 * g]);
    }
}

// ---- EdgeConv gather-reduce ----
struct EdgeParams {
    const float* UV;   // (B, N, 2*Cout) point-major: U in [0,Cout), V in [Cout, 2*Cout)
    const int* idx;    // (B, N, k)
    float* ymax; float* ymin;  // (B, Cout, N)
    double* stats;     // (B, ceil(N/32), G, 2)
    int N, k, Cout, G;
};

template <int CPL>  // channels per lane = Cout / 32
__global__ void __launch_bounds__(256) edge_reduce_kernel(EdgeParams p) {
    extern __shared__ __align__(16) unsigned char esm[];
    const int Cout = 32 * CPL;
    float* smax = reinterpret_cast<float*>(esm);            // [Cout][33]
    float* smin = smax + Cout * 33;                          // [Cout][33]
    double* sred = reinterpret_cast<double*>(smin + Cout * 33);  // [8 warps][32 lanes][2]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y, n0 = blockIdx.x * 32;
    const float* UV = p.UV + (long long)b * p.N * 2 * Cout;
    const int* idx = p.idx + (long long)b * p.N * p.k;
    double d1 = 0.0, d2 = 0.0;
    for (int pi = 0; pi < 4; ++pi) {
        const int nl = warp * 4 + pi, n = n0 + nl;
        float mx[CPL], mn[CPL], s1[CPL], s2[CPL], v[CPL];
        if (n < p.N) {
            {
                const float* vr = UV + (long long)n * 2 * Cout + Cout + lane * CPL;
                if (CPL == 4) { float4 t = __ldg(reinterpret_cast<const float4*>(vr)); v[0] = t.x; v[1] = t.y; v[2 % CPL] = t.z; v[3 % CPL] = t.w; }
                else { float2 t = __ldg(reinterpret_cast<const float2*>(vr)); v[0] = t.x; v[1] = t.y; }
            }
#pragma unroll
            for (int e = 0; e < CPL; ++e) { mx[e] = -INFINITY; mn[e] = INFINITY; s1[e] = 0.f; s2[e] = 0.f; }
            for (int jb = 0; jb < p.k; jb += 32) {
                const int mine = (jb + lane < p.k) ? __ldg(idx + (long long)n * p.k + jb + lane) : 0;
                const int cntj = min(32, p.k - jb);
                for (int j4 = 0; j4 < cntj; j4 += 4) {
                    float u[4][CPL];
#pragma unroll
                    for (int u4 = 0; u4 < 4; ++u4) {
                        const int nb = __shfl_sync(0xffffffffu, mine, (j4 + u4) & 31);
                        const float* ur = UV + (long long)nb * 2 * Cout + lane * CPL;
                        if (CPL == 4) { float4 t = __ldg(reinterpret_cast<const float4*>(ur)); u[u4][0] = t.x; u[u4][1] = t.y; u[u4][2 % CPL] = t.z; u[u4][3 % CPL] = t.w; }
                        else { float2 t = __ldg(reinterpret_cast<const float2*>(ur)); u[u4][0] = t.x; u[u4][1] = t.y; }
                    }
#pragma unroll
                    for (int u4 = 0; u4 < 4; ++u4) {
                        if (j4 + u4 < cntj) {
#pragma unroll
                            for (int e = 0; e < CPL; ++e) {
                                const float y = u[u4][e] + v[e];
                                mx[e] = fmaxf(mx[e], y); mn[e] = fminf(mn[e], y);
                                s1[e] += y; s2[e] = fmaf(y, y, s2[e]);
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int e = 0; e < CPL; ++e) {
                smax[(lane * CPL + e) * 33 + nl] = mx[e];
                smin[(lane * CPL + e) * 33 + nl] = mn[e];
                d1 += (double)s1[e]; d2 += (double)s2[e];
            }
        }
    }
    sred[(warp * 32 + lane) * 2] = d1;
    sred[(warp * 32 + lane) * 2 + 1] = d2;
    __syncthreads();
    // coalesced channel-major stores: 32 consecutive points per channel row
    for (int e = tid; e < Cout * 32; e += 256) {
        const int c = e >> 5, nl = e & 31;
        if (n0 + nl < p.N) {
            const long long o = ((long long)b * Cout + c) * p.N + n0 + nl;
            p.ymax[o] = smax[c * 33 + nl];
            p.ymin[o] = smin[c * 33 + nl];
        }
    }
    if (tid < p.G) {
        const int gs = Cout / p.G;           // channels per group
        const int lpg = gs / CPL;            // lanes per group
        double t1 = 0.0, t2 = 0.0;
        for (int w = 0; w < 8; ++w)
            for (int l = tid * lpg; l < (tid + 1) * lpg; ++l) { t1 += sred[(w * 32 + l) * 2]; t2 += sred[(w * 32 + l) * 2 + 1]; }
        double* o = p.stats + (((long long)b * gridDim.x + blockIdx.x) * p.G + tid) * 2;
        o[0] = t1; o[1] = t2;
    }
}

// x[b, c, n] = leaky(a * (a >= 0 ? ymax : ymin) + s): max over k of GN+LeakyReLU (reference src/SEDNet.py:81-82).
__global__ void edge_finalize_kernel(const float* __restrict__ ymax, const float* __restrict__ ymin,
                                     const float* __restrict__ a, const float* __restrict__ s, int C, int N,
                                     float slope, float* __restrict__ out, long long out_bstride) {
    const int b = blockIdx.z, c = blockIdx.y;
    const float av = a[(long long)b * C + c], sv = s[(long long)b * C + c];
    const float* src = (av >= 0.f ? ymax : ymin) + ((long long)b * C + c) * N;
    float* dst = out + (long long)b * out_bstride + (long long)c * N;
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
        const float y = fmaf(av, src[n], sv);
        dst[n] = y >= 0.f ? y : slope * y;
    }
}

// x4[b, c] = relu(a * (a >= 0 ? max_n y : min_n y) + s) from the tile partials (reference src/SEDNet.py:95-96).
__global__ void pool_finalize_kernel(const float* __restrict__ mm, int P, int C, const float* __restrict__ a,
                                     const float* __restrict__ s, float* __restrict__ out) {
    const int b = blockIdx.y, c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float mx = -INFINITY, mn = INFINITY;
    for (int q = 0; q < P; ++q) {
        const float* o = mm + (((long long)b * P + q) * C + c) * 2;
        mx = fmaxf(mx, o[0]); mn = fminf(mn, o[1]);
    }
    const float av = a[(long long)b * C + c], sv = s[(long long)b * C + c];
    out[(long long)b * C + c] = fmaxf(fmaf(av, av >= 0.f ? mx : mn, sv), 0.f);
}

// out[b, co] = bias[co] + sum_c W[co*ldw + c] * v[b, c]: the global-feature half of conv1 hoisted out of the
// per-point GEMM (reference src/SEDNet.py:300-303 repeats the 1024-d vector N times instead).
__global__ void gemv_bias_kernel(const float* __restrict__ Wt, int ldw, const float* __restrict__ bias,
                                 const float* __restrict__ v, int Cin, int Cout, float* __restrict__ out) {
    const int b = blockIdx.y, co = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (co >= Cout) return;
    float acc = 0.f;
    for (int c = lane; c < Cin; c += 32) acc = fmaf(__ldg(Wt + (long long)co * ldw + c), v[(long long)b * Cin + c], acc);
    acc = warp_sum_f(acc);
    if (lane == 0) out[(long long)b * Cout + co] = acc + (bias ? bias[co] : 0.f);
}

// x = w * relu(aa*ya + sa) + relu(as*ys + ss); x = x + w * relu(pe)   (reference src/SEDNet.py:320-326)
__global__ void head_combine_kernel(const float* __restrict__ ys, const float* __restrict__ as_, const float* __restrict__ ss,
                                    const float* __restrict__ ya, const float* __restrict__ aa, const float* __restrict__ sa,
                                    const float* __restrict__ pe, float w, int C, int N, float* __restrict__ out) {
    const int b = blockIdx.z, c = blockIdx.y;
    const long long row = ((long long)b * C + c) * N;
    const float a1 = as_[(long long)b * C + c], s1 = ss[(long long)b * C + c];
    const float a2 = aa[(long long)b * C + c], s2 = sa[(long long)b * C + c];
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
        const float xs = fmaxf(fmaf(a1, ys[row + n], s1), 0.f);
        const float xa = fmaxf(fmaf(a2, ya[row + n], s2), 0.f);
        float x = __fadd_rn(__fmul_rn(w, xa), xs);
        x = __fadd_rn(x, __fmul_rn(w, fmaxf(pe[row + n], 0.f)));
        out[row + n] = x;
    }
}

// log_softmax over C (<= 16) channels of a channel-major tensor with row stride ld (reference src/SEDNet.py:314).
__global__ void log_softmax_kernel(const float* __restrict__ x, long long x_bstride, int C, int N,
                                   float* __restrict__ out) {
    const int b = blockIdx.y, n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float* xb = x + (long long)b * x_bstride;
    float v[16];
    float m = -INFINITY;
    for (int c = 0; c < C; ++c) { v[c] = xb[(long long)c * N + n]; m = fmaxf(m, v[c]); }
    float sum = 0.f;
    for (int c = 0; c < C; ++c) sum += expf(v[c] - m);
    const float l = logf(sum);
    for (int c = 0; c < C; ++c) out[((long long)b * C + c) * N + n] = (v[c] - m) - l;
}

// one_hot[n, label[n]] = 1 (reference src/segment_utils.py:536-545)
__global__ void one_hot_kernel(const long long* __restrict__ labels, int N, int K, float* __restrict__ out) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const long long l = labels[n];
    for (int k = 0; k < K; ++k) out[(long long)n * K + k] = (k == l) ? 1.f : 0.f;
}


// Wf[co][c] = W[co][c] (U half, applied to x_j), Wf[Cout+co][c] = W[co][Cin+c] - W[co][c] (V half, applied to x_i):
// W.[x_j - x_i ; x_i] = W1.x_j + (W2 - W1).x_i.   W is (Cout, 2*Cin) as in encoder.convK.0.weight.
__global__ void edge_fold_kernel(const float* __restrict__ W, int Cout, int Cin, float* __restrict__ Wf) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= Cout * Cin) return;
    const int co = e / Cin, c = e - co * Cin;
    const float w1 = W[(long long)co * 2 * Cin + c], w2 = W[(long long)co * 2 * Cin + Cin + c];
    Wf[(long long)co * Cin + c] = w1;
    Wf[(long long)(Cout + co) * Cin + c] = w2 - w1;
}

// out (B,2C,N,k) = cat([x_j - x_i, x_i]) (reference src/PointNet.py:161-170); API parity only.
__global__ void graph_feature_kernel(const float* __restrict__ x, const long long* __restrict__ idx, int C, int N,
                                     int k, float* __restrict__ out) {
    const int b = blockIdx.z, c = blockIdx.y;
    const float* xb = x + ((long long)b * C + c) * N;
    const long long* ib = idx + (long long)b * N * k;
    float* o1 = out + ((long long)b * 2 * C + c) * N * k;
    float* o2 = out + ((long long)b * 2 * C + C + c) * N * k;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < (long long)N * k;
         e += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(e / k);
        const float xi = xb[n];
        o1[e] = xb[ib[e]] - xi;
        o2[e] = xi;
    }
}

// Few-output-channel layers (Cout <= 8: src/SEDNet.py:305 mlp_prim_prob2 256 -> 6 and :317 edge_module.2 128 -> 2):
// next to no arithmetic, so a tensor-core tile is all overhead (60 / 45 us on pw_tc_kernel for 0.06 us of MMA work) and the
// layer is bound by reading its input once.  A CTA owns 128 points; each of its eight warps walks one eighth of the input
// channels with 16-byte loads (a lane = 4 consecutive points, 512 contiguous bytes per warp and channel, eight loads in flight),
// applies the producer's affine + activation and keeps 4 x Cout sums in registers against weights broadcast from shared
// memory; the eight partial sums meet in shared memory.  (A first version with one thread per point over ALL channels ran at
// 142 / 74 us: 550 threads per SM cannot cover the latency of 256 dependent-issue loads.)
constexpr int PWS_MAX_COUT = 8, PWS_THREADS = 256, PWS_PTS = 128, PWS_WARPS = PWS_THREADS / 32;

__global__ void __launch_bounds__(PWS_THREADS) pw_small_kernel(PwParams p) {
    extern __shared__ __align__(16) float sm[];        // W [Cin][8] | a [Cin] | s [Cin] | red [8 warps][8][128]
    float* Ws = sm;
    float* as_ = Ws + p.Cin * PWS_MAX_COUT;
    float* ss = as_ + p.Cin;
    float* red = ss + p.Cin;
    const int b = blockIdx.y, n0 = blockIdx.x * PWS_PTS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int e = threadIdx.x; e < p.Cin * PWS_MAX_COUT; e += PWS_THREADS) {
        const int c = e >> 3, co = e & 7;
        Ws[e] = co < p.Cout ? p.Wt[(long long)co * p.ldw + c] : 0.f;
    }
    for (int c = threadIdx.x; c < p.Cin; c += PWS_THREADS) {
        as_[c] = p.in_a ? p.in_a[(long long)b * p.Cin + c] : 1.f;
        ss[c] = p.in_s ? p.in_s[(long long)b * p.Cin + c] : 0.f;
    }
    __syncthreads();
    const int n = n0 + lane * 4;
    const float* X = p.X + (long long)b * p.x_bstride + n;
    const bool vec = ((p.ldx & 3) == 0) && ((p.x_bstride & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.X) & 15) == 0) &&
                     (n + 3 < p.N);
    float acc[PWS_MAX_COUT][4];
#pragma unroll
    for (int co = 0; co < PWS_MAX_COUT; ++co)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[co][e] = 0.f;
    const bool affine = p.in_a != nullptr;
    const int per = (p.Cin + PWS_WARPS - 1) / PWS_WARPS;
    const int c_begin = warp * per, c_end = min(p.Cin, c_begin + per);
    constexpr int U = 8;
    for (int c0 = c_begin; c0 < c_end; c0 += U) {
        float4 x[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {                  // all loads of the group first
            const int c = c0 + u;
            x[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < c_end) {
                const float* xr = X + (long long)c * p.ldx;
                if (vec) x[u] = __ldg(reinterpret_cast<const float4*>(xr));
                else {
                    if (n < p.N) x[u].x = __ldg(xr);
                    if (n + 1 < p.N) x[u].y = __ldg(xr + 1);
                    if (n + 2 < p.N) x[u].z = __ldg(xr + 2);
                    if (n + 3 < p.N) x[u].w = __ldg(xr + 3);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int c = c0 + u;
            if (c < c_end) {
                float v[4] = {x[u].x, x[u].y, x[u].z, x[u].w};
                if (affine) {
                    const float av = as_[c], sv = ss[c];
#pragma unroll
                    for (int e = 0; e < 4; ++e) v[e] = apply_act(fmaf(av, v[e], sv), p.in_act);
                }
                const float4 w0 = *reinterpret_cast<const float4*>(Ws + c * PWS_MAX_COUT);
                const float4 w1 = *reinterpret_cast<const float4*>(Ws + c * PWS_MAX_COUT + 4);
                const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                for (int co = 0; co < PWS_MAX_COUT; ++co)
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc[co][e] = fmaf(w[co], v[e], acc[co][e]);
            }
        }
    }
#pragma unroll
    for (int co = 0; co < PWS_MAX_COUT; ++co)
        *reinterpret_cast<float4*>(red + ((warp * PWS_MAX_COUT + co) * PWS_PTS) + lane * 4) =
            make_float4(acc[co][0], acc[co][1], acc[co][2], acc[co][3]);
    __syncthreads();
    // thread = (point, half of the output channels): fixed-order sum over the eight warps
    const int pt = threadIdx.x & (PWS_PTS - 1), half = threadIdx.x >> 7;
    const int nn = n0 + pt;
    if (nn >= p.N) return;
    float* Y = p.Y + (long long)b * p.y_bstride;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int co = half * 4 + q;
        if (co < p.Cout) {
            float y = 0.f;
#pragma unroll
            for (int w = 0; w < PWS_WARPS; ++w) y += red[(w * PWS_MAX_COUT + co) * PWS_PTS + pt];
            y = __fadd_rn(y, p.bias ? p.bias[(long long)b * p.bias_bstride + co] : 0.f);
            if (p.y_point_major) Y[(long long)nn * p.ldy + co] = y;
            else Y[(long long)co * p.ldy + nn] = y;
        }
    }
}

int pw_gemm(const float* X, long long x_bstride, int ldx, const float* Wt, int ldw, const float* bias,
            long long bias_bstride, const float* in_a, const float* in_s, int in_act, float* Y, long long y_bstride,
            int ldy, int y_point_major, double* stats, float* mm, int B, int Cin, int Cout, int N, cudaStream_t stream) {
    if (!X || !Wt || B <= 0 || Cin <= 0 || Cout <= 0 || N <= 0) return SED_ERR_ARG;
    if ((in_a == nullptr) != (in_s == nullptr)) return SED_ERR_ARG;
    // SEDNET_B200_PW=ffma forces the CUDA-core kernel of this file (A/B comparisons); default: tensor cores
    static const bool ffma = [] { const char* e = getenv("SEDNET_B200_PW"); return e && !strcmp(e, "ffma"); }();
    if (!ffma && Cout <= PWS_MAX_COUT && Cin >= 32 && Cin <= 1024 && Y && !stats && !mm) {
        PwParams ps{X, x_bstride, ldx, Wt, ldw, bias, bias_bstride, in_a, in_s, in_act, Y, y_bstride, ldy, y_point_major,
                    nullptr, nullptr, Cin, Cout, N};
        const size_t smem = ((size_t)(PWS_MAX_COUT + 2) * Cin + (size_t)PWS_WARPS * PWS_MAX_COUT * PWS_PTS) * sizeof(float);
        // 32 KB of partial sums + up to 40 KB of weights: above the 48 KB default
        SED_CUDA(cudaFuncSetAttribute(pw_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
        pw_small_kernel<<<dim3((N + PWS_PTS - 1) / PWS_PTS, B), PWS_THREADS, smem, stream>>>(ps);
        SED_CHECK_LAUNCH();
        return SED_OK;
    }
    // SEDNET_B200_PW=tc1 keeps every shape on the first tensor-core kernel (one tile per CTA, pointwise_tc.cu)
    static const bool tc1 = [] { const char* e = getenv("SEDNET_B200_PW"); return e && !strcmp(e, "tc1"); }();
    if (!ffma && !tc1) {
        const int rc = pw_gemm_tc2(X, x_bstride, ldx, Wt, ldw, bias, bias_bstride, in_a, in_s, in_act, Y, y_bstride, ldy,
                                   y_point_major, stats, mm, B, Cin, Cout, N, stream);
        if (rc != SED_ERR_UNSUPPORTED) return rc;
    }
    if (!ffma) {
        const int rc = pw_gemm_tc(X, x_bstride, ldx, Wt, ldw, bias, bias_bstride, in_a, in_s, in_act, Y, y_bstride, ldy,
                                  y_point_major, stats, mm, B, Cin, Cout, N, stream);
        if (rc != SED_ERR_UNSUPPORTED) return rc;
    }
    PwParams p{X, x_bstride, ldx, Wt, ldw, bias, bias_bstride, in_a, in_s, in_act, Y, y_bstride, ldy, y_point_major,
               stats, mm, Cin, Cout, N};
    dim3 grid((N + PW_BN - 1) / PW_BN, (Cout + PW_BM - 1) / PW_BM, B);
    pw_gemm_kernel<<<grid, PW_THREADS, 0, stream>>>(p);
    SED_CHECK_LAUNCH();
    return SED_OK;
}

int gn_finalize(const double* part, int P, int NBLK, int blocks_per_group, double count, const float* gamma,
                const float* beta, int B, int C, int G, float eps, float* a_out, float* s_out, cudaStream_t stream) {
    if (!part || !gamma || !beta || !a_out || !s_out || G <= 0 || C % G) return SED_ERR_ARG;
    gn_finalize_kernel<<<B, 256, G * 2 * sizeof(double), stream>>>(part, P, NBLK, blocks_per_group, count, gamma, beta,
                                                                    C, G, eps, a_out, s_out);
    SED_CHECK_LAUNCH();
    return SED_OK;
}

int edge_fold_weights(const float* W, int Cout, int Cin, float* Wf, cudaStream_t stream) {
    edge_fold_kernel<<<(Cout * Cin + 255) / 256, 256, 0, stream>>>(W, Cout, Cin, Wf);
    SED_CHECK_LAUNCH();
    return SED_OK;
}

// stats partial layout (B, ceil(N/32), G, 2) doubles.
int edge_reduce(const float* UV, const int* idx, float* ymax, float* ymin, double* stats, int B, int N, int k,
                int Cout, int G, cudaStream_t stream) {
    if (!UV || !idx || !ymax || !ymin || !stats || G <= 0 || G > 32) return SED_ERR_ARG;
    EdgeParams p{UV, idx, ymax, ymin, stats, N, k, Cout, G};
    dim3 grid((N + 31) / 32, B);
    size_t smem = (size_t)Cout * 33 * 2 * sizeof(float) + 8 * 32 * 2 * sizeof(double);
    if (Cout == 64 && (64 / G) % 2 == 0) {
        edge_reduce_kernel<2><<<grid, 256, smem, stream>>>(p);
    } else if (Cout == 128 && (128 / G) % 4 == 0) {
        SED_CUDA(cudaFuncSetAttribute(edge_reduce_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        edge_reduce_kernel<4><<<grid, 256, smem, stream>>>(p);
    } else {
        return SED_ERR_UNSUPPORTED;
    }
    SED_CHECK_LAUNCH();
    return SED_OK;
}

int edge_finalize(const float* ymax, const float* ymin, const float* a, const float* s, int B, int C, int N,
                  float slope, float* out, long long out_bstride, cudaStream_t stream) {
    dim3 grid((N + 1023) / 1024, C, B);
    edge_finalize_kernel<<<grid, 256, 0, stream>>>(ymax, ymin, a, s, C, N, slope, out, out_bstride);
    SED_CHECK_LAUNCH();
    return SED_OK;
}

int pool_finalize(const float* mm, int P, int B, int C, const float* a, const float* s, float* out, cudaStream_t stream) {
    dim3 grid((C + 127) / 128, B);
    pool_finalize_kernel<<<grid, 128, 0, stream>>>(mm, P, C, a, s, out);
    SED_CHECK_LAUNCH();
    return SED_OK;
}

int gemv_bias(const float* Wt, int ldw, const float* bias, const float* v, int B, int Cin, int Cout, float* out,
              cudaStream_t stream) {
    dim3 grid((Cout + 7) / 8, B);
    gemv_bias_kernel<<<grid, 256, 0, stream>>>(Wt, ldw, bias, v, Cin, Cout, out);
    SED_CHECK_LAUNCH();
    return SED_OK;
}

int head_combine(const float* ys, const float* as_, const float* ss, const float* ya, const float* aa, const float* sa,
                 const float* pe, float w, int B, int C, int N, float* out, cudaStream_t stream) {
    dim3 grid((N + 1023) / 1024, C, B);
    head_combine_kernel<<<grid, 256, 0, stream>>>(ys, as_, ss, ya, aa, sa, pe, w, C, N, out);
    SED_CHECK_LAUNCH();
    return SED_OK;
}

int log_softmax(const float* x, long long x_bstride, int B, int C, int N, float* out, cudaStream_t stream) {
    if (C > 16) return SED_ERR_UNSUPPORTED;
    dim3 grid((N + 255) / 256, B);
    log_softmax_kernel<<<grid, 256, 0, stream>>>(x, x_bstride, C, N, out);
    SED_CHECK_LAUNCH();
    return SED_OK;
}

}  // namespace sed

using namespace sed;

extern "C" {

int sed_one_hot(const int64_t* labels, int N, int K, float* out, sed_stream_t stream) {
    if (!labels || !out || N <= 0 || K <= 0) return SED_ERR_ARG;
    one_hot_kernel<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const long long*)labels, N, K, out);
    SED_CHECK_LAUNCH();
    return SED_OK;
}

int sed_graph_feature(const float* x, const int64_t* idx, int B, int C, int N, int k, float* out, sed_stream_t stream) {
    if (!x || !idx || !out || B <= 0 || C <= 0 || N <= 0 || k <= 0) return SED_ERR_ARG;
    dim3 grid((unsigned)(((long long)N * k + 1023) / 1024), C, B);
    graph_feature_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, (const long long*)idx, C, N, k, out);
    SED_CHECK_LAUNCH();
    return SED_OK;
}

}  // extern "C"
