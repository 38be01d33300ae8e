// Batched weighted least-squares primitive fits and point->primitive residuals.
//
// Reference: Fit.fit_{plane,sphere,cylinder,cone}_torch src/primitive_forward.py:712-847, LeastSquares.lstsq /
// best_lambda src/fitting_utils.py:32-85, ComputePrimitiveDistance src/primitives.py:47-206, the per-segment
// dispatch of fit_one_shape_torch src/primitive_forward.py:929-1051.
//
// One CTA per (cloud, segment).  The reference runs one cuSOLVER SVD / QR and a host np.linalg.cond per segment;
// here every fit reduces to a handful of 3x3 moment matrices accumulated in FP64 over the segment's points
// (deterministic block reductions), followed by a 3x3 Jacobi eigen-decomposition in FP64:
//   * right singular vectors of an (n x 3) matrix A   = eigenvectors of A^T A;
//   * singular values (rank test, condition number)   = sqrt of its eigenvalues;
//   * the QR least-squares solution                    = (A^T A)^-1 A^T Y  (full rank), and the reference's
//     regularised branch is literally (A^T A + lambda I)^-1 A^T Y with lambda from the 1e-6 * 10^i ladder.
#include <algorithm>

#include "internal.h"

namespace sed {

constexpr int FIT_THREADS = 256;
constexpr double kEps32 = 1.1920928955078125e-07;  // np.finfo(np.float32).eps (EPS in the reference)

struct Sym3 { double xx, xy, xz, yy, yz, zz; };

// Cyclic Jacobi for a symmetric 3x3 matrix. Eigenvalues descending in w, matching unit eigenvectors in V[:, i].
__device__ void eig3(const Sym3& S, double w[3], double V[3][3]) {
    double A[3][3] = {{S.xx, S.xy, S.xz}, {S.xy, S.yy, S.yz}, {S.xz, S.yz, S.zz}};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 24; ++sweep) {
        const double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
        const double dia = A[0][0] * A[0][0] + A[1][1] * A[1][1] + A[2][2] * A[2][2];
        if (off <= 1e-34 * dia || off == 0.0) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (A[p][q] == 0.0) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) {  // A <- A J
                    const double akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - s * akq;
                    A[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {  // A <- J^T A
                    const double apk = A[p][k], aqk = A[q][k];
                    A[p][k] = c * apk - s * aqk;
                    A[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    const double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - s * vkq;
                    V[k][q] = s * vkp + c * vkq;
                }
            }
    }
    w[0] = A[0][0]; w[1] = A[1][1]; w[2] = A[2][2];
    for (int i = 0; i < 2; ++i)      // sort descending (3 elements)
        for (int j = 0; j < 2 - i; ++j)
            if (w[j] < w[j + 1]) {
                double t = w[j]; w[j] = w[j + 1]; w[j + 1] = t;
                for (int k = 0; k < 3; ++k) { t = V[k][j]; V[k][j] = V[k][j + 1]; V[k][j + 1] = t; }
            }
}

// Smallest right singular vector (V[:, -1] of torch.svd, src/fitting_utils.py:440) with a canonical sign:
// the component of largest magnitude is positive (LAPACK's sign is arbitrary; parity is modulo this sign).
__device__ void smallest_vec(const Sym3& S, double a[3], double w[3]) {
    double V[3][3];
    eig3(S, w, V);
    a[0] = V[0][2]; a[1] = V[1][2]; a[2] = V[2][2];
    int m = 0;
    if (fabs(a[1]) > fabs(a[m])) m = 1;
    if (fabs(a[2]) > fabs(a[m])) m = 2;
    if (a[m] < 0.0) { a[0] = -a[0]; a[1] = -a[1]; a[2] = -a[2]; }
}

// LeastSquares.lstsq (src/fitting_utils.py:36-65) on the normal equations G = A^T A, g = A^T Y of an (nrows x 3)
// system.  Returns 0 (full rank, QR branch) or 2 (regularised branch, best_lambda :68-85).
__device__ int lstsq3(const Sym3& G, const double g[3], int nrows, double x[3]) {
    double mu[3], V[3][3];
    eig3(G, mu, V);
    const double s0 = sqrt(fmax(mu[0], 0.0)), s2 = sqrt(fmax(mu[2], 0.0));
    // torch.linalg.matrix_rank default tolerance: sigma_max * max(m, n) * eps(float32)
    const double tol = s0 * (double)max(nrows, 3) * kEps32;
    double lam = 0.0;
    int status = 0;
    if (!(s2 > tol)) {
        status = 2;
        lam = 1e-6;
        for (int i = 0; i < 7; ++i) {
            // rank(AtA + lam I) == 3  <=>  smallest singular value mu2 + lam above (mu0 + lam) * 3 * eps
            if ((mu[2] + lam) > (mu[0] + lam) * 3.0 * kEps32) break;
            lam *= 10.0;
        }
    }
    double y[3];
    for (int i = 0; i < 3; ++i) {
        const double proj = V[0][i] * g[0] + V[1][i] * g[1] + V[2][i] * g[2];
        const double den = mu[i] + lam;
        y[i] = den != 0.0 ? proj / den : 0.0;
    }
    for (int k = 0; k < 3; ++k) x[k] = V[k][0] * y[0] + V[k][1] * y[1] + V[k][2] * y[2];
    return status;
}

struct FitParams {
    const float* pts; const float* nrm; const float* wts; const long long* labels; const int* seg_type;
    int N, S, min_pts;
    float* params; int* status;
};

struct SegIter {
    const float* pts; const float* nrm; const float* wts; const long long* labels; int N, s;
    const unsigned char* mask = nullptr;   // stage-2 centre crop: members with mask[i] == 0 are left out
    __device__ __forceinline__ bool member(int i) const {
        return (!labels || labels[i] == (long long)s) && (!mask || mask[i]);
    }
    __device__ __forceinline__ double w(int i) const { return wts ? (double)wts[i] : 1.0; }
    __device__ __forceinline__ void p(int i, double v[3]) const {
        v[0] = pts[3 * i]; v[1] = pts[3 * i + 1]; v[2] = pts[3 * i + 2];
    }
    __device__ __forceinline__ void n(int i, double v[3]) const {
        v[0] = nrm[3 * i]; v[1] = nrm[3 * i + 1]; v[2] = nrm[3 * i + 2];
    }
};

struct Shared {
    double scratch[16 * (FIT_THREADS / 32)];
    double out[16];
};

template <int NV>
__device__ __forceinline__ void reduce(double (&v)[NV], Shared& sh) { block_sum_d<NV>(v, sh.scratch, sh.out); }

__device__ __forceinline__ void add_outer(double (&acc)[6], double s, const double v[3]) {
    acc[0] += s * v[0] * v[0]; acc[1] += s * v[0] * v[1]; acc[2] += s * v[0] * v[2];
    acc[3] += s * v[1] * v[1]; acc[4] += s * v[1] * v[2]; acc[5] += s * v[2] * v[2];
}
__device__ __forceinline__ Sym3 sym(const double* a) { return Sym3{a[0], a[1], a[2], a[3], a[4], a[5]}; }

// fit_plane_torch on `use_normals ? normals : points` (src/primitive_forward.py:712-733): returns the unit axis
// (canonical sign) and d.
__device__ void plane_fit(const SegIter& it, bool use_normals, Shared& sh, double a[3], double& dval) {
    double m[4] = {0, 0, 0, 0};
    for (int i = threadIdx.x; i < it.N; i += FIT_THREADS) {
        if (!it.member(i)) continue;
        double v[3]; use_normals ? it.n(i, v) : it.p(i, v);
        const double w = it.w(i);
        m[0] += w; m[1] += w * v[0]; m[2] += w * v[1]; m[3] += w * v[2];
    }
    reduce<4>(m, sh);
    const double wsum = m[0] + kEps32;
    const double c[3] = {m[1] / wsum, m[2] / wsum, m[3] / wsum};
    double acc[6] = {0, 0, 0, 0, 0, 0};
    for (int i = threadIdx.x; i < it.N; i += FIT_THREADS) {
        if (!it.member(i)) continue;
        double v[3]; use_normals ? it.n(i, v) : it.p(i, v);
        const double w = it.w(i);
        const double dv[3] = {v[0] - c[0], v[1] - c[1], v[2] - c[2]};
        add_outer(acc, w * w, dv);
    }
    reduce<6>(acc, sh);
    double ev[3];
    smallest_vec(sym(acc), a, ev);
    dval = (a[0] * m[1] + a[1] * m[2] + a[2] * m[3]) / wsum;  // sum w (a.p) / (sum w + eps)
}

// fit_sphere_torch (src/primitive_forward.py:750-773) on the points, optionally projected orthogonally to `axis`
// (the cylinder's circle fit, :806-809).  Returns lstsq status.
__device__ int sphere_fit(const SegIter& it, const double* axis, int nrows, Shared& sh, double c[3], double& radius) {
    auto load = [&](int i, double v[3]) {
        it.p(i, v);
        if (axis) {
            const double t = v[0] * axis[0] + v[1] * axis[1] + v[2] * axis[2];
            v[0] -= t * axis[0]; v[1] -= t * axis[1]; v[2] -= t * axis[2];
        }
    };
    double m[5] = {0, 0, 0, 0, 0};
    for (int i = threadIdx.x; i < it.N; i += FIT_THREADS) {
        if (!it.member(i)) continue;
        double v[3]; load(i, v);
        const double w = it.w(i);
        m[0] += w; m[1] += w * v[0]; m[2] += w * v[1]; m[3] += w * v[2];
        m[4] += w * (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    }
    reduce<5>(m, sh);
    const double wsum = m[0] + kEps32;
    const double mean[3] = {m[1] / wsum, m[2] / wsum, m[3] / wsum};
    const double tq = m[4] / wsum;
    // A_i = w_i * 2 (mean - v_i);  Y_i = w_i * (w_i |v_i|^2 - tq)   (weights enter Y twice, :756-763)
    double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = threadIdx.x; i < it.N; i += FIT_THREADS) {
        if (!it.member(i)) continue;
        double v[3]; load(i, v);
        const double w = it.w(i);
        const double A[3] = {2.0 * w * (mean[0] - v[0]), 2.0 * w * (mean[1] - v[1]), 2.0 * w * (mean[2] - v[2])};
        const double Y = w * (w * (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) - tq);
        acc[0] += A[0] * A[0]; acc[1] += A[0] * A[1]; acc[2] += A[0] * A[2];
        acc[3] += A[1] * A[1]; acc[4] += A[1] * A[2]; acc[5] += A[2] * A[2];
        acc[6] += A[0] * Y; acc[7] += A[1] * Y; acc[8] += A[2] * Y;
    }
    reduce<9>(acc, sh);
    double x[3];
    const int st = lstsq3(sym(acc), acc + 6, nrows, x);
    c[0] = -x[0]; c[1] = -x[1]; c[2] = -x[2];
    double r2[1] = {0};
    for (int i = threadIdx.x; i < it.N; i += FIT_THREADS) {
        if (!it.member(i)) continue;
        double v[3]; load(i, v);
        const double dx = v[0] - c[0], dy = v[1] - c[1], dz = v[2] - c[2];
        r2[0] += it.w(i) * (dx * dx + dy * dy + dz * dz);
    }
    reduce<1>(r2, sh);
    radius = sqrt(fmax(fmax(r2[0] / wsum, 1e-3), 1e-5));  // clamp(min=1e-3) then guard_sqrt (:771-772)
    return st;
}

__global__ void __launch_bounds__(FIT_THREADS) fit_segments_kernel(FitParams p) {
    __shared__ Shared sh;
    const int s = blockIdx.x, b = blockIdx.y;
    const long long base = (long long)b * p.N;
    SegIter it{p.pts + base * 3, p.nrm ? p.nrm + base * 3 : nullptr, p.wts ? p.wts + base : nullptr,
               p.labels ? p.labels + base : nullptr, p.N, s};
    float* out = p.params + ((long long)b * p.S + s) * SED_FIT_PARAMS;
    int* status = p.status + (long long)b * p.S + s;
    const int type = p.seg_type[(long long)b * p.S + s];

    double cnt[1] = {0};
    for (int i = threadIdx.x; i < it.N; i += FIT_THREADS) cnt[0] += it.member(i) ? 1.0 : 0.0;
    reduce<1>(cnt, sh);
    const int nrows = (int)cnt[0];
    const bool analytic = type == SED_PRIM_PLANE || type == SED_PRIM_CONE || type == SED_PRIM_CYLINDER || type == SED_PRIM_SPHERE;
    double res[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int st = 1;
    if (analytic && nrows >= p.min_pts && nrows > 0 && (it.nrm || type == SED_PRIM_PLANE || type == SED_PRIM_SPHERE)) {
        if (type == SED_PRIM_PLANE) {
            double a[3], d;
            plane_fit(it, false, sh, a, d);
            res[0] = a[0]; res[1] = a[1]; res[2] = a[2]; res[3] = d;
            st = 0;
        } else if (type == SED_PRIM_SPHERE) {
            double c[3], r;
            st = sphere_fit(it, nullptr, nrows, sh, c, r);
            res[0] = c[0]; res[1] = c[1]; res[2] = c[2]; res[3] = r;
        } else if (type == SED_PRIM_CYLINDER) {
            // axis = V[:, -1] of svd(w * normals), normalised (:798-804)
            double acc[6] = {0, 0, 0, 0, 0, 0};
            for (int i = threadIdx.x; i < it.N; i += FIT_THREADS) {
                if (!it.member(i)) continue;
                double n[3]; it.n(i, n);
                const double w = it.w(i);
                add_outer(acc, w * w, n);
            }
            reduce<6>(acc, sh);
            double a[3], ev[3];
            smallest_vec(sym(acc), a, ev);
            const double an = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]) + kEps32;
            a[0] /= an; a[1] /= an; a[2] /= an;
            double c[3], r;
            st = sphere_fit(it, a, nrows, sh, c, r);
            res[0] = a[0]; res[1] = a[1]; res[2] = a[2]; res[3] = c[0]; res[4] = c[1]; res[5] = c[2]; res[6] = r;
        } else {  // cone (:812-847)
            double acc[16];
            for (int k = 0; k < 16; ++k) acc[k] = 0.0;
            for (int i = threadIdx.x; i < it.N; i += FIT_THREADS) {
                if (!it.member(i)) continue;
                double n[3], v[3]; it.n(i, n); it.p(i, v);
                const double w = it.w(i);
                double* G = acc;
                G[0] += w * w * n[0] * n[0]; G[1] += w * w * n[0] * n[1]; G[2] += w * w * n[0] * n[2];
                G[3] += w * w * n[1] * n[1]; G[4] += w * w * n[1] * n[2]; G[5] += w * w * n[2] * n[2];
                const double y = w * w * (n[0] * v[0] + n[1] * v[1] + n[2] * v[2]);  // A^T Y = sum (w n)(w n.p)
                acc[6] += y * n[0]; acc[7] += y * n[1]; acc[8] += y * n[2];
                acc[9] += n[0]; acc[10] += n[1]; acc[11] += n[2];  // unweighted sum of normals (sign rule :832)
                acc[12] += w;
            }
            reduce<16>(acc, sh);
            double mu[3], V[3][3];
            eig3(sym(acc), mu, V);
            const bool degenerate = !(mu[2] > 0.0) || sqrt(mu[0] / mu[2]) > 1e5;  // np.linalg.cond(A) > 1e5 (:822)
            if (degenerate) {
                res[3] = 1.0;  // apex 0, axis (1,0,0), theta 0 (:824-827)
                st = 3;
            } else {
                double c[3];
                st = lstsq3(sym(acc), acc + 6, nrows, c);
                double a[3], dummy;
                plane_fit(it, true, sh, a, dummy);  // axis = fit_plane_torch(normals, None, weights)[0] (:831)
                if (acc[9] * a[0] + acc[10] * a[1] + acc[11] * a[2] > 0.0) { a[0] = -a[0]; a[1] = -a[1]; a[2] = -a[2]; }
                double th[1] = {0};
                for (int i = threadIdx.x; i < it.N; i += FIT_THREADS) {
                    if (!it.member(i)) continue;
                    double v[3]; it.p(i, v);
                    v[0] -= c[0]; v[1] -= c[1]; v[2] -= c[2];
                    const double nv = fmax(sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]), 1e-12);
                    double dt = fabs((v[0] * a[0] + v[1] * a[1] + v[2] * a[2]) / nv);
                    dt = fmin(dt, 0.999);
                    th[0] += it.w(i) * acos(dt);
                }
                reduce<1>(th, sh);
                double theta = th[0] / (acc[12] + kEps32);
                theta = fmin(fmax(theta, 1e-3), 3.142 / 2 - 1e-3);
                res[0] = c[0]; res[1] = c[1]; res[2] = c[2]; res[3] = a[0]; res[4] = a[1]; res[5] = a[2]; res[6] = theta;
            }
        }
    }
    if (threadIdx.x < SED_FIT_PARAMS) out[threadIdx.x] = (float)res[threadIdx.x];
    if (threadIdx.x == 0) *status = st;
}

// ------------------------------------------------------------------------------------------------ stage-2 fits
// Fitting_patches_and_edges/primitive_forward_v2.py:716-891 (+ circle_fit_utils.py:11-113): the same moment-matrix
// fits on a CENTRE CROP of the segment -- the m points nearest to the segment's mean, m = int(n * filter_ratio)
// (plane), n // 3 when n > 600 (cylinder), n // 2 (cone) -- with the cylinder's circle found by an algebraic 2-D fit
// in the plane of the projected points and the cone's apex / axis by the stage-2 rules.  The crop is an exact
// selection: a 4 x 8-bit radix select over the FP32 squared distances (shared-memory histogram), ties at the m-th
// distance taken in index order.
struct FitParamsV2 {
    const float* pts; const float* nrm; const float* wts; const long long* labels; const int* seg_type;
    int N, S, min_pts;
    double plane_ratio;
    unsigned char* mask;   // (B,N) scratch
    float* params; int* status;
};

struct CropShared {
    unsigned int hist[256];
    unsigned int prefix, less, want;
    unsigned long long first;
};

__device__ __forceinline__ float crop_d2(const SegIter& it, int i, float cx, float cy, float cz) {
    // (points - center).pow(2).sum(-1) in FP32 (:723-724), no FMA contraction
    const float dx = __fsub_rn(it.pts[3 * i], cx), dy = __fsub_rn(it.pts[3 * i + 1], cy), dz = __fsub_rn(it.pts[3 * i + 2], cz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// Marks in `mask` the m members of the segment nearest to its (unweighted) mean; returns in first_idx the nearest one
// (points[0] of the reference's sorted crop).  All threads of the CTA call.
__device__ void crop_nearest(const SegIter& it, int n, int m, unsigned char* mask, Shared& sh, CropShared& cs, int& first_idx) {
    double c[3] = {0, 0, 0};
    for (int i = threadIdx.x; i < it.N; i += FIT_THREADS) {
        if (!it.member(i)) continue;
        c[0] += it.pts[3 * i]; c[1] += it.pts[3 * i + 1]; c[2] += it.pts[3 * i + 2];
    }
    reduce<3>(c, sh);
    const float cx = (float)(c[0] / n), cy = (float)(c[1] / n), cz = (float)(c[2] / n);
    if (threadIdx.x == 0) { cs.prefix = 0; cs.want = (unsigned)m; cs.less = 0; cs.first = ~0ull; }
    __syncthreads();
    unsigned long long best = ~0ull;
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        for (int i = threadIdx.x; i < 256; i += FIT_THREADS) cs.hist[i] = 0;
        __syncthreads();
        const unsigned prefix = cs.prefix;
        for (int i = threadIdx.x; i < it.N; i += FIT_THREADS) {
            if (!it.member(i)) continue;
            const unsigned key = __float_as_uint(crop_d2(it, i, cx, cy, cz));   // d2 >= 0: integer order = float order
            if (pass == 0) {
                const unsigned long long e = ((unsigned long long)key << 32) | (unsigned)i;
                best = e < best ? e : best;
            }
            if (pass == 0 || (key >> (shift + 8)) == prefix) atomicAdd(&cs.hist[(key >> shift) & 255u], 1u);
        }
        if (pass == 0) atomicMin(&cs.first, best);
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned want = cs.want, bin = 0;
            for (; bin < 256; ++bin) {
                if (cs.hist[bin] >= want) break;
                want -= cs.hist[bin];
            }
            bin = bin > 255u ? 255u : bin;
            cs.prefix = (prefix << 8) | bin;
            cs.want = want;            // rank still to resolve inside the chosen bin
        }
        __syncthreads();
    }
    const unsigned T = cs.prefix;      // bit pattern of the m-th smallest distance
    const unsigned quota = cs.want;    // how many of the entries equal to T belong to the crop
    unsigned ties_l = 0;
    for (int i = threadIdx.x; i < it.N; i += FIT_THREADS) {
        if (!it.member(i)) continue;
        const unsigned key = __float_as_uint(crop_d2(it, i, cx, cy, cz));
        mask[i] = key <= T ? 1 : 0;
        ties_l += key == T ? 1u : 0u;
    }
    double tt[1] = {(double)ties_l};
    reduce<1>(tt, sh);
    if ((unsigned)tt[0] > quota) {     // rare: more equal distances than the crop takes -> the first `quota` in index order
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned seen = 0;
            for (int i = 0; i < it.N; ++i) {
                if (!it.member(i)) continue;
                if (__float_as_uint(crop_d2(it, i, cx, cy, cz)) == T) { if (seen >= quota) mask[i] = 0; ++seen; }
            }
        }
    }
    __syncthreads();
    first_idx = (int)(cs.first & 0xffffffffull);
}

// a^-1 b for a symmetric 3x3 system through its eigen-decomposition, directions below sigma_max * rows * eps dropped
// (the minimum-norm least-squares solution of np.linalg.lstsq / torch.linalg.lstsq on the normal equations)
__device__ void solve_sym3(const Sym3& G, const double g[3], int nrows, double x[3]) {
    double mu[3], V[3][3];
    eig3(G, mu, V);
    const double rt = (double)max(nrows, 3) * kEps32;
    const double tol = mu[0] * rt * rt;   // on eigenvalues = singular values squared
    x[0] = x[1] = x[2] = 0.0;
    for (int i = 0; i < 3; ++i) {
        if (!(mu[i] > tol)) continue;
        const double y = (V[0][i] * g[0] + V[1][i] * g[1] + V[2][i] * g[2]) / mu[i];
        x[0] += V[0][i] * y; x[1] += V[1][i] * y; x[2] += V[2][i] * y;
    }
}

// Rodrigues rotation taking unit vector n0 to n1 applied to v (circle_fit_utils.py:11-28)
__device__ void rodrigues(const double n0[3], const double n1[3], const double v[3], double out[3]) {
    double k[3] = {n0[1] * n1[2] - n0[2] * n1[1], n0[2] * n1[0] - n0[0] * n1[2], n0[0] * n1[1] - n0[1] * n1[0]};
    const double kn = sqrt(k[0] * k[0] + k[1] * k[1] + k[2] * k[2]);
    k[0] /= kn; k[1] /= kn; k[2] /= kn;                       // n0 parallel to n1: 0/0 = NaN, as in the reference
    const double ct = fmin(fmax(n0[0] * n1[0] + n0[1] * n1[1] + n0[2] * n1[2], -1.0), 1.0);
    const double st = sin(acos(ct));
    const double kv = k[0] * v[0] + k[1] * v[1] + k[2] * v[2];
    const double cr[3] = {k[1] * v[2] - k[2] * v[1], k[2] * v[0] - k[0] * v[2], k[0] * v[1] - k[1] * v[0]};
    for (int i = 0; i < 3; ++i) out[i] = v[i] * ct + cr[i] * st + k[i] * kv * (1.0 - ct);
}

__global__ void __launch_bounds__(FIT_THREADS) fit_segments_v2_kernel(FitParamsV2 p) {
    __shared__ Shared sh;
    __shared__ CropShared cs;
    const int s = blockIdx.x, b = blockIdx.y;
    const long long base = (long long)b * p.N;
    SegIter it{p.pts + base * 3, p.nrm ? p.nrm + base * 3 : nullptr, p.wts ? p.wts + base : nullptr,
               p.labels ? p.labels + base : nullptr, p.N, s};
    unsigned char* mask = p.mask + base;
    float* out = p.params + ((long long)b * p.S + s) * SED_FIT_PARAMS;
    int* status = p.status + (long long)b * p.S + s;
    const int type = p.seg_type[(long long)b * p.S + s];

    double cnt[1] = {0};
    for (int i = threadIdx.x; i < it.N; i += FIT_THREADS) cnt[0] += it.member(i) ? 1.0 : 0.0;
    reduce<1>(cnt, sh);
    const int n = (int)cnt[0];
    const bool analytic = type == SED_PRIM_PLANE || type == SED_PRIM_CONE || type == SED_PRIM_CYLINDER || type == SED_PRIM_SPHERE;
    double res[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int st = 1;
    if (analytic && n >= p.min_pts && n > 0 && (it.nrm || type == SED_PRIM_PLANE || type == SED_PRIM_SPHERE)) {
        int m = n, first = 0;
        if (type == SED_PRIM_PLANE && p.plane_ratio > 0.0) m = (int)((double)n * p.plane_ratio);   // :722-725
        if (type == SED_PRIM_CYLINDER && n > 600) m = n / 3;                                          // :830-835
        if (type == SED_PRIM_CONE) m = n / 2;                                                         // :854-856
        m = max(m, 1);
        if (m < n || type == SED_PRIM_CONE) {
            crop_nearest(it, n, m, mask, sh, cs, first);
            it.mask = mask;
        }
        if (type == SED_PRIM_PLANE) {
            double a[3], d;
            plane_fit(it, false, sh, a, d);
            res[0] = a[0]; res[1] = a[1]; res[2] = a[2]; res[3] = d;
            st = 0;
        } else if (type == SED_PRIM_SPHERE) {      // :772-796, the stage-1 algebra
            double c[3], r;
            st = sphere_fit(it, nullptr, n, sh, c, r);
            res[0] = c[0]; res[1] = c[1]; res[2] = c[2]; res[3] = r;
        } else if (type == SED_PRIM_CYLINDER) {
            double acc[6] = {0, 0, 0, 0, 0, 0};
            for (int i = threadIdx.x; i < it.N; i += FIT_THREADS) {
                if (!it.member(i)) continue;
                double nn[3]; it.n(i, nn);
                const double w = it.w(i);
                add_outer(acc, w * w, nn);
            }
            reduce<6>(acc, sh);
            double a[3], ev[3];
            smallest_vec(sym(acc), a, ev);
            const double an = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]) + kEps32;
            a[0] /= an; a[1] /= an; a[2] /= an;
            // circle_segmentation (circle_fit_utils.py:80-113) of prj = p - (p.a) a
            auto prj = [&](int i, double v[3]) {
                it.p(i, v);
                const double t = v[0] * a[0] + v[1] * a[1] + v[2] * a[2];
                v[0] -= t * a[0]; v[1] -= t * a[1]; v[2] -= t * a[2];
            };
            double mm[4] = {0, 0, 0, 0};
            for (int i = threadIdx.x; i < it.N; i += FIT_THREADS) {
                if (!it.member(i)) continue;
                double v[3]; prj(i, v);
                mm[0] += 1.0; mm[1] += v[0]; mm[2] += v[1]; mm[3] += v[2];
            }
            reduce<4>(mm, sh);
            const double mean[3] = {mm[1] / mm[0], mm[2] / mm[0], mm[3] / mm[0]};
            double cov[6] = {0, 0, 0, 0, 0, 0};
            for (int i = threadIdx.x; i < it.N; i += FIT_THREADS) {
                if (!it.member(i)) continue;
                double v[3]; prj(i, v);
                const double dv[3] = {v[0] - mean[0], v[1] - mean[1], v[2] - mean[2]};
                add_outer(cov, 1.0, dv);
            }
            reduce<6>(cov, sh);
            double nv[3];
            smallest_vec(sym(cov), nv, ev);
            const double z[3] = {0.0, 0.0, 1.0};
            // [x y 1] c = x^2 + y^2 in the rotated frame
            double q[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            for (int i = threadIdx.x; i < it.N; i += FIT_THREADS) {
                if (!it.member(i)) continue;
                double v[3], r3[3]; prj(i, v);
                v[0] -= mean[0]; v[1] -= mean[1]; v[2] -= mean[2];
                rodrigues(nv, z, v, r3);
                const double x = r3[0], y = r3[1], bb = x * x + y * y;
                q[0] += x * x; q[1] += x * y; q[2] += x; q[3] += y * y; q[4] += y; q[5] += 1.0;
                q[6] += x * bb; q[7] += y * bb; q[8] += bb;
            }
            reduce<9>(q, sh);
            double cc[3];
            solve_sym3(sym(q), q + 6, m, cc);
            const double xc = cc[0] / 2, yc = cc[1] / 2;
            const double r = sqrt(cc[2] + xc * xc + yc * yc);
            const double c2[3] = {xc, yc, 0.0};
            double C[3];
            rodrigues(z, nv, c2, C);
            res[0] = a[0]; res[1] = a[1]; res[2] = a[2];
            res[3] = C[0] + mean[0]; res[4] = C[1] + mean[1]; res[5] = C[2] + mean[2]; res[6] = r;
            st = 0;
        } else {   // cone :851-891
            double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            for (int i = threadIdx.x; i < it.N; i += FIT_THREADS) {
                if (!it.member(i)) continue;
                double nn[3], v[3]; it.n(i, nn); it.p(i, v);
                add_outer(*reinterpret_cast<double(*)[6]>(acc), 1.0, nn);
                const double y = nn[0] * v[0] + nn[1] * v[1] + nn[2] * v[2];
                acc[6] += y * nn[0]; acc[7] += y * nn[1]; acc[8] += y * nn[2];
            }
            reduce<9>(acc, sh);
            double c[3];
            solve_sym3(sym(acc), acc + 6, m, c);                 // torch.lstsq(Y, A=normals) (:862)
            double a[3], dummy;
            plane_fit(it, false, sh, a, dummy);                  // fit_plane_torch(points, ..., nofilter=True) (:866)
            double p0[3]; it.p(first, p0);
            if ((c[0] - p0[0]) * a[0] + (c[1] - p0[1]) * a[1] + (c[2] - p0[2]) * a[2] < 0.0) { a[0] = -a[0]; a[1] = -a[1]; a[2] = -a[2]; }
            // FP32 compares of the reference (:873-879)
            for (int i = 0; i < 3; ++i) {
                if (fabsf((float)a[i]) >= 0.98f) {
                    const double sg = a[i] > 0 ? 1.0 : -1.0;
                    a[0] = a[1] = a[2] = 0.0;
                    a[i] = sg;
                }
                if (fabsf((float)c[i]) <= 0.1f) c[i] = 0.0;
            }
            double th[2] = {0, 0};
            for (int i = threadIdx.x; i < it.N; i += FIT_THREADS) {
                if (!it.member(i)) continue;
                double v[3]; it.p(i, v);
                v[0] -= c[0]; v[1] -= c[1]; v[2] -= c[2];
                const double nv = fmax(sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]), 1e-12);
                const double dt = fmin(fabs((v[0] * a[0] + v[1] * a[1] + v[2] * a[2]) / nv), 0.999);
                th[0] += it.w(i) * acos(dt); th[1] += it.w(i);
            }
            reduce<2>(th, sh);
            double theta = th[0] / (th[1] + kEps32);
            theta = fmin(fmax(theta, 1e-3), 3.142 / 2 - 1e-3);
            res[0] = c[0]; res[1] = c[1]; res[2] = c[2]; res[3] = a[0]; res[4] = a[1]; res[5] = a[2]; res[6] = theta;
            st = 0;
        }
    }
    if (threadIdx.x < SED_FIT_PARAMS) out[threadIdx.x] = (float)res[threadIdx.x];
    if (threadIdx.x == 0) *status = st;
}

// LeastSquares.lstsq(A, Y) for one (m x 3) system (src/fitting_utils.py:36-65) and the right singular system of an
// (m x 3) matrix (customsvd forward, src/fitting_utils.py:420-455: S descending, V columns).  Single CTA.
__global__ void __launch_bounds__(FIT_THREADS) lstsq3_kernel(const float* __restrict__ A, const float* __restrict__ Y, int m,
                                                             float* __restrict__ x, int* __restrict__ status) {
    __shared__ Shared sh;
    double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = threadIdx.x; i < m; i += FIT_THREADS) {
        const double a[3] = {A[3 * i], A[3 * i + 1], A[3 * i + 2]};
        const double y = Y[i];
        acc[0] += a[0] * a[0]; acc[1] += a[0] * a[1]; acc[2] += a[0] * a[2];
        acc[3] += a[1] * a[1]; acc[4] += a[1] * a[2]; acc[5] += a[2] * a[2];
        acc[6] += a[0] * y; acc[7] += a[1] * y; acc[8] += a[2] * y;
    }
    reduce<9>(acc, sh);
    double sol[3];
    const int st = lstsq3(sym(acc), acc + 6, m, sol);
    if (threadIdx.x < 3) x[threadIdx.x] = (float)sol[threadIdx.x];
    if (threadIdx.x == 0 && status) *status = st;
}

__global__ void __launch_bounds__(FIT_THREADS) svd3_kernel(const float* __restrict__ A, int m, float* __restrict__ S,
                                                           float* __restrict__ V) {
    __shared__ Shared sh;
    double acc[6] = {0, 0, 0, 0, 0, 0};
    for (int i = threadIdx.x; i < m; i += FIT_THREADS) {
        const double a[3] = {A[3 * i], A[3 * i + 1], A[3 * i + 2]};
        add_outer(acc, 1.0, a);
    }
    reduce<6>(acc, sh);
    double w[3], Vd[3][3];
    eig3(sym(acc), w, Vd);
    if (threadIdx.x < 3) S[threadIdx.x] = (float)sqrt(fmax(w[threadIdx.x], 0.0));
    if (threadIdx.x < 9) V[threadIdx.x] = (float)Vd[threadIdx.x / 3][threadIdx.x % 3];
}

// CustomSVD.backward (src/fitting_utils.py:385-417, :449-452): only grad_V flows back,
//     grad_input = 2 U diag(S) sym(K^T o (V^T grad_V)) V^T,   K_ij = 1 / ((S_i - S_j)(S_i + S_j)) off the diagonal, with
//     |S_i - S_j| floored at 1e-6 (svd_grad_K).  The 3 x 3 factor is formed per thread in FP32 in the reference's operation
//     order; every thread then multiplies its rows of U by it.
__global__ void __launch_bounds__(256) svd3_backward_kernel(const float* __restrict__ U, const float* __restrict__ S,
                                                            const float* __restrict__ V, const float* __restrict__ gV, int m,
                                                            float* __restrict__ gin) {
    float s[3], v[3][3], g[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        s[i] = S[i];
#pragma unroll
        for (int j = 0; j < 3; ++j) { v[i][j] = V[3 * i + j]; g[i][j] = gV[3 * i + j]; }
    }
    float K[3][3], inner[3][3], W[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float diff = s[i] - s[j];                                  // s2 - s1: row index first
            const float sgn = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
            float kneg = sgn * fmaxf(fabsf(diff), 1e-6f);
            if (i == j) kneg = 1e-6f;
            K[i][j] = (i == j) ? 0.f : (1.0f / kneg) * (1.0f / (s[i] + s[j]));   // K_neg * K_pos * rm_diag
        }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float vtg = v[0][i] * g[0][j] + v[1][i] * g[1][j] + v[2][i] * g[2][j];   // (V^T grad_V)_ij
            inner[i][j] = K[j][i] * vtg;                                                    // K.T * (...)
        }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            // (S inner_sym V^T)_ij = s_i sum_c sym_ic V_jc
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) acc += ((inner[i][c] + inner[c][i]) / 2.0f) * v[j][c];
            W[i][j] = 2.0f * s[i] * acc;
        }
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < m; r += gridDim.x * blockDim.x) {
        const float u0 = U[3 * r], u1 = U[3 * r + 1], u2 = U[3 * r + 2];
#pragma unroll
        for (int j = 0; j < 3; ++j) gin[3 * r + j] = u0 * W[0][j] + u1 * W[1][j] + u2 * W[2][j];
    }
}

// ---- point -> primitive distances, FP32 in the reference's operation order (src/primitives.py)
__device__ __forceinline__ float guard_sqrtf(float x) { return sqrtf(fmaxf(x, 1e-5f)); }  // src/guard.py:12-14

__device__ float prim_distance(int prim, const float* q, float x, float y, float z) {
    if (prim == SED_PRIM_PLANE) {          // :89-111  (p.a - d)^2
        const float t = x * q[0] + y * q[1] + z * q[2] - q[3];
        return t * t;
    }
    if (prim == SED_PRIM_SPHERE) {         // :113-127 (|p - c| - r)^2
        const float dx = x - q[0], dy = y - q[1], dz = z - q[2];
        const float t = sqrtf(dx * dx + dy * dy + dz * dz) - q[3];
        return t * t;
    }
    if (prim == SED_PRIM_CYLINDER) {       // :129-161
        const float vx = x - q[3], vy = y - q[4], vz = z - q[5];
        const float pr = vx * q[0] + vy * q[1] + vz * q[2];
        const float d2 = fmaxf((vx * vx + vy * vy + vz * vz) - pr * pr, 1e-5f);
        const float t = sqrtf(d2) - q[6];
        return t * t;
    }
    if (prim == SED_PRIM_CONE) {           // :166-195
        const float vx = x - q[0] + 1e-8f, vy = y - q[1] + 1e-8f, vz = z - q[2] + 1e-8f;
        const float mod = sqrtf(vx * vx + vy * vy + vz * vz);
        float al = (vx * q[3] + vy * q[4] + vz * q[5]) / (mod + 1e-7f);
        al = fminf(fmaxf(al, -0.999f), 0.999f);
        const float da = fminf(fabsf(acosf(al) - q[6]), 3.142f / 2.0f);
        const float t = mod * sinf(da);
        return t * t;
    }
    if (prim == 7) {                       // torus :58-87  [axis3 center3 major minor]
        const float an = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
        const float ax = q[0] / an, ay = q[1] / an, az = q[2] / an;
        const float cx = x - q[3], cy = y - q[4], cz = z - q[5];
        const float zz = cx * ax + cy * ay + cz * az;
        const float xx = guard_sqrtf((cx * cx + cy * cy + cz * cz) - zz * zz);
        const float r1 = guard_sqrtf((xx - q[6]) * (xx - q[6]) + zz * zz) - q[7];
        const float r2 = guard_sqrtf((xx + q[6]) * (xx + q[6]) + zz * zz) - q[7];
        return fminf(r1 * r1, r2 * r2);
    }
    return 0.f;
}

__global__ void __launch_bounds__(FIT_THREADS) residual_segments_kernel(const float* __restrict__ pts,
                                                                        const long long* __restrict__ labels,
                                                                        const int* __restrict__ seg_type,
                                                                        const float* __restrict__ params,
                                                                        const int* __restrict__ status, int N, int S,
                                                                        int use_sqrt, float* __restrict__ residual) {
    __shared__ Shared sh;
    __shared__ float q[SED_FIT_PARAMS];
    const int s = blockIdx.x, b = blockIdx.y;
    const long long seg = (long long)b * S + s;
    const int st = status[seg];
    if (st == 1) { if (threadIdx.x == 0) residual[seg] = 0.f; return; }
    if (threadIdx.x < SED_FIT_PARAMS) q[threadIdx.x] = params[seg * SED_FIT_PARAMS + threadIdx.x];
    __syncthreads();
    const int type = seg_type[seg];
    const float* P = pts + (long long)b * N * 3;
    const long long* L = labels ? labels + (long long)b * N : nullptr;
    double acc[2] = {0, 0};
    for (int i = threadIdx.x; i < N; i += FIT_THREADS) {
        if (L && L[i] != (long long)s) continue;
        float dist = prim_distance(type, q, P[3 * i], P[3 * i + 1], P[3 * i + 2]);
        if (use_sqrt) dist = guard_sqrtf(dist);
        acc[0] += (double)dist; acc[1] += 1.0;
    }
    reduce<2>(acc, sh);
    if (threadIdx.x == 0) residual[seg] = acc[1] > 0 ? (float)(acc[0] / acc[1]) : 0.f;
}

__global__ void primitive_distance_kernel(const float* __restrict__ pts, int n, int prim, const float* __restrict__ params,
                                          int use_sqrt, float* __restrict__ out) {
    __shared__ float q[SED_FIT_PARAMS];
    if (threadIdx.x < SED_FIT_PARAMS) q[threadIdx.x] = params[threadIdx.x];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float dist = prim_distance(prim, q, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
    out[i] = use_sqrt ? guard_sqrtf(dist) : dist;
}

// pred_type = argmax_c log_prob[b, c, n] (first maximum, generate_predictions_aug.py:365)
__global__ void pred_type_kernel(const float* __restrict__ lp, int P, int N, int* __restrict__ pred) {
    const int b = blockIdx.y, n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float* x = lp + (long long)b * P * N + n;
    float best = x[0]; int bi = 0;
    for (int c = 1; c < P; ++c) { const float v = x[(long long)c * N]; if (v > best) { best = v; bi = c; } }
    pred[(long long)b * N + n] = bi;
}

// seg_type = mode of pred_type over the segment (np.bincount(...).argmax(): lowest id on ties); one CTA per cloud.
__global__ void __launch_bounds__(1024) segment_vote_kernel(const int* __restrict__ pred, const long long* __restrict__ labels,
                                                            int N, int S, int* __restrict__ seg_type,
                                                            int* __restrict__ seg_count) {
    extern __shared__ int hist[];  // [S][16]
    const int b = blockIdx.x;
    for (int i = threadIdx.x; i < S * 16; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const long long l = labels[(long long)b * N + i];
        const int t = pred[(long long)b * N + i];
        if (l >= 0 && l < S && t >= 0 && t < 16) atomicAdd(&hist[l * 16 + t], 1);
    }
    __syncthreads();
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
        int best = 0, bt = 0, tot = 0;
        for (int t = 0; t < 16; ++t) {
            const int c = hist[s * 16 + t];
            tot += c;
            if (c > best) { best = c; bt = t; }
        }
        seg_type[(long long)b * S + s] = bt;
        if (seg_count) seg_count[(long long)b * S + s] = tot;
    }
}

}  // namespace sed

using namespace sed;

extern "C" {

int sed_fit_segments(const float* points, const float* normals, const float* weights, const int64_t* labels,
                     const int* seg_type, int B, int N, int S, int min_pts, float* params, int* status,
                     sed_stream_t stream) {
    if (!points || !seg_type || !params || !status || B <= 0 || N <= 0 || S <= 0) return SED_ERR_ARG;
    FitParams p{points, normals, weights, (const long long*)labels, seg_type, N, S, min_pts, params, status};
    fit_segments_kernel<<<dim3(S, B), FIT_THREADS, 0, (cudaStream_t)stream>>>(p);
    SED_CHECK_LAUNCH();
    return SED_OK;
}

int sed_fit_segments_v2(const float* points, const float* normals, const float* weights, const int64_t* labels,
                        const int* seg_type, int B, int N, int S, int min_pts, double plane_filter_ratio, float* params,
                        int* status, sed_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!points || !seg_type || !params || !status || B <= 0 || N <= 0 || S <= 0 || !(plane_filter_ratio <= 1.0)) return SED_ERR_ARG;
    ensure_pool_config();
    unsigned char* mask = nullptr;
    SED_CUDA(cudaMallocAsync((void**)&mask, (size_t)B * N, st));
    FitParamsV2 p{points, normals, weights, (const long long*)labels, seg_type, N, S, min_pts, plane_filter_ratio, mask,
                  params, status};
    fit_segments_v2_kernel<<<dim3(S, B), FIT_THREADS, 0, st>>>(p);
    const cudaError_t e = cudaGetLastError();
    ++g_sed_launches;
    cudaFreeAsync(mask, st);
    return e == cudaSuccess ? SED_OK : SED_ERR_CUDA_BASE - (int)e;
}

int sed_lstsq3(const float* A, const float* Y, int m, float* x, int* status, sed_stream_t stream) {
    if (!A || !Y || !x || m <= 0) return SED_ERR_ARG;
    lstsq3_kernel<<<1, FIT_THREADS, 0, (cudaStream_t)stream>>>(A, Y, m, x, status);
    SED_CHECK_LAUNCH();
    return SED_OK;
}

int sed_svd3(const float* A, int m, float* S, float* V, sed_stream_t stream) {
    if (!A || !S || !V || m <= 0) return SED_ERR_ARG;
    svd3_kernel<<<1, FIT_THREADS, 0, (cudaStream_t)stream>>>(A, m, S, V);
    SED_CHECK_LAUNCH();
    return SED_OK;
}

int sed_svd3_backward(const float* U, const float* S, const float* V, const float* grad_V, int m, float* grad_input,
                      sed_stream_t stream) {
    if (!U || !S || !V || !grad_V || !grad_input || m <= 0) return SED_ERR_ARG;
    svd3_backward_kernel<<<std::min((m + 255) / 256, 4 * kNumSMs), 256, 0, (cudaStream_t)stream>>>(U, S, V, grad_V, m, grad_input);
    SED_CHECK_LAUNCH();
    return SED_OK;
}

int sed_residual_segments(const float* points, const int64_t* labels, const int* seg_type, const float* params,
                          const int* status, int B, int N, int S, int use_sqrt, float* residual, sed_stream_t stream) {
    if (!points || !seg_type || !params || !status || !residual || B <= 0 || N <= 0 || S <= 0) return SED_ERR_ARG;
    residual_segments_kernel<<<dim3(S, B), FIT_THREADS, 0, (cudaStream_t)stream>>>(
        points, (const long long*)labels, seg_type, params, status, N, S, use_sqrt, residual);
    SED_CHECK_LAUNCH();
    return SED_OK;
}

int sed_primitive_distance(const float* points, int n, int prim, const float* params, int use_sqrt, float* out,
                           sed_stream_t stream) {
    if (!points || !params || !out || n <= 0) return SED_ERR_ARG;
    if (prim != SED_PRIM_PLANE && prim != SED_PRIM_SPHERE && prim != SED_PRIM_CYLINDER && prim != SED_PRIM_CONE && prim != 7)
        return SED_ERR_ARG;
    primitive_distance_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(points, n, prim, params, use_sqrt, out);
    SED_CHECK_LAUNCH();
    return SED_OK;
}

int sed_segment_types(const float* log_prob, const int64_t* labels, int B, int P, int N, int S, int* pred_type,
                      int* seg_type, int* seg_count, sed_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!pred_type || B <= 0 || P <= 0 || P > 16 || N <= 0) return SED_ERR_ARG;
    if (log_prob) {
        pred_type_kernel<<<dim3((N + 255) / 256, B), 256, 0, st>>>(log_prob, P, N, pred_type);
        SED_CHECK_LAUNCH();
    }
    if (labels && seg_type) {
        if (S <= 0 || (size_t)S * 16 * sizeof(int) > 48 * 1024) return SED_ERR_ARG;
        segment_vote_kernel<<<B, 1024, (size_t)S * 16 * sizeof(int), st>>>(pred_type, (const long long*)labels, N, S,
                                                                          seg_type, seg_count);
        SED_CHECK_LAUNCH();
    }
    return SED_OK;
}

}  // extern "C"
