// Cross-translation-unit declarations of the library's internal (C++ linkage) launchers.
#pragma once
#include "../../include/sednet_b200.h"
#include "common.cuh"

namespace sed {

// ---- knn.cu
// sorted = 0: the k neighbours of a row may come in any order (enough for EdgeConv, which reduces over them)
int knn_l2(const float* x, long long bstride, int B, int C, int N, int k, void* idx, int idx64, cudaStream_t st,
           int sorted = 1);
int knn_pn(const float* x6, long long bstride, int B, int N, int k, float W, void* idx, int idx64, cudaStream_t st,
           int sorted = 1);
int knn_far(const float* x, long long bstride, int B, int C, int N, int k, void* idx, int idx64, cudaStream_t st);
int nearest_cos(const float* Q, const float* Cand, int B, int Nq, int Nc, const int* nc_ptr, int d, void* out,
                int idx64, cudaStream_t st);

// ---- pointwise.cu
// Fused 1x1 convolution Y = W . act(in_a * X + in_s) + bias over channel-major activations.
//   X[b*x_bstride + c*ldx + n], Wt[co*ldw + c]; bias nullable, per-cloud when bias_bstride != 0;
//   in_a/in_s (B,Cin) nullable affine applied while loading, in_act 0 none / 1 relu / 2 leaky(0.2);
//   Y nullable; y_point_major: Y[b*y_bstride + n*ldy + co] else Y[b*y_bstride + co*ldy + n];
//   stats nullable: (B, ceil(N/128), ceil(Cout/32), 2) doubles, partial sum / sum of squares per 32 channels;
//   mm nullable: (B, ceil(N/128), Cout, 2) floats, partial max / min over points.
int pw_gemm(const float* X, long long x_bstride, int ldx, const float* Wt, int ldw, const float* bias,
            long long bias_bstride, const float* in_a, const float* in_s, int in_act, float* Y, long long y_bstride,
            int ldy, int y_point_major, double* stats, float* mm, int B, int Cin, int Cout, int N, cudaStream_t st);
// tensor-core implementation (pointwise_tc.cu); SED_ERR_UNSUPPORTED outside its range (Cin < 32)
int pw_gemm_tc(const float* X, long long x_bstride, int ldx, const float* Wt, int ldw, const float* bias,
               long long bias_bstride, const float* in_a, const float* in_s, int in_act, float* Y, long long y_bstride,
               int ldy, int y_point_major, double* stats, float* mm, int B, int Cin, int Cout, int N, cudaStream_t st);
// weight-stationary persistent tensor-core implementation (pointwise_tc2.cu); SED_ERR_UNSUPPORTED outside its range
// (Cin < 32, Cin > 256, Cout < 64)
int pw_gemm_tc2(const float* X, long long x_bstride, int ldx, const float* Wt, int ldw, const float* bias,
                long long bias_bstride, const float* in_a, const float* in_s, int in_act, float* Y, long long y_bstride,
                int ldy, int y_point_major, double* stats, float* mm, int B, int Cin, int Cout, int N, cudaStream_t st);
int gn_finalize(const double* part, int P, int NBLK, int blocks_per_group, double count, const float* gamma,
                const float* beta, int B, int C, int G, float eps, float* a_out, float* s_out, cudaStream_t st);
int edge_fold_weights(const float* W, int Cout, int Cin, float* Wf, cudaStream_t st);
int edge_reduce(const float* UV, const int* idx, float* ymax, float* ymin, double* stats, int B, int N, int k,
                int Cout, int G, cudaStream_t st);
int edge_finalize(const float* ymax, const float* ymin, const float* a, const float* s, int B, int C, int N,
                  float slope, float* out, long long out_bstride, cudaStream_t st);
int pool_finalize(const float* mm, int P, int B, int C, const float* a, const float* s, float* out, cudaStream_t st);
int gemv_bias(const float* Wt, int ldw, const float* bias, const float* v, int B, int Cin, int Cout, float* out,
              cudaStream_t st);
int head_combine(const float* ys, const float* as_, const float* ss, const float* ya, const float* aa,
                 const float* sa, const float* pe, float w, int B, int C, int N, float* out, cudaStream_t st);
int log_softmax(const float* x, long long x_bstride, int B, int C, int N, float* out, cudaStream_t st);

// Stream-ordered scratch comes from the device's default memory pool; keep freed blocks cached across synchronisation
// points (the default release threshold of 0 hands them back to the driver at every sync).
inline void ensure_pool_config() {
    static bool done = false;
    if (done) return;
    int dev = 0;
    cudaMemPool_t pool;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    done = true;
}

inline int64_t align_up(int64_t v, int64_t a = 256) { return (v + a - 1) / a * a; }

// Prepared 1x1-convolution weights (FP16 hi / lo split, per-row power-of-two scale) of a caller whose FP32 weights do not
// change between explicit invalidations -- the pipeline handle (sed_pipeline_set_weights).  While a cache is bound to the
// calling thread, the tensor-core pointwise GEMMs look weight matrices that lie inside one of its registered address ranges up
// by (pointer, pitch, shape, padding) and split them ONCE instead of on every call (round 1 re-split ~24 matrices per
// network and step).  Anything outside the ranges -- workspace buffers such as the folded EdgeConv weights, whose pointer is
// reused for different contents -- takes the stream-ordered temporary path.
struct PwCache;
PwCache* pw_cache_create();
void pw_cache_destroy(PwCache* c);
void pw_cache_clear(PwCache* c);                                      // the weights changed: drop every prepared matrix
void pw_cache_add_range(PwCache* c, const void* lo, size_t bytes);    // an address range of immutable weights
void pw_cache_bind(PwCache* c);                                       // thread-local; nullptr switches the cache off
struct PwCacheScope {                                                 // binds for a scope (early returns included)
    explicit PwCacheScope(PwCache* c) { pw_cache_bind(c); }
    ~PwCacheScope() { pw_cache_bind(nullptr); }
};
// device buffer [Wh | Wl | rscale] (Cout x pad halves, twice, then Cout floats; parts aligned to 256 B) of W (Cout, Cin) with
// row pitch ldw: from the bound cache, or a temporary of stream `st` that the caller releases with cudaFreeAsync
int pw_prepared(const float* W, int ldw, int Cout, int Cin, int pad, cudaStream_t st, char** buf, bool* temporary);

// bump allocator over a caller-provided workspace
struct Arena {
    char* base; int64_t off, cap;
    Arena(void* p, int64_t c) : base((char*)p), off(0), cap(c) {}
    template <typename T> T* take(int64_t n) {
        int64_t o = off;
        off = align_up(off + n * (int64_t)sizeof(T));
        return (base && off <= cap) ? reinterpret_cast<T*>(base + o) : nullptr;
    }
};

}  // namespace sed
