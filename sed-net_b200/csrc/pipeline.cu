// End-to-end inference step behind one C call (reference generate_predictions_aug.py:213-236,365,379-387 followed by
// the analytic fits of Fitting_patches_and_edges/residual_utils.py:210-331): device buffers, weights and workspaces
// live in an opaque handle; sed_pipeline_run_host takes HOST buffers and performs the copies itself.
#include <string.h>

#include <new>

#include "internal.h"

std::atomic<long long> g_sed_launches{0};

namespace sed {

static int64_t param_numel(int i, int E, int NP) {
    switch (i) {
        case SED_P_ENC_CONV1_W: return 64 * 12;
        case SED_P_ENC_BN1_W: case SED_P_ENC_BN1_B: return 64;
        case SED_P_ENC_CONV2_W: return 64 * 128;
        case SED_P_ENC_BN2_W: case SED_P_ENC_BN2_B: return 64;
        case SED_P_ENC_CONV3_W: return 128 * 128;
        case SED_P_ENC_BN3_W: case SED_P_ENC_BN3_B: return 128;
        case SED_P_ENC_MLP1_W: return 1024 * 256;
        case SED_P_ENC_MLP1_B: case SED_P_ENC_BNMLP1_W: case SED_P_ENC_BNMLP1_B: return 1024;
        case SED_P_CONV1_W: return 512 * 1280;
        case SED_P_CONV1_B: case SED_P_BN1_W: case SED_P_BN1_B: return 512;
        case SED_P_CONV2_W: return 256 * 512;
        case SED_P_CONV2_B: case SED_P_BN2_W: case SED_P_BN2_B: return 256;
        case SED_P_PRIM1_W: case SED_P_SEG1_W: case SED_P_ASIS0_W: return 256 * 256;
        case SED_P_PRIM1_B: case SED_P_BN_PRIM1_W: case SED_P_BN_PRIM1_B: return 256;
        case SED_P_PRIM2_W: return NP * 256;
        case SED_P_PRIM2_B: return NP;
        case SED_P_EDGE0_W: return 128 * 256;
        case SED_P_EDGE0_B: case SED_P_EDGE1_W: case SED_P_EDGE1_B: return 128;
        case SED_P_EDGE2_W: return 2 * 128;
        case SED_P_EDGE2_B: return 2;
        case SED_P_SEG1_B: case SED_P_BN_SEG1_W: case SED_P_BN_SEG1_B: return 256;
        case SED_P_ASIS0_B: case SED_P_ASIS1_W: case SED_P_ASIS1_B: return 256;
        case SED_P_PRIMENC_W: return 256 * (NP + 2);
        case SED_P_PRIMENC_B: return 256;
        case SED_P_SEG2_W: return E * 256;
        case SED_P_SEG2_B: return E;
    }
    return 0;
}

// inp[b, 0:3, n] = points[b, n, :], inp[b, 3:6, n] = normals[b, n, :]  (generate_predictions_aug.py:223-225)
__global__ void pack_input_kernel(const float* __restrict__ pts, const float* __restrict__ nrm, int N, float* __restrict__ inp) {
    const int b = blockIdx.y, n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float* p = pts + ((long long)b * N + n) * 3;
    const float* q = nrm + ((long long)b * N + n) * 3;
    float* o = inp + (long long)b * 6 * N + n;
    o[0] = p[0]; o[(long long)N] = p[1]; o[2LL * N] = p[2];
    o[3LL * N] = q[0]; o[4LL * N] = q[1]; o[5LL * N] = q[2];
}

}  // namespace sed

using namespace sed;

struct sed_pipeline {
    int max_B, N, k, S, E, NP, d;   // d: width of the rows of X for the clustering half (128 = the network's embedding)
    int dmax;                       // allocated width of X / shifted / centres (192: room for the 148-column hpnet embedding)
    // weights
    float* wbuf[2];
    const float* wptr[2][SED_P_COUNT];
    bool have_weights;
    PwCache* pw_cache;              // FP16 hi / lo images of the 1x1-convolution weights, split once per set_weights
    // device buffers
    float *pts, *nrm, *inp, *emb, *logp, *edges, *emb2, *logp2, *edges2, *X, *shifted, *tmp, *kth, *bw, *centers, *params,
        *residual;
    long long* labels;
    int *pred_type, *seg_type, *seg_count, *status, *center_ids, *n_centers, *n_labels, *idx1;
    void *fwd_ws, *fwd_ws2, *nms_ws;
    int64_t fwd_ws_bytes;
    // the type network runs on its own stream, concurrently with the instance network (both read the same input and the
    // same first-layer graph; the tails of one network's kernels overlap the other's)
    cudaStream_t side;
    cudaEvent_t ev_fork, ev_join;
    // pinned host scratch
    int* h_counts;  // [2*max_B]: n_labels, n_centers
    // stage boundaries of the last run: start, first-layer graph, both forwards + normalise, bandwidth, shift, nms, fits
    cudaEvent_t ev[7];
    int retries;
};

static int pipe_alloc(void** p, size_t bytes) {
    cudaError_t e = cudaMalloc(p, bytes);
    return e == cudaSuccess ? SED_OK : SED_ERR_CUDA_BASE - (int)e;
}

extern "C" {

int sed_version(void) { return 100; }

const char* sed_error_string(int code) {
    if (code == SED_OK) return "ok";
    if (code == SED_ERR_ARG) return "invalid argument";
    if (code == SED_ERR_UNSUPPORTED) return "shape outside the compiled range";
    if (code == SED_ERR_GUARD) return "guarded mean-shift did not reach <= 49 clusters before K = int(quantile * 10000) exceeded N";
    if (code <= SED_ERR_CUDA_BASE) return cudaGetErrorString((cudaError_t)(-(code - SED_ERR_CUDA_BASE)));
    return "unknown error";
}

int64_t sed_launch_count(int reset) {
    return reset ? g_sed_launches.exchange(0) : g_sed_launches.load();
}

void sed_pipeline_destroy(sed_pipeline_t* p) {
    if (!p) return;
    void* bufs[] = {p->wbuf[0], p->wbuf[1], p->pts, p->nrm, p->inp, p->emb, p->logp, p->edges, p->emb2, p->logp2, p->edges2,
                    p->X, p->shifted, p->tmp, p->kth, p->bw, p->centers, p->params, p->residual, p->labels, p->pred_type,
                    p->seg_type, p->seg_count, p->status, p->center_ids, p->n_centers, p->n_labels, p->idx1, p->fwd_ws,
                    p->fwd_ws2, p->nms_ws};
    for (void* b : bufs)
        if (b) cudaFree(b);
    if (p->h_counts) cudaFreeHost(p->h_counts);
    for (auto& e : p->ev)
        if (e) cudaEventDestroy(e);
    if (p->ev_fork) cudaEventDestroy(p->ev_fork);
    if (p->ev_join) cudaEventDestroy(p->ev_join);
    if (p->side) cudaStreamDestroy(p->side);
    if (p->pw_cache) pw_cache_destroy(p->pw_cache);
    delete p;
}

int sed_pipeline_create(int max_B, int N, int k, int max_segments, sed_pipeline_t** out) {
    if (!out || max_B <= 0 || N <= 0 || k <= 0 || k > N || max_segments <= 0 || max_segments > 512) return SED_ERR_ARG;
    // The driver calls mean_shift(X, 10000, q, ...): compute_bandwidth takes K = int(q * 10000) whatever N is, over all N
    // rows only while N <= 10000 (src/mean_shift.py:122-129 subsamples 10 000 rows with NumPy's RNG above that); the
    // handle runs the all-rows form, so larger clouds are refused instead of silently getting a smaller bandwidth.
    if (N > 10000) return SED_ERR_UNSUPPORTED;
    sed_pipeline* p = new (std::nothrow) sed_pipeline();
    if (!p) return SED_ERR_ARG;
    memset(p, 0, sizeof(*p));
    p->max_B = max_B; p->N = N; p->k = k; p->S = max_segments; p->E = 128; p->NP = 6; p->d = 128; p->dmax = 192;
    const size_t B = max_B, n = N, S = max_segments, d = p->dmax;
    int64_t wtot = 0;
    for (int i = 0; i < SED_P_COUNT; ++i) wtot += align_up(param_numel(i, p->E, p->NP) * 4);
    p->fwd_ws_bytes = sed_sednet_workspace_bytes(max_B, N, k);
    int rc = SED_OK;
#define PALLOC(field, bytes) if (rc == SED_OK) rc = pipe_alloc((void**)&p->field, (bytes))
    PALLOC(wbuf[0], wtot); PALLOC(wbuf[1], wtot);
    PALLOC(pts, B * n * 3 * 4); PALLOC(nrm, B * n * 3 * 4); PALLOC(inp, B * 6 * n * 4);
    PALLOC(emb, B * p->E * n * 4); PALLOC(logp, B * p->NP * n * 4); PALLOC(edges, B * 2 * n * 4);
    PALLOC(emb2, B * p->E * n * 4); PALLOC(logp2, B * p->NP * n * 4); PALLOC(edges2, B * 2 * n * 4);
    PALLOC(idx1, B * n * (size_t)k * 4); PALLOC(fwd_ws2, p->fwd_ws_bytes);
    PALLOC(X, B * n * d * 4); PALLOC(shifted, B * n * d * 4); PALLOC(tmp, B * n * d * 4);
    PALLOC(kth, B * n * 4); PALLOC(bw, B * 4); PALLOC(centers, B * S * d * 4);
    PALLOC(params, B * S * SED_FIT_PARAMS * 4); PALLOC(residual, B * S * 4);
    PALLOC(labels, B * n * 8); PALLOC(pred_type, B * n * 4); PALLOC(seg_type, B * S * 4); PALLOC(seg_count, B * S * 4);
    PALLOC(status, B * S * 4); PALLOC(center_ids, B * S * 4); PALLOC(n_centers, B * 4); PALLOC(n_labels, B * 4);
    PALLOC(fwd_ws, p->fwd_ws_bytes); PALLOC(nms_ws, sed_ms_nms_workspace_bytes(max_B, N));
#undef PALLOC
    if (rc == SED_OK && cudaMallocHost((void**)&p->h_counts, 2 * B * sizeof(int)) != cudaSuccess) rc = SED_ERR_CUDA_BASE - 2;
    for (auto& e : p->ev)
        if (rc == SED_OK && cudaEventCreate(&e) != cudaSuccess) rc = SED_ERR_CUDA_BASE - 2;
    if (rc == SED_OK && cudaStreamCreateWithFlags(&p->side, cudaStreamNonBlocking) != cudaSuccess) rc = SED_ERR_CUDA_BASE - 2;
    if (rc == SED_OK && cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming) != cudaSuccess) rc = SED_ERR_CUDA_BASE - 2;
    if (rc == SED_OK && cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming) != cudaSuccess) rc = SED_ERR_CUDA_BASE - 2;
    if (rc != SED_OK) { sed_pipeline_destroy(p); return rc; }
    p->pw_cache = pw_cache_create();
    pw_cache_add_range(p->pw_cache, p->wbuf[0], (size_t)wtot);
    pw_cache_add_range(p->pw_cache, p->wbuf[1], (size_t)wtot);
    *out = p;
    return SED_OK;
}

int sed_pipeline_set_weights(sed_pipeline_t* p, const float* const* type_params_host, const float* const* inst_params_host) {
    if (!p || !type_params_host || !inst_params_host) return SED_ERR_ARG;
    const float* const* src[2] = {type_params_host, inst_params_host};
    for (int m = 0; m < 2; ++m) {
        int64_t off = 0;
        for (int i = 0; i < SED_P_COUNT; ++i) {
            const int64_t bytes = param_numel(i, p->E, p->NP) * 4;
            if (!src[m][i]) return SED_ERR_ARG;
            float* dst = (float*)((char*)p->wbuf[m] + off);
            SED_CUDA(cudaMemcpy(dst, src[m][i], bytes, cudaMemcpyHostToDevice));
            p->wptr[m][i] = dst;
            off += align_up(bytes);
        }
    }
    pw_cache_clear(p->pw_cache);   // the prepared images belong to the previous weights
    p->have_weights = true;
    return SED_OK;
}

// mean-shift of clouds [b0, b0+nb) with K = int(quantile * 10000) (generate_predictions_aug.py:25-35, src/mean_shift.py:19-43)
static int pipe_mean_shift(sed_pipeline* p, int b0, int nb, double quantile, int iterations, int prec_mode, bool mark,
                           cudaStream_t st) {
    const int64_t N = p->N, d = p->d, S = p->S;
    const int K = (int)(quantile * 10000.0);
    const float* X = p->X + b0 * N * d;
    SED_TRY(sed_ms_bandwidth(X, nb, (int)N, (int)d, K, 0.003f, p->kth + b0 * N, p->bw + b0, st));
    if (mark) SED_CUDA(cudaEventRecord(p->ev[3], st));
    SED_TRY(sed_ms_shift(X, p->bw + b0, nb, (int)N, (int)d, iterations, 0, prec_mode, p->shifted + b0 * N * d,
                         p->tmp + b0 * N * d, st));
    if (mark) SED_CUDA(cudaEventRecord(p->ev[4], st));
    SED_TRY(sed_ms_nms(p->shifted + b0 * N * d, X, p->bw + b0, nb, (int)N, (int)d, (int)S, (int64_t*)p->labels + b0 * N,
                       p->center_ids + b0 * S, p->n_centers + b0, p->n_labels + b0, p->centers + b0 * S * d, p->nms_ws, st));
    if (mark) SED_CUDA(cudaEventRecord(p->ev[5], st));
    return SED_OK;
}

int sed_pipeline_run_device(sed_pipeline_t* p, const float* points_dev, const float* normals_dev, int B, double quantile,
                            int iterations, int prec_mode, sed_stream_t stream) {
    SED_TRY(sed_pipeline_run_forward(p, points_dev, normals_dev, B, stream));
    return sed_pipeline_run_cluster(p, points_dev, normals_dev, B, quantile, iterations, prec_mode, stream);
}

int sed_pipeline_run_forward(sed_pipeline_t* p, const float* points_dev, const float* normals_dev, int B,
                             sed_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!p || !points_dev || !normals_dev || B <= 0 || B > p->max_B || !p->have_weights) return SED_ERR_ARG;
    const int N = p->N, S = p->S;
    PwCacheScope weights_prepared_once(p->pw_cache);
    SED_CUDA(cudaEventRecord(p->ev[0], st));
    pack_input_kernel<<<dim3((N + 255) / 256, B), 256, 0, st>>>(points_dev, normals_dev, N, p->inp);
    SED_CHECK_LAUNCH();
    // first EdgeConv layer's graph: a function of the input only (src/SEDNet.py:80, src/PointNet.py:90-137), shared by
    // the two networks
    SED_TRY(knn_pn(p->inp, 6LL * N, B, N, p->k, 1.0f, p->idx1, 0, st, 0));   // neighbour order is irrelevant to EdgeConv
    SED_CUDA(cudaEventRecord(p->ev[1], st));
    // type network (side stream) and instance network (generate_predictions_aug.py:224-229); only the outputs the
    // driver keeps
    SED_CUDA(cudaEventRecord(p->ev_fork, st));
    SED_CUDA(cudaStreamWaitEvent(p->side, p->ev_fork, 0));
    SED_TRY(sed_sednet_forward_g1(p->wptr[0], p->inp, p->idx1, B, N, p->k, 1.0f, 0.2f, p->E, p->NP, p->emb2, p->logp2,
                                  p->edges2, nullptr, nullptr, p->fwd_ws2, p->fwd_ws_bytes, p->side));
    SED_TRY(sed_segment_types(p->logp2, nullptr, B, p->NP, N, S, p->pred_type, nullptr, nullptr, p->side));
    SED_CUDA(cudaEventRecord(p->ev_join, p->side));
    SED_TRY(sed_sednet_forward_g1(p->wptr[1], p->inp, p->idx1, B, N, p->k, 1.0f, 0.2f, p->E, p->NP, p->emb, p->logp,
                                  p->edges, nullptr, nullptr, p->fwd_ws, p->fwd_ws_bytes, st));
    p->d = p->E;                               // X = the network's own embedding again
    SED_TRY(sed_normalize_transpose(p->emb, B, p->E, N, p->X, st));
    SED_CUDA(cudaStreamWaitEvent(st, p->ev_join, 0));
    SED_CUDA(cudaEventRecord(p->ev[2], st));
    return SED_OK;
}

int sed_pipeline_run_cluster(sed_pipeline_t* p, const float* points_dev, const float* normals_dev, int B, double quantile,
                             int iterations, int prec_mode, sed_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!p || !points_dev || !normals_dev || B <= 0 || B > p->max_B) return SED_ERR_ARG;
    const int N = p->N, S = p->S;
    p->retries = 0;
    // guarded mean-shift: re-run a cloud with quantile * 1.2 while it has more than 49 labels
    SED_TRY(pipe_mean_shift(p, 0, B, (double)quantile, iterations, prec_mode, true, st));
    SED_CUDA(cudaMemcpyAsync(p->h_counts, p->n_labels, B * sizeof(int), cudaMemcpyDeviceToHost, st));
    SED_CUDA(cudaMemcpyAsync(p->h_counts + p->max_B, p->n_centers, B * sizeof(int), cudaMemcpyDeviceToHost, st));
    SED_CUDA(cudaStreamSynchronize(st));
    // The reference's loop is `while True: ...; if unique(labels) > 49: quantile *= 1.2 else break`
    // (generate_predictions_aug.py:25-35); it ends by raising from topk once K = int(quantile * 10000) exceeds N.  Same
    // here: K runs 150, 180, 216, 259, ... up to N (the K-th select has 16-bit counters for K > 255), and a cloud that
    // still has > 49 labels (or more centres than the handle's max_segments) when K would pass N fails the call with
    // SED_ERR_GUARD -- never SED_OK with labels that the type vote / fits cannot index.
    for (int b = 0; b < B; ++b) {
        double q = quantile;
        while (p->h_counts[b] > 49 || p->h_counts[p->max_B + b] < 0) {
            q *= 1.2;
            if ((int)(q * 10000.0) > N) return SED_ERR_GUARD;
            ++p->retries;
            SED_TRY(pipe_mean_shift(p, b, 1, q, iterations, prec_mode, false, st));
            SED_CUDA(cudaMemcpyAsync(p->h_counts + b, p->n_labels + b, sizeof(int), cudaMemcpyDeviceToHost, st));
            SED_CUDA(cudaMemcpyAsync(p->h_counts + p->max_B + b, p->n_centers + b, sizeof(int), cudaMemcpyDeviceToHost, st));
            SED_CUDA(cudaStreamSynchronize(st));
        }
    }
    // per-segment primitive type, fits, residuals
    SED_TRY(sed_segment_types(nullptr, (const int64_t*)p->labels, B, p->NP, N, S, p->pred_type, p->seg_type, p->seg_count,
                              st));
    SED_TRY(sed_fit_segments(points_dev, normals_dev, nullptr, (const int64_t*)p->labels,
                             p->seg_type, B, N, S, 20, p->params, p->status, st));
    SED_TRY(sed_residual_segments(points_dev, (const int64_t*)p->labels, p->seg_type, p->params, p->status, B, N, S, 1,
                                  p->residual, st));
    SED_CUDA(cudaEventRecord(p->ev[6], st));
    return SED_OK;
}

int sed_pipeline_set_cluster_width(sed_pipeline_t* p, int d) {
    if (!p || d <= 0 || d > p->dmax || (d & 3)) return SED_ERR_ARG;
    p->d = d;
    return SED_OK;
}

int sed_pipeline_stage_ms(sed_pipeline_t* p, float* ms6_host, int* retries_host) {
    if (!p || !ms6_host) return SED_ERR_ARG;
    SED_CUDA(cudaEventSynchronize(p->ev[6]));
    for (int i = 0; i < 6; ++i) SED_CUDA(cudaEventElapsedTime(&ms6_host[i], p->ev[i], p->ev[i + 1]));
    if (retries_host) *retries_host = p->retries;
    return SED_OK;
}

int sed_pipeline_run_host(sed_pipeline_t* p, const float* points_host, const float* normals_host, int B, double quantile,
                          int iterations, int prec_mode, int64_t* labels_host, int* pred_type_host, int* seg_type_host,
                          float* params_host, int* status_host, float* residual_host, float* bw_host, int* n_labels_host,
                          sed_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!p || !points_host || !normals_host || B <= 0 || B > p->max_B) return SED_ERR_ARG;
    const size_t N = p->N, S = p->S;
    SED_CUDA(cudaMemcpyAsync(p->pts, points_host, B * N * 3 * 4, cudaMemcpyHostToDevice, st));
    SED_CUDA(cudaMemcpyAsync(p->nrm, normals_host, B * N * 3 * 4, cudaMemcpyHostToDevice, st));
    SED_TRY(sed_pipeline_run_device(p, p->pts, p->nrm, B, quantile, iterations, prec_mode, stream));
    if (labels_host) SED_CUDA(cudaMemcpyAsync(labels_host, p->labels, B * N * 8, cudaMemcpyDeviceToHost, st));
    if (pred_type_host) SED_CUDA(cudaMemcpyAsync(pred_type_host, p->pred_type, B * N * 4, cudaMemcpyDeviceToHost, st));
    if (seg_type_host) SED_CUDA(cudaMemcpyAsync(seg_type_host, p->seg_type, B * S * 4, cudaMemcpyDeviceToHost, st));
    if (params_host) SED_CUDA(cudaMemcpyAsync(params_host, p->params, B * S * SED_FIT_PARAMS * 4, cudaMemcpyDeviceToHost, st));
    if (status_host) SED_CUDA(cudaMemcpyAsync(status_host, p->status, B * S * 4, cudaMemcpyDeviceToHost, st));
    if (residual_host) SED_CUDA(cudaMemcpyAsync(residual_host, p->residual, B * S * 4, cudaMemcpyDeviceToHost, st));
    if (bw_host) SED_CUDA(cudaMemcpyAsync(bw_host, p->bw, B * 4, cudaMemcpyDeviceToHost, st));
    if (n_labels_host) SED_CUDA(cudaMemcpyAsync(n_labels_host, p->n_labels, B * 4, cudaMemcpyDeviceToHost, st));
    SED_CUDA(cudaStreamSynchronize(st));
    return SED_OK;
}

void* sed_pipeline_device_ptr(sed_pipeline_t* p, const char* name) {
    if (!p || !name) return nullptr;
    struct { const char* n; void* v; } tab[] = {
        {"points", p->pts}, {"normals", p->nrm}, {"input", p->inp}, {"embedding", p->emb}, {"log_prob", p->logp},
        {"type_log_prob", p->logp2}, {"graph1", p->idx1},
        {"edges", p->edges}, {"X", p->X}, {"shifted", p->shifted}, {"bw", p->bw}, {"centers", p->centers},
        {"params", p->params}, {"residual", p->residual}, {"labels", p->labels}, {"pred_type", p->pred_type},
        {"seg_type", p->seg_type}, {"seg_count", p->seg_count}, {"status", p->status}, {"center_ids", p->center_ids},
        {"n_centers", p->n_centers}, {"n_labels", p->n_labels}};
    for (auto& t : tab)
        if (!strcmp(t.n, name)) return t.v;
    return nullptr;
}

}  // extern "C"
