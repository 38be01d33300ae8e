// Host-side orchestration of the SEDNet forward (reference src/SEDNet.py:78-98 encoder, :292-342 heads) over the
// kernels of knn.cu and pointwise.cu.  No activation tensor of width N*k is ever written: EdgeConv runs as a
// per-point GEMM (U|V) + gather-reduce, GroupNorm is applied in the consumer's load path, the 1024-d global
// feature enters conv1 as a per-cloud bias.
#include "internal.h"

namespace sed {

static constexpr float kGnEps = 1e-5f;

struct EdgeWs {
    float *UV, *ymax, *ymin, *Wf, *a, *s;
    double* stats;
};

static void carve_edge(Arena& A, int B, int N, int Cout, EdgeWs& w) {
    w.UV = A.take<float>((int64_t)B * N * 2 * Cout);
    w.ymax = A.take<float>((int64_t)B * Cout * N);
    w.ymin = A.take<float>((int64_t)B * Cout * N);
    w.Wf = A.take<float>((int64_t)2 * Cout * 128);
    w.a = A.take<float>((int64_t)B * Cout);
    w.s = A.take<float>((int64_t)B * Cout);
    w.stats = A.take<double>((int64_t)B * ((N + 31) / 32) * 32 * 2);
}

// One EdgeConv block given the neighbour table.
static int edgeconv(const float* x, long long x_bstride, const int* idx, const float* W, const float* gamma,
                    const float* beta, int B, int Cin, int Cout, int N, int k, int G, float eps, float slope,
                    float* out, long long out_bstride, const EdgeWs& w, cudaStream_t st) {
    if (Cin > 128) return SED_ERR_UNSUPPORTED;
    SED_TRY(edge_fold_weights(W, Cout, Cin, w.Wf, st));
    SED_TRY(pw_gemm(x, x_bstride, N, w.Wf, Cin, nullptr, 0, nullptr, nullptr, 0, w.UV, (long long)N * 2 * Cout,
                    2 * Cout, 1, nullptr, nullptr, B, Cin, 2 * Cout, N, st));
    SED_TRY(edge_reduce(w.UV, idx, w.ymax, w.ymin, w.stats, B, N, k, Cout, G, st));
    SED_TRY(gn_finalize(w.stats, (N + 31) / 32, G, 1, (double)(Cout / G) * N * k, gamma, beta, B, Cout, G, eps, w.a,
                        w.s, st));
    SED_TRY(edge_finalize(w.ymax, w.ymin, w.a, w.s, B, Cout, N, slope, out, out_bstride, st));
    return SED_OK;
}

struct FwdWs {
    int* idx;
    EdgeWs e;
    float *feats, *x4, *gbias, *Y1, *Y2, *Y3, *Y4, *Y5, *Y6, *TE, *PE, *XS;
    float *a[8], *s[8];
    double* stats;
    float* mm;
};

static void carve_fwd(Arena& A, int B, int N, int k, FwdWs& w) {
    const int P = (N + 127) / 128;
    w.idx = A.take<int>((int64_t)B * N * k);
    carve_edge(A, B, N, 128, w.e);
    w.feats = A.take<float>((int64_t)B * 256 * N);
    w.x4 = A.take<float>((int64_t)B * 1024);
    w.gbias = A.take<float>((int64_t)B * 512);
    w.Y1 = A.take<float>((int64_t)B * 512 * N);
    w.Y2 = A.take<float>((int64_t)B * 256 * N);
    w.Y3 = A.take<float>((int64_t)B * 256 * N);
    w.Y4 = A.take<float>((int64_t)B * 128 * N);
    w.Y5 = A.take<float>((int64_t)B * 256 * N);
    w.Y6 = A.take<float>((int64_t)B * 256 * N);
    w.TE = A.take<float>((int64_t)B * 16 * N);
    w.PE = A.take<float>((int64_t)B * 256 * N);
    w.XS = A.take<float>((int64_t)B * 256 * N);
    for (int i = 0; i < 8; ++i) {
        w.a[i] = A.take<float>((int64_t)B * 1024);
        w.s[i] = A.take<float>((int64_t)B * 1024);
    }
    w.stats = A.take<double>((int64_t)B * P * 32 * 2);
    w.mm = A.take<float>((int64_t)B * P * 1024 * 2);
}

// conv (1x1, bias) + GroupNorm statistics; the normalised tensor is never written, (a, s) are handed to consumers.
static int conv_gn(const float* X, long long x_bstride, const float* W, int ldw, const float* bias,
                   long long bias_bstride, const float* in_a, const float* in_s, int in_act, float* Y, int B, int Cin,
                   int Cout, int N, int G, const float* gamma, const float* beta, float* a, float* s, FwdWs& w,
                   float* mm, cudaStream_t st) {
    const int P = (N + 127) / 128, NBLK = (Cout + 31) / 32;
    SED_TRY(pw_gemm(X, x_bstride, N, W, ldw, bias, bias_bstride, in_a, in_s, in_act, Y, (long long)Cout * N, N, 0,
                    w.stats, mm, B, Cin, Cout, N, st));
    SED_TRY(gn_finalize(w.stats, P, NBLK, (Cout / G) / 32, (double)(Cout / G) * N, gamma, beta, B, Cout, G, kGnEps, a, s,
                        st));
    return SED_OK;
}

// DGCNNEncoderGn.forward, mode 5 (src/SEDNet.py:78-98): leaves x_features in w.feats and x4 in w.x4.
static int encoder(const float* const* P, const float* points, const int* idx1, int B, int N, int k,
                   float normal_metric_W, FwdWs& w, cudaStream_t st) {
    const long long fb = 256LL * N;  // batch stride of feats
    // ---- encoder: three EdgeConv blocks (src/SEDNet.py:80-92)
    // (the first layer's graph depends on the input only: a caller running several networks on the same clouds
    // computes it once with sed_knn_pn and passes it in)
    if (!idx1) SED_TRY(knn_pn(points, 6LL * N, B, N, k, normal_metric_W, w.idx, 0, st, 0));
    SED_TRY(edgeconv(points, 6LL * N, idx1 ? idx1 : w.idx, P[SED_P_ENC_CONV1_W], P[SED_P_ENC_BN1_W], P[SED_P_ENC_BN1_B], B, 6, 64, N,
                     k, 2, kGnEps, 0.2f, w.feats, fb, w.e, st));
    SED_TRY(knn_l2(w.feats, fb, B, 64, N, k, w.idx, 0, st, 0));
    SED_TRY(edgeconv(w.feats, fb, w.idx, P[SED_P_ENC_CONV2_W], P[SED_P_ENC_BN2_W], P[SED_P_ENC_BN2_B], B, 64, 64, N, k,
                     2, kGnEps, 0.2f, w.feats + 64LL * N, fb, w.e, st));
    SED_TRY(knn_l2(w.feats + 64LL * N, fb, B, 64, N, k, w.idx, 0, st, 0));
    SED_TRY(edgeconv(w.feats + 64LL * N, fb, w.idx, P[SED_P_ENC_CONV3_W], P[SED_P_ENC_BN3_W], P[SED_P_ENC_BN3_B], B, 64,
                     128, N, k, 2, kGnEps, 0.2f, w.feats + 128LL * N, fb, w.e, st));

    // ---- mlp1 + GroupNorm(8) + ReLU + max over N (src/SEDNet.py:95-96); the (B,1024,N) tensor is never written
    SED_TRY(conv_gn(w.feats, fb, P[SED_P_ENC_MLP1_W], 256, P[SED_P_ENC_MLP1_B], 0, nullptr, nullptr, 0, nullptr, B, 256,
                    1024, N, 8, P[SED_P_ENC_BNMLP1_W], P[SED_P_ENC_BNMLP1_B], w.a[0], w.s[0], w, w.mm, st));
    SED_TRY(pool_finalize(w.mm, (N + 127) / 128, B, 1024, w.a[0], w.s[0], w.x4, st));

    return SED_OK;
}

}  // namespace sed

using namespace sed;

extern "C" {

int64_t sed_edgeconv_workspace_bytes(int B, int N, int Cout) {
    Arena A(nullptr, 0);
    EdgeWs w;
    carve_edge(A, B, N, Cout, w);
    return A.off;
}

int sed_edgeconv_forward(const float* x, int64_t x_bstride, const int* idx, const float* W, const float* gamma,
                         const float* beta, int B, int Cin, int Cout, int N, int k, int G, float eps, float slope,
                         float* out, int64_t out_bstride, void* workspace, sed_stream_t stream) {
    if (!x || !idx || !W || !gamma || !beta || !out || !workspace) return SED_ERR_ARG;
    if (Cout != 64 && Cout != 128) return SED_ERR_UNSUPPORTED;
    Arena A(workspace, sed_edgeconv_workspace_bytes(B, N, Cout));
    EdgeWs w;
    carve_edge(A, B, N, Cout, w);
    return edgeconv(x, x_bstride, idx, W, gamma, beta, B, Cin, Cout, N, k, G, eps, slope, out, out_bstride, w,
                    (cudaStream_t)stream);
}

int64_t sed_sednet_workspace_bytes(int B, int N, int k) {
    Arena A(nullptr, 0);
    FwdWs w;
    carve_fwd(A, B, N, k, w);
    return A.off;
}

int sed_pointwise_forward(const float* x, int64_t x_bstride, const float* W, int ldw, const float* bias, const float* in_a,
                          const float* in_s, int in_act, float* y, int64_t y_bstride, double* stats, float* mm, int B,
                          int Cin, int Cout, int N, sed_stream_t stream) {
    if (!x || !W || !y || in_act < 0 || in_act > 2 || ldw < Cin) return SED_ERR_ARG;
    return pw_gemm(x, x_bstride, N, W, ldw, bias, 0, in_a, in_s, in_act, y, y_bstride, N, 0, stats, mm, B, Cin, Cout, N,
                   (cudaStream_t)stream);
}

int sed_sednet_forward(const float* const* P, const float* points, int B, int N, int k, float normal_metric_W,
                       float w_pos_enc, int E, int NP, float* embedding, float* log_prob, float* edges, float* x4_out,
                       float* feats_out, void* workspace, int64_t workspace_bytes, sed_stream_t stream) {
    return sed_sednet_forward_g1(P, points, nullptr, B, N, k, normal_metric_W, w_pos_enc, E, NP, embedding, log_prob,
                                 edges, x4_out, feats_out, workspace, workspace_bytes, stream);
}

int sed_sednet_forward_g1(const float* const* P, const float* points, const int* idx1, int B, int N, int k,
                          float normal_metric_W, float w_pos_enc, int E, int NP, float* embedding, float* log_prob,
                          float* edges, float* x4_out, float* feats_out, void* workspace, int64_t workspace_bytes,
                          sed_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!P || !points || !embedding || !log_prob || !edges || !workspace) return SED_ERR_ARG;
    if (B <= 0 || N < k || k <= 0 || k > 256 || E <= 0 || E > 256 || NP <= 0 || NP > 8) return SED_ERR_ARG;
    for (int i = 0; i < SED_P_COUNT; ++i)
        if (!P[i]) return SED_ERR_ARG;
    if (workspace_bytes < sed_sednet_workspace_bytes(B, N, k)) return SED_ERR_ARG;
    Arena A(workspace, workspace_bytes);
    FwdWs w;
    carve_fwd(A, B, N, k, w);
    const long long fb = 256LL * N;  // batch stride of feats
    SED_TRY(encoder(P, points, idx1, B, N, k, normal_metric_W, w, st));

    // ---- conv1 over cat([x4 repeated, feats]) (src/SEDNet.py:300-303): global half as a per-cloud bias
    SED_TRY(gemv_bias(P[SED_P_CONV1_W], 1280, P[SED_P_CONV1_B], w.x4, B, 1024, 512, w.gbias, st));
    SED_TRY(conv_gn(w.feats, fb, P[SED_P_CONV1_W] + 1024, 1280, w.gbias, 512, nullptr, nullptr, 0, w.Y1, B, 256, 512, N,
                    8, P[SED_P_BN1_W], P[SED_P_BN1_B], w.a[1], w.s[1], w, nullptr, st));
    // conv2 (:304) -> x_all = relu(a2*Y2+s2)
    SED_TRY(conv_gn(w.Y1, 512LL * N, P[SED_P_CONV2_W], 512, P[SED_P_CONV2_B], 0, w.a[1], w.s[1], 1, w.Y2, B, 512, 256, N,
                    4, P[SED_P_BN2_W], P[SED_P_BN2_B], w.a[2], w.s[2], w, nullptr, st));
    // type head (:312-314) -> x_type = relu(a3*Y3+s3)
    SED_TRY(conv_gn(w.Y2, 256LL * N, P[SED_P_PRIM1_W], 256, P[SED_P_PRIM1_B], 0, w.a[2], w.s[2], 1, w.Y3, B, 256, 256, N,
                    4, P[SED_P_BN_PRIM1_W], P[SED_P_BN_PRIM1_B], w.a[3], w.s[3], w, nullptr, st));
    const long long teb = 16LL * N;  // TE rows: [0,NP) type logits, [NP,NP+2) edge logits
    SED_TRY(pw_gemm(w.Y3, 256LL * N, N, P[SED_P_PRIM2_W], 256, P[SED_P_PRIM2_B], 0, w.a[3], w.s[3], 1, w.TE, teb, N, 0,
                    nullptr, nullptr, B, 256, NP, N, st));
    SED_TRY(log_softmax(w.TE, teb, B, NP, N, log_prob, st));
    // edge head (:316-317, def :249-253): conv -> GN(4,128) -> conv, no activation
    SED_TRY(conv_gn(w.Y3, 256LL * N, P[SED_P_EDGE0_W], 256, P[SED_P_EDGE0_B], 0, w.a[3], w.s[3], 1, w.Y4, B, 256, 128, N,
                    4, P[SED_P_EDGE1_W], P[SED_P_EDGE1_B], w.a[4], w.s[4], w, nullptr, st));
    SED_TRY(pw_gemm(w.Y4, 128LL * N, N, P[SED_P_EDGE2_W], 128, P[SED_P_EDGE2_B], 0, w.a[4], w.s[4], 0,
                    w.TE + (long long)NP * N, teb, N, 0, nullptr, nullptr, B, 128, 2, N, st));
    SED_CUDA(cudaMemcpy2DAsync(edges, 2LL * N * sizeof(float), w.TE + (long long)NP * N, teb * sizeof(float),
                               2LL * N * sizeof(float), B, cudaMemcpyDeviceToDevice, st));
    // embedding head (:320-329)
    SED_TRY(conv_gn(w.Y2, 256LL * N, P[SED_P_SEG1_W], 256, P[SED_P_SEG1_B], 0, w.a[2], w.s[2], 1, w.Y5, B, 256, 256, N, 4,
                    P[SED_P_BN_SEG1_W], P[SED_P_BN_SEG1_B], w.a[5], w.s[5], w, nullptr, st));
    SED_TRY(conv_gn(w.Y3, 256LL * N, P[SED_P_ASIS0_W], 256, P[SED_P_ASIS0_B], 0, w.a[3], w.s[3], 1, w.Y6, B, 256, 256, N,
                    4, P[SED_P_ASIS1_W], P[SED_P_ASIS1_B], w.a[6], w.s[6], w, nullptr, st));
    SED_TRY(pw_gemm(w.TE, teb, N, P[SED_P_PRIMENC_W], NP + 2, P[SED_P_PRIMENC_B], 0, nullptr, nullptr, 0, w.PE,
                    256LL * N, N, 0, nullptr, nullptr, B, NP + 2, 256, N, st));
    SED_TRY(head_combine(w.Y5, w.a[5], w.s[5], w.Y6, w.a[6], w.s[6], w.PE, w_pos_enc, B, 256, N, w.XS, st));
    SED_TRY(pw_gemm(w.XS, 256LL * N, N, P[SED_P_SEG2_W], 256, P[SED_P_SEG2_B], 0, nullptr, nullptr, 0, embedding,
                    (long long)E * N, N, 0, nullptr, nullptr, B, 256, E, N, st));

    if (x4_out) SED_CUDA(cudaMemcpyAsync(x4_out, w.x4, (size_t)B * 1024 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (feats_out)
        SED_CUDA(cudaMemcpyAsync(feats_out, w.feats, (size_t)B * 256 * N * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return SED_OK;
}

int sed_encoder_forward(const float* const* P, const float* points, int B, int N, int k, float normal_metric_W,
                        float* x4_out, float* feats_out, void* workspace, int64_t workspace_bytes, sed_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!P || !points || !x4_out || !feats_out || !workspace) return SED_ERR_ARG;
    if (B <= 0 || N < k || k <= 0 || k > 256) return SED_ERR_ARG;
    for (int i = 0; i <= SED_P_ENC_BNMLP1_B; ++i)
        if (!P[i]) return SED_ERR_ARG;
    if (workspace_bytes < sed_sednet_workspace_bytes(B, N, k)) return SED_ERR_ARG;
    Arena A(workspace, workspace_bytes);
    FwdWs w;
    carve_fwd(A, B, N, k, w);
    SED_TRY(encoder(P, points, nullptr, B, N, k, normal_metric_W, w, st));
    SED_CUDA(cudaMemcpyAsync(x4_out, w.x4, (size_t)B * 1024 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    SED_CUDA(cudaMemcpyAsync(feats_out, w.feats, (size_t)B * 256 * N * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return SED_OK;
}

}  // extern "C"
