/*
 * sednet_b200.h -- C ABI of libsednet_b200.so: the sm_100a (B200) implementation of SED-Net's per-point
 * inference hot path (kNN graph + EdgeConv encoder, per-point heads, mean-shift clustering, primitive fits).
 *
 * The reference (yuanqili78/SED-Net) has no FFI: its boundary for this path is a set of Python functions and
 * nn.Module methods.  Every entry point below names the reference function (file:line under the reference
 * root) it stands in for; the Python mirror of those functions (the .py files of sed-net_b200/src) binds these symbols
 * through ctypes, see INTEGRATION.md.
 *
 * Conventions
 *   - plain C types only; all pointers are DEVICE pointers unless the name ends in _host;
 *   - every function returns int: 0 = ok, SED_ERR_ARG (-1) bad argument, SED_ERR_UNSUPPORTED (-2) shape outside
 *     the compiled range, SED_ERR_GUARD (-3) the guarded mean-shift ran out of quantile (the reference raises there),
 *     -(1000 + cudaError_t) for CUDA failures; never throws, never exits;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); no entry point synchronises the device
 *     unless documented ("host sync");
 *   - tensors are dense row-major in the stated shape, f32 unless stated; indices/labels are int64 where the
 *     reference returns torch int64.
 */
#ifndef SEDNET_B200_H
#define SEDNET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SED_OK 0
#define SED_ERR_ARG (-1)
#define SED_ERR_UNSUPPORTED (-2)
#define SED_ERR_GUARD (-3)
#define SED_ERR_CUDA_BASE (-1000)

typedef void* sed_stream_t; /* cudaStream_t */

/* library / device info: returns the ABI version (major*100+minor). */
int sed_version(void);
/* human-readable text for a return code (static storage). */
const char* sed_error_string(int code);

/* ------------------------------------------------------------------ graph ops (src/PointNet.py) */

/* src/PointNet.py:62-87  knn(x, k, k): x (B,C,N) -> idx (B,N,k), nearest first (self first).
 * idx64 != 0 -> int64 output (torch convention), else int32.  C <= 128, k <= 256, k <= N. */
int sed_knn_l2(const float* x, int B, int C, int N, int k, void* idx, int idx64, sed_stream_t stream);

/* src/PointNet.py:90-137  knn_points_normals(x, k, k, W): x (B,6,N) rows 0-2 points, 3-5 normals;
 * metric p_dist * (1 + W * (2 - 2 n.n')). */
int sed_knn_pn(const float* x6, int B, int N, int k, float W, void* idx, int idx64, sed_stream_t stream);

/* src/PointNet.py:140-171 / :174-208  get_graph_feature[_with_normals]: gathers cat([x_j - x_i, x_i]).
 * x (B,C,N), idx (B,N,k) int64 -> out (B,2C,N,k).  (API parity only: the fused forward never materialises this.) */
int sed_graph_feature(const float* x, const int64_t* idx, int B, int C, int N, int k, float* out,
                      sed_stream_t stream);

/* ------------------------------------------------------------------ network (src/SEDNet.py) */

/* Parameter table of SEDNet with the inference driver's kwargs (generate_predictions_aug.py:142-154): device
 * pointers to the state_dict tensors, in this order (src/SEDNet.py:31-48, :228-290). */
enum sed_param {
    SED_P_ENC_CONV1_W = 0, /* encoder.conv1.0.weight (64,12)   */
    SED_P_ENC_BN1_W,       /* encoder.bn1.weight (64)          */
    SED_P_ENC_BN1_B,
    SED_P_ENC_CONV2_W,     /* encoder.conv2.0.weight (64,128)  */
    SED_P_ENC_BN2_W,
    SED_P_ENC_BN2_B,
    SED_P_ENC_CONV3_W,     /* encoder.conv3.0.weight (128,128) */
    SED_P_ENC_BN3_W,
    SED_P_ENC_BN3_B,
    SED_P_ENC_MLP1_W,      /* encoder.mlp1.weight (1024,256)   */
    SED_P_ENC_MLP1_B,
    SED_P_ENC_BNMLP1_W,
    SED_P_ENC_BNMLP1_B,
    SED_P_CONV1_W,         /* conv1.weight (512,1280)          */
    SED_P_CONV1_B,
    SED_P_BN1_W,
    SED_P_BN1_B,
    SED_P_CONV2_W,         /* conv2.weight (256,512)           */
    SED_P_CONV2_B,
    SED_P_BN2_W,
    SED_P_BN2_B,
    SED_P_PRIM1_W,         /* mlp_prim_prob1.weight (256,256)  */
    SED_P_PRIM1_B,
    SED_P_BN_PRIM1_W,
    SED_P_BN_PRIM1_B,
    SED_P_PRIM2_W,         /* mlp_prim_prob2.weight (P,256)    */
    SED_P_PRIM2_B,
    SED_P_EDGE0_W,         /* edge_module.0.weight (128,256)   */
    SED_P_EDGE0_B,
    SED_P_EDGE1_W,         /* edge_module.1 GroupNorm(4,128)   */
    SED_P_EDGE1_B,
    SED_P_EDGE2_W,         /* edge_module.2.weight (2,128)     */
    SED_P_EDGE2_B,
    SED_P_SEG1_W,          /* mlp_seg_prob1.weight (256,256)   */
    SED_P_SEG1_B,
    SED_P_BN_SEG1_W,
    SED_P_BN_SEG1_B,
    SED_P_ASIS0_W,         /* asis.0.weight (256,256)          */
    SED_P_ASIS0_B,
    SED_P_ASIS1_W,
    SED_P_ASIS1_B,
    SED_P_PRIMENC_W,       /* prim_encoding.0.weight (256,P+2) */
    SED_P_PRIMENC_B,
    SED_P_SEG2_W,          /* mlp_seg_prob2.weight (E,256)     */
    SED_P_SEG2_B,
    SED_P_COUNT
};

/* bytes of scratch sed_sednet_forward needs for (B, N, k). */
int64_t sed_sednet_workspace_bytes(int B, int N, int k);

/* src/SEDNet.py:292-342  SEDNet.forward(points, None, False) with mode=5, primitives, embedding,
 * combine_label_prim, edge_module, late_fusion (the driver's configuration), including the encoder
 * DGCNNEncoderGn.forward src/SEDNet.py:78-98.
 *   params      SED_P_COUNT device pointers (host array of pointers)
 *   points      (B,6,N)
 *   k           nn_nb;  normal_metric_W, w_pos_enc as in the constructor;  emb_size E (<=128), num_primitives P (<=8)
 *   embedding   (B,E,N), log_prob (B,P,N), edges (B,2,N)   outputs
 *   x4 (B,1024), x_features (B,256,N)  optional encoder outputs (may be NULL)
 *   workspace   sed_sednet_workspace_bytes(B,N,k) bytes */
int sed_sednet_forward(const float* const* params_host, const float* points, int B, int N, int k,
                       float normal_metric_W, float w_pos_enc, int emb_size, int num_primitives, float* embedding,
                       float* log_prob, float* edges, float* x4, float* x_features, void* workspace,
                       int64_t workspace_bytes, sed_stream_t stream);

/* Same forward with the first EdgeConv layer's neighbour table given: idx1 (B,N,k) int32 = sed_knn_pn(points, ..., k,
 * normal_metric_W, idx64 = 0).  That table depends on the input only, so a driver that runs several networks on the
 * same clouds (type net and instance net, generate_predictions_aug.py:224-229) builds it once.  idx1 NULL = compute. */
int sed_sednet_forward_g1(const float* const* params_host, const float* points, const int* idx1, int B, int N, int k,
                          float normal_metric_W, float w_pos_enc, int emb_size, int num_primitives, float* embedding,
                          float* log_prob, float* edges, float* x4, float* x_features, void* workspace,
                          int64_t workspace_bytes, sed_stream_t stream);

/* src/SEDNet.py:78-98  DGCNNEncoderGn.forward (mode 5) alone: points (B,6,N) -> x4 (B,1024), x_features (B,256,N).
 * Only the encoder entries of the parameter table (SED_P_ENC_*) are read.  Workspace as for sed_sednet_forward. */
int sed_encoder_forward(const float* const* params_host, const float* points, int B, int N, int k,
                        float normal_metric_W, float* x4, float* x_features, void* workspace, int64_t workspace_bytes,
                        sed_stream_t stream);

/* One EdgeConv block, src/SEDNet.py:37-45,81-92: Conv2d(2C->Cout,1x1,no bias) over cat([x_j-x_i, x_i]) ->
 * GroupNorm(G) -> LeakyReLU(slope) -> max over k.  x (B,Cin,N) with batch stride x_bstride elements,
 * idx (B,N,k) int32, W (Cout,2Cin), out (B,Cout,N) with batch stride out_bstride.  Cout in {64,128}.
 * workspace: sed_edgeconv_workspace_bytes(B,N,Cout). */
int64_t sed_edgeconv_workspace_bytes(int B, int N, int Cout);
int sed_edgeconv_forward(const float* x, int64_t x_bstride, const int* idx, const float* W, const float* gamma,
                         const float* beta, int B, int Cin, int Cout, int N, int k, int G, float eps, float slope,
                         float* out, int64_t out_bstride, void* workspace, sed_stream_t stream);

/* One 1x1 convolution of the network, src/SEDNet.py:292-342 (torch.nn.Conv1d(Cin, Cout, 1) applied to the
 * GroupNorm + activation of the previous layer):  y[b,co,n] = bias[co] + sum_c W[co,c] * act(in_a[b,c] * x[b,c,n] + in_s[b,c]).
 * x (B,Cin,N) with batch stride x_bstride elements, W (Cout,Cin) row pitch ldw, bias (Cout) or NULL, in_a / in_s (B,Cin)
 * the folded GroupNorm scale / shift of the input (both NULL: x is used as is), in_act 0 none, 1 ReLU, 2 LeakyReLU(0.2).
 * y (B,Cout,N) with batch stride y_bstride.  Optional epilogue outputs: stats (B, ceil(N/128), ceil(Cout/32), 2) doubles =
 * per 128-point tile and 32-channel block the sum and sum of squares of y; mm (B, ceil(N/128), Cout, 2) = max and min. */
int sed_pointwise_forward(const float* x, int64_t x_bstride, const float* W, int ldw, const float* bias, const float* in_a,
                          const float* in_s, int in_act, float* y, int64_t y_bstride, double* stats, float* mm, int B,
                          int Cin, int Cout, int N, sed_stream_t stream);

/* ------------------------------------------------------------------ mean-shift (src/mean_shift.py) */

/* F.normalize(embedding[b].T, p=2, dim=1) (generate_predictions_aug.py:379-380): emb (B,d,N) -> X (B,N,d). */
int sed_normalize_transpose(const float* emb, int B, int d, int N, float* X, sed_stream_t stream);

/* src/mean_shift.py:115-137 compute_bandwidth on the rows given (+ the clamp(min=min_bw) of :34):
 * X (B,N,d) unit rows, K = int(quantile*num_samples) -> bw (B).  kth_ws: B*N floats scratch. */
int sed_ms_bandwidth(const float* X, int B, int N, int d, int K, float min_bw, float* kth_ws, float* bw,
                     sed_stream_t stream);

/* src/mean_shift.py:45-79 mean_shift_: `iterations` fixed shifts on the unit hypersphere.
 *   X (B,N,d), d <= 256 a multiple of 4, bw (B) device, kernel_type 0 gaussian / 1 epanechnikov.
 *   prec_mode 0 = FP32 FFMA in the reference's operation order (any d <= 256).  Tensor-core modes (tcgen05, FP16 hi/lo
 *   split operands, FP32 accumulation; rows zero-padded to 128 or 192 columns):
 *     4 = S from 3 MMAs (Qh.Xh + Qh.Xl + Ql.Xh), O from 3 (Ph.Xh + Ph.Xl + Pl.Xh): every operand of both legs carries
 *         22 significant bits -- FP32-faithful (d <= 128; for 128 < d <= 256 the call runs mode 0);
 *     1 = S from 3 MMAs, O from 2 (Ph.Xh + Ph.Xl): scores and X at 22 bits, the weights P are single FP16 values (11 bits,
 *         unbiased rounding: ~2e-6 per iteration and point against FP32, <= 1e-4 on BASELINE's configs; d <= 128;
 *         for 128 < d <= 256 the call runs mode 0, which is at least as accurate);
 *     3 = S from 3 MMAs, O from 1 (Ph.Xh): the exponent is FP32-faithful, the weighted mean carries one FP16 rounding of
 *         X per term (d <= 192; wider rows run mode 0);
 *     2 = single FP16 pass for both legs (d <= 128; fast, not FP32-faithful); for 128 < d <= 192 it runs as mode 3.
 *   out (B,N,d); tmp (B,N,d) scratch (ping-pong). */
int sed_ms_shift(const float* X, const float* bw, int B, int N, int d, int iterations, int kernel_type,
                 int prec_mode, float* out, float* tmp, sed_stream_t stream);
/* The same iterations started from positions Q0 (B,N,d) instead of from the keys themselves (Q0 NULL or == X: sed_ms_shift):
 * one step of a longer run -- the training path calls it with iterations = 1 and keeps every state for
 * sed_ms_shift_backward_step.  Rows of 129..192 columns take the FP32 FFMA kernel here. */
int sed_ms_shift_from(const float* X, const float* Q0, const float* bw, int B, int N, int d, int iterations, int kernel_type,
                      int prec_mode, float* out, float* tmp, sed_stream_t stream);

/* Backward of ONE mean-shift iteration (Gaussian kernel) -- the training path: src/segment_loss.py:50-56 runs
 * mean_shift(..., iterations = 5, nms = False) inside the triplet loss and differentiates src/mean_shift.py:45-79 with
 * autograd.  Q (B,N,d) are the positions that entered the iteration, X (B,N,d) the keys, G (B,N,d) = dL/d(positions after the
 * iteration), bw (B) the bandwidths (constants of the graph, as in the reference: computed under no_grad).  Writes
 * dQ (B,N,d) = dL/dQ and ADDS the iteration's dL/dX to dX_accum (B,N,d); the caller adds the final dQ (iteration 0 starts
 * from Q = X).  d <= 128, a multiple of 4.  P and dS are recomputed tile by tile; workspace:
 * sed_ms_shift_backward_workspace_bytes(B, N). */
int64_t sed_ms_shift_backward_workspace_bytes(int B, int N);
int sed_ms_shift_backward_step(const float* Q, const float* X, const float* G, const float* bw, int B, int N, int d,
                               float* dQ, float* dX_accum, void* workspace, sed_stream_t stream);

/* src/mean_shift.py:139-179 nms: centers = shifted points (B,N,d), X (B,N,d), bw (B).
 *   labels (B,N) int64; center_ids (B,max_centers) int32 ascending, n_centers (B) int32, n_labels (B) int32
 *   (number of distinct labels in use = torch.unique(labels).shape[0]); centers_out (B,max_centers,d).
 *   workspace: sed_ms_nms_workspace_bytes(B,N). Clouds with more than max_centers centres get n_centers = -1. */
int64_t sed_ms_nms_workspace_bytes(int B, int N);
int sed_ms_nms(const float* centers, const float* X, const float* bw, int B, int N, int d, int max_centers,
               int64_t* labels, int* center_ids, int* n_centers, int* n_labels, float* centers_out,
               void* workspace, sed_stream_t stream);

/* src/segment_utils.py:536-545 to_one_hot(target, maxx): labels (N) int64 -> (N,maxx) f32. */
int sed_one_hot(const int64_t* labels, int N, int maxx, float* out, sed_stream_t stream);

/* ------------------------------------------------------------------ primitive fits (src/primitive_forward.py) */

#define SED_PRIM_PLANE 1    /* src/primitive_forward.py:1006 */
#define SED_PRIM_CONE 3     /* :1011 */
#define SED_PRIM_CYLINDER 4 /* :1016 */
#define SED_PRIM_SPHERE 5   /* :1021 */
#define SED_FIT_PARAMS 8    /* floats per segment */

/* Batched weighted least-squares fits, one CTA per (cloud, segment):
 *   Fit.fit_plane_torch    src/primitive_forward.py:712-733 -> params [a0 a1 a2 d]
 *   Fit.fit_sphere_torch   :750-773                          -> [c0 c1 c2 r]
 *   Fit.fit_cylinder_torch :788-810                          -> [a0 a1 a2 c0 c1 c2 r]
 *   Fit.fit_cone_torch     :812-847                          -> [c0 c1 c2 a0 a1 a2 theta]
 *   with LeastSquares.lstsq / best_lambda src/fitting_utils.py:32-85 and the dispatch + 20-point minimum of
 *   fit_one_shape_torch :929-1051.
 * points, normals (B,N,3); weights (B,N) or NULL (= 1); labels (B,N) int64 or NULL (every point belongs to
 * segment 0); seg_type (B,S) int32; params (B,S,8); status (B,S) int32: 0 fitted (full-rank), 1 skipped,
 * 2 fitted through the regularised branch, 3 degenerate cone (cond > 1e5, :822-827). */
int sed_fit_segments(const float* points, const float* normals, const float* weights, const int64_t* labels,
                     const int* seg_type, int B, int N, int S, int min_pts, float* params, int* status,
                     sed_stream_t stream);

/* ------------------------------------------------------------------ stage 2 (Fitting_patches_and_edges/) */

/* Stage-2 variants of the four fits, Fitting_patches_and_edges/primitive_forward_v2.py:716-891 with
 * circle_fit_utils.py:11-113, same argument and output layout as sed_fit_segments (type ids SED_PRIM_*; the stage-2
 * dispatcher's own ids 1 plane / 3 cone / 2 cylinder / 4 sphere, :953-976, are mapped by the caller):
 *   plane     fit on the int(n * plane_filter_ratio) points nearest to the segment mean (ratio <= 0: no crop)
 *   sphere    as stage 1
 *   cylinder  axis from the SVD of the weighted normals (nearest n // 3 points when n > 600), circle by an algebraic
 *             2-D fit in the plane of the projected points -> [a0 a1 a2 c0 c1 c2 r]
 *   cone      nearest n // 2 points; apex from the unweighted least squares normals.c = normals.p, axis from the plane
 *             fit of the points, oriented and snapped by the rules of :868-879 -> [c0 c1 c2 a0 a1 a2 theta]
 * status: 0 fitted, 1 skipped. */
int sed_fit_segments_v2(const float* points, const float* normals, const float* weights, const int64_t* labels,
                        const int* seg_type, int B, int N, int S, int min_pts, double plane_filter_ratio, float* params,
                        int* status, sed_stream_t stream);

/* pointnet2 three_nn, Fitting_patches_and_edges/pointnet2/_ext_src/src/interpolate_gpu.cu:14-66: for every row of
 * unknown (B,n,3) the three nearest rows of known (B,m,3): dist2 (B,n,3) squared distances ascending, idx (B,n,3) int32;
 * equal distances keep the lowest index. */
int sed_three_nn(const float* unknown, const float* known, int B, int n, int m, float* dist2, int* idx,
                 sed_stream_t stream);

/* get_edges_between_insts, Fitting_patches_and_edges/proj_2_edge_utils.py:45-60: idx3 (n,3) = three_nn of the cloud on
 * itself, insts (n) int64 -> out (n) uint8: first (strict: and second) non-self neighbour belongs to another instance. */
int sed_inst_edges(const int* idx3, const int64_t* insts, int n, int strict, uint8_t* out, sed_stream_t stream);

/* face_face_inter_map, proj_2_edge_utils.py:63-110: mat (30,30) uint8, mat[a][b] = 1 when at least nn_num_thresh first /
 * second neighbours of instance a's points belong to instance b; an instance of primitive_ids (n_ids, int64) without
 * any neighbour gets the instance of the point nearest to its first point. */
int sed_face_face_map(const float* points, const int64_t* insts, const int* idx3, const int64_t* primitive_ids, int n_ids,
                      int n, int nn_num_thresh, uint8_t* mat, sed_stream_t stream);

/* ------------------------------------------------------------------ HPNet-style embedding weights */

/* compute_entropy, src/smooth_normal_matrix.py:95-154: features (N,K) f32 -> out3 (device, 3 floats) =
 * { E, average_dst, alpha }.  Only the first 5 * chunk points enter the pairwise sums (the reference's ITER = 5 chunks of
 * CHUNK rows), which are divided by N * N as in the reference.  No host synchronisation. */
int sed_compute_entropy(const float* features, int N, int K, int chunk, float* out3, sed_stream_t stream);

/* knn_idx, src/smooth_normal_matrix.py:31-39: xyz (B,N,3) -> idx (B,N,k) int32, the k LARGEST entries of every row of the
 * squared-distance matrix (-2 x.y + |x|^2 + |y|^2), largest first -- torch.topk's default largest=True makes the
 * reference's "knn" the k FARTHEST points; kept as is.  k <= 256. */
int sed_far_idx(const float* xyz, int B, int N, int k, int* idx, sed_stream_t stream);

/* construction_affinity_matrix_normal, src/smooth_normal_matrix.py:42-92, in factored form (the reference's dense (N,N)
 * matrix is A = (M + M^T) / 2, M_ij = a_ij / sqrt(D_i D_j), a_ij = w_ij at the k scattered entries of row i where
 * w_ij != 0 and 1e-12 everywhere else, D_i = sum_j a_ij):  normals (B,N,3), idx (B,N,k) from sed_far_idx ->
 * w (B,N,k) = exp(-acos(clamp(n_i.n_j, -0.99, 0.99))^2 / (2 sigma^2)), dinv (B,N) = 1 / sqrt(D_i). */
int sed_affinity_normal_build(const float* normals, const int* idx, int B, int N, int k, float sigma, float* w, float* dinv,
                              sed_stream_t stream);

/* Block product with that matrix for ONE cloud (what torch.lobpcg needs from it, :191): prepare builds the transposed
 * adjacency in `workspace` (sed_affinity_workspace_bytes(N,k) bytes) once; matmul computes Y (N,m) = A X for X (N,m)
 * row-major, m <= 64.  idx (N,k), w (N,k), dinv (N) of that cloud. */
int64_t sed_affinity_workspace_bytes(int N, int k);
int sed_affinity_prepare(const int* idx, const float* w, int N, int k, void* workspace, sed_stream_t stream);
int sed_affinity_matmul(const int* idx, const float* w, const float* dinv, void* workspace, const float* X, int N, int k, int m,
                        float* Y, sed_stream_t stream);

/* ------------------------------------------------------------------ evaluation (src/segment_utils.py, src/utils.py) */

/* The integer tables from which the driver's matching / IoU metrics are computed (SIOU_matched_segments[_usecd]
 * src/segment_utils.py:140-243, mean_IOU_primitive_segment[_usecd] :359-495, relaxed_iou_fast :609-627,
 * compute_type_miou_abc :300-357), one pass over the points.  pred, gt (B,N) int64 labels (values outside [0,K) are
 * ignored, e.g. the -1 background of :326-331); type_pred, type_gt (B,N) int64 in [0,T) or NULL.  K <= 64, T <= 16.
 *   confusion (B,K,K) [r][c] = #points with pred r and gt c;   npred, ngt (B,K) segment sizes;
 *   pred_types, gt_types (B,K,T) histogram of the point types of every segment;
 *   gt_first (B,K) index of the first point of every gt segment (N when empty: `gt_prim[gt_indices][0]`, :411,478). */
int sed_segment_tables(const int64_t* pred, const int64_t* gt, const int64_t* type_pred, const int64_t* type_gt, int B, int N,
                       int K, int T, int* confusion, int* npred, int* ngt, int* pred_types, int* gt_types, int* gt_first,
                       sed_stream_t stream);

/* primitive_type_segment_torch, src/segment_utils.py:509-517: out (T,K)[l][k] = sum_n [types[n] == l] * weights[n,k]
 * (types (N) int64, weights (N,K) f32; the caller takes the argmax over l).  T <= 16. */
int sed_type_vote_weighted(const int64_t* types, const float* weights, int N, int K, int T, float* out, sed_stream_t stream);

/* chamfer_distance, src/utils.py:273-296: a (B,n,3), b (B,m,3) -> min_a (B,n) = min_j |a_i - b_j|^2, min_b (B,m) (FP32,
 * the reference's operation order); the caller averages ((mean min_a + mean min_b) / 2). */
int sed_chamfer_min(const float* a, const float* b, int B, int n, int m, float* min_a, float* min_b, sed_stream_t stream);

/* The reference's chamfer extension, src/chamfer_distance/chamfer_distance.cu:6-158 (forward_cuda) and :161-187
 * (backward_cuda), bound by src/chamfer_distance/chamfer_distance.py:44-75: xyz1 (B,n,3), xyz2 (B,m,3) ->
 * dist1 (B,n) / idx1 (B,n) int32 = squared distance to / index of the nearest point of xyz2 (lowest index on ties), dist2 /
 * idx2 (B,m) the other way round; backward: grad_xyz1 (B,n,3), grad_xyz2 (B,m,3) from grad_dist1 / grad_dist2 (zeroed here). */
int sed_chamfer_forward(const float* xyz1, const float* xyz2, int B, int n, int m, float* dist1, float* dist2, int* idx1,
                        int* idx2, sed_stream_t stream);
int sed_chamfer_backward(const float* xyz1, const float* xyz2, const float* grad_dist1, const float* grad_dist2, const int* idx1,
                         const int* idx2, int B, int n, int m, float* grad_xyz1, float* grad_xyz2, sed_stream_t stream);

/* The chamfer terms of ALL matched segment pairs of mean_IOU_primitive_segment_usecd (src/segment_utils.py:473) in one
 * pass: pred2gt, gt2pred (B,K) int32 matched partner of every segment or -1.  min_pred (B,N)[i] = distance^2 from point
 * i to the nearest point of the gt segment matched to i's predicted segment; min_gt (B,N)[i] = to the nearest point of
 * the predicted segment matched to i's gt segment; +inf when unmatched.  K <= 254. */
int sed_matched_chamfer(const float* points, const int64_t* pred, const int64_t* gt, const int* pred2gt, const int* gt2pred, int B,
                        int N, int K, float* min_pred, float* min_gt, sed_stream_t stream);

/* LeastSquares.lstsq(A, Y, lamb) src/fitting_utils.py:36-65 for one (m,3) system: x (3); status 0 full rank (QR
 * branch), 2 regularised branch (best_lambda :68-85). */
int sed_lstsq3(const float* A, const float* Y, int m, float* x, int* status, sed_stream_t stream);

/* customsvd forward (src/fitting_utils.py:420-455, torch.svd(some=True)) of an (m,3) matrix: singular values S (3)
 * descending and right singular vectors V (3,3) as columns (sign of each column is arbitrary, as in LAPACK). */
int sed_svd3(const float* A, int m, float* S, float* V, sed_stream_t stream);

/* CustomSVD.backward (src/fitting_utils.py:385-417, :449-452; only grad_V flows back, as in the reference): U (m,3), S (3),
 * V (3,3) of the forward, grad_V (3,3) -> grad_input (m,3) = 2 U diag(S) sym(K^T o (V^T grad_V)) V^T with svd_grad_K's 1e-6
 * floor on |S_i - S_j|. */
int sed_svd3_backward(const float* U, const float* S, const float* V, const float* grad_V, int m, float* grad_input,
                      sed_stream_t stream);

/* ResidualLoss.residual_loss(points, parameters, sqrt) with reduce=True, src/primitives.py:36-44 over
 * ComputePrimitiveDistance.distance_from_* :89-195: mean (guarded-sqrt) distance of each segment's points to
 * its fitted primitive.  residual (B,S) f32 (0 for skipped segments). */
int sed_residual_segments(const float* points, const int64_t* labels, const int* seg_type, const float* params,
                          const int* status, int B, int N, int S, int use_sqrt, float* residual,
                          sed_stream_t stream);

/* ComputePrimitiveDistance.distance_from_{plane,sphere,cylinder,cone,torus} with reduce=False
 * (src/primitives.py:89-111,113-127,129-161,166-195,58-87): points (n,3), params[8] device, out (n).
 * prim: SED_PRIM_* or 7 for torus ([axis3 center3 major minor]). */
int sed_primitive_distance(const float* points, int n, int prim, const float* params, int use_sqrt, float* out,
                           sed_stream_t stream);

/* Per-segment vote of the predicted primitive type (Fitting_patches_and_edges/residual_utils.py:245-285):
 * log_prob (B,P,N) -> pred_type (B,N) int32 (argmax, generate_predictions_aug.py:365; with log_prob NULL the
 * given pred_type is used as is); with labels (B,N) int64: seg_type (B,S) = mode over the segment (lowest id on
 * ties), seg_count (B,S). */
int sed_segment_types(const float* log_prob, const int64_t* labels, int B, int P, int N, int S, int* pred_type,
                      int* seg_type, int* seg_count, sed_stream_t stream);

/* ------------------------------------------------------------------ end-to-end step with HOST buffers */

/* One inference step over a batch, generate_predictions_aug.py:213-236,365,379-387 followed by the analytic
 * fits of every predicted segment (residual_utils.py:210-331): two forwards (type net, instance net), type
 * argmax, normalise, guarded mean-shift (quantile *= 1.2 while > 49 labels), per-segment type vote, fits,
 * residuals.  All *_host pointers are HOST memory (pinned for best speed); device buffers live in the handle.
 * Synchronises `stream` before returning (host sync). */
typedef struct sed_pipeline sed_pipeline_t;
int sed_pipeline_create(int max_B, int N, int k, int max_segments, sed_pipeline_t** out);
void sed_pipeline_destroy(sed_pipeline_t* p);
/* upload the two weight sets (host arrays of HOST pointers, SED_P_COUNT each; shapes as in enum sed_param). */
int sed_pipeline_set_weights(sed_pipeline_t* p, const float* const* type_params_host,
                             const float* const* inst_params_host);
/* points_host, normals_host (B,N,3); outputs: labels_host (B,N) int64, pred_type_host (B,N) int32,
 * seg_type_host (B,S) int32, params_host (B,S,8), status_host (B,S) int32, residual_host (B,S),
 * bw_host (B), n_labels_host (B) int32.  S = max_segments of the handle. prec_mode as in sed_ms_shift. */
int sed_pipeline_run_host(sed_pipeline_t* p, const float* points_host, const float* normals_host, int B,
                          double quantile, int iterations, int prec_mode, int64_t* labels_host,
                          int* pred_type_host, int* seg_type_host, float* params_host, int* status_host,
                          float* residual_host, float* bw_host, int* n_labels_host, sed_stream_t stream);
/* same step with inputs already resident on the device ((B,N,3) each) and results left on the device in the
 * handle; used for the device-resident throughput number.  sed_pipeline_device_ptr returns the named buffer. */
int sed_pipeline_run_device(sed_pipeline_t* p, const float* points_dev, const float* normals_dev, int B,
                            double quantile, int iterations, int prec_mode, sed_stream_t stream);
/* The two halves of sed_pipeline_run_device, for a driver that post-processes the embedding in between (the
 * reference inserts hpnet_process there, generate_predictions_aug.py:371-380):
 *   run_forward  pack the input, first-layer graph, both networks, type argmax, X = normalised embedding (B,N,128)
 *                (device buffers "X", "embedding", "log_prob", "type_log_prob", "pred_type" of the handle);
 *   run_cluster  guarded mean-shift of the handle's X (which the caller may have rewritten), per-segment type vote
 *                from pred_type, fits, residuals.  Host sync per guard try. */
int sed_pipeline_run_forward(sed_pipeline_t* p, const float* points_dev, const float* normals_dev, int B,
                             sed_stream_t stream);
int sed_pipeline_run_cluster(sed_pipeline_t* p, const float* points_dev, const float* normals_dev, int B, double quantile,
                             int iterations, int prec_mode, sed_stream_t stream);
/* Width of the rows the clustering half reads from the handle's X buffer, (B,N,d) dense: 128 after run_forward (the network's
 * embedding); a driver that replaced X by a wider embedding -- the reference's hpnet_process concatenation is 148 columns,
 * generate_predictions_aug.py:371-380 -- sets it before run_cluster.  d a multiple of 4, <= 192 (the buffers' allocated width).
 * run_forward resets it to 128.  Above 128 columns pass prec_mode 3 to run_cluster (the 192-wide tensor-core kernel); mode 1
 * runs the FP32 FFMA kernel there (see sed_ms_shift). */
int sed_pipeline_set_cluster_width(sed_pipeline_t* p, int d);
void* sed_pipeline_device_ptr(sed_pipeline_t* p, const char* name);
/* device time (ms, CUDA events on the run's stream) of the stages of the last run: first-layer graph (shared by the
 * two networks), both forwards (type net on an internal side stream, concurrently with the instance net) +
 * normalise, bandwidth, shift iterations, nms (+ guard retries), type vote + fits + residuals; retries = number of
 * guard re-runs. Waits for the run to finish. */
int sed_pipeline_stage_ms(sed_pipeline_t* p, float* ms6_host, int* retries_host);
/* number of kernels the library launched since the last call with reset != 0 (for bench accounting). */
int64_t sed_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif /* SEDNET_B200_H */
