// 1x1 convolutions of the SEDNet forward, weight-stationary and persistent (tcgen05 + TMEM), sm_100a.
//
// Same contract as pw_gemm (pointwise.cu):  Y = W . act(in_a * X + in_s) + bias  over channel-major activations, with the
// GroupNorm partial statistics and the max / min pooling of the output produced in the epilogue -- reference
// src/SEDNet.py:78-98 (EdgeConv per-point GEMM, mlp1) and :292-342 (conv1 and the heads).
//
// pw_tc_kernel (pointwise_tc.cu) gives every CTA one 128-point x 256-channel tile: per tile it pays a prologue (TMEM
// allocation, barriers, first loads), streams 64 KB of weights per 64 input channels from L2 and ends in an epilogue that
// nothing overlaps -- its tensor pipe is 3-19 % busy (profiles/pw_tc_r1.md).  This kernel turns the product around:
//     D[co, point] (TMEM, FP32) = W[co, c] . X[c, point]
//   * A = a 128-row tile of the WEIGHTS, loaded ONCE per CTA into tensor memory (FP16 hi | lo of w * 2^k, k per row:
//     <= 256 of the 512 columns) and kept there while the CTA walks over its share of the 128-point tiles;
//   * B = activations.  FP32 channel-major in HBM, they need the previous layer's GroupNorm affine + activation: TMA stages
//     the raw 64-channel x 128-point FP32 boxes in a 3-deep ring (96 KB in flight per SM hide the L2 latency that register
//     prefetching could not), eight producer warps read them, transform, split each value into FP16 hi + lo and write the
//     canonical MN-major 128B-swizzle image (points contiguous) into a 3-stage operand ring;
//   * D += Wh.Xh + Wh.Xl + Wl.Xh (22 significant bits per operand, FP32 accumulation) into one of TWO accumulators, so the
//     MMA warp works on tile t + 1 while four epilogue warps drain tile t;
//   * epilogue: thread = output channel (TMEM lane), columns = points: row scale 2^-k, bias, statistics and max / min are
//     per-thread running values (no shuffles per element), channel-major stores are 128 contiguous bytes per thread.
// Weights never travel again after the first load; what is re-read (from L2) is the activation tile, once per 128 output
// channels.
#include <algorithm>

#include "tc_common.cuh"

namespace sed {

constexpr int P2_THREADS = 448;          // warp 0 TMEM + TMA (raw activations), warp 1 MMA, warps 2-9 producers, warps 10-13 epilogue
constexpr int P2_M = 128;                // output channels per CTA (UMMA M, TMEM lanes)
constexpr int P2_N = 128;                // points per tile (UMMA N)
constexpr int P2_KC = 64;                // input channels per stage
constexpr int P2_STAGES = 3;             // FP16 operand stages
constexpr int P2_RAW = 3;                // raw FP32 stages (TMA)
constexpr int P2_KMAX = 256;             // input channels held in tensor memory (hi + lo: 256 columns)
constexpr uint32_t P2_PART = P2_N * P2_KC * 2;       // 16 KB: [2 halves of 64 points][64 channels][64 points] fp16
constexpr uint32_t P2_STAGE = 2 * P2_PART;            // hi + lo
constexpr uint32_t P2_RAW_BYTES = P2_N * P2_KC * 4;   // 32 KB: [64 channels][128 points] f32
constexpr uint32_t P2_TR_PITCH = 36;                  // floats per scratch row: 144 B keeps 16-byte accesses aligned and conflict-free
constexpr uint32_t P2_TR_BYTES = 32 * P2_TR_PITCH * 4; // one epilogue warp's transpose scratch
constexpr uint32_t P2_COL_D = 256;                    // accumulators at columns [256,384) and [384,512)

struct P2Params {
    const float* X; long long x_bstride; int ldx;
    const __half* Wh; const __half* Wl;        // (Cout, kpad) pre-split weights
    const float* rscale;                       // (Cout) 2^-k of the weight rows
    const float* bias; long long bias_bstride;
    const float* in_a; const float* in_s; int in_act;
    float* Y; long long y_bstride; int ldy; int y_point_major;
    double* stats; float* mm;
    int Cin, Cout, N, kpad, B, tiles_per_cloud, ctas_per_group;
};

__device__ __forceinline__ float p2_act(float v, int act) {
    if (act == 1) return fmaxf(v, 0.f);
    if (act == 2) return v >= 0.f ? v : 0.2f * v;
    return v;
}

// A thread owns 8 consecutive points of a channel row = two 16-byte pieces of the raw box.  With a 32-byte lane stride, reading
// "first piece, then second piece" puts lanes c and c + 4 of a quarter-warp on the same banks (2-way conflict on every raw
// LDS.128).  Lanes with sel = 1 therefore read their pieces in the opposite order: v[0..3] holds piece `sel`, v[4..7] piece
// `1 - sel` (conflict-free in both loads), and the FP16 image gets two 8-byte stores whose order follows `sel` (sel also
// flips between the two 64-point halves, so the half-warps of the STS.64 stay conflict-free).
__device__ __forceinline__ int p2_sel(int chunk) { return ((chunk >> 2) ^ (chunk >> 3)) & 1; }

// FP16 hi / lo split of the 8 points and their stores into the operand image (16-byte chunk at hi_addr / lo_addr)
__device__ __forceinline__ void p2_split_store(const float (&v)[8], uint32_t hi_addr, uint32_t lo_addr, int sel) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const __half2 h = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
        const float2 hf = __half22float2(h);
        const __half2 l = __floats2half2_rn(v[2 * e] - hf.x, v[2 * e + 1] - hf.y);
        hi[e] = *reinterpret_cast<const uint32_t*>(&h);
        lo[e] = *reinterpret_cast<const uint32_t*>(&l);
    }
    const uint32_t o0 = (uint32_t)sel * 8u, o1 = 8u - o0;      // where v[0..3] and v[4..7] belong inside the 16-byte chunk
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(hi_addr + o0), "r"(hi[0]), "r"(hi[1]));
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(hi_addr + o1), "r"(hi[2]), "r"(hi[3]));
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(lo_addr + o0), "r"(lo[0]), "r"(lo[1]));
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(lo_addr + o1), "r"(lo[2]), "r"(lo[3]));
}

// Interior boxes (all 128 points and all 64 channels exist -- every box but those of a cloud's last tile): no bounds tests,
// the activation known at compile time, the eight rows' scale / shift loaded up front.  ACT: -1 raw (no affine), 0 affine only,
// 1 affine + ReLU, 2 affine + LeakyReLU(0.2).
template <int ACT>
__device__ __forceinline__ void p2_convert_interior(float (&cur)[8][8], const float* __restrict__ ia, const float* __restrict__ is,
                                                    int c_first, int cl_first, int chunk, uint32_t x_hi, uint32_t x_lo, int sel) {
    float av[8], sv[8];
    if (ACT >= 0) {
#pragma unroll
        for (int q = 0; q < 8; ++q) { av[q] = __ldg(ia + c_first + q * 2); sv[q] = __ldg(is + c_first + q * 2); }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int cl = cl_first + q * 2;
        float (&v)[8] = cur[q];
        if (ACT >= 0) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                float t = fmaf(av[q], v[e], sv[q]);
                if (ACT == 1) t = fmaxf(t, 0.f);
                if (ACT == 2) t = t >= 0.f ? t : 0.2f * t;
                v[e] = t;
            }
        }
        const uint32_t off = (uint32_t)(chunk >> 3) * (P2_PART / 2) + (uint32_t)cl * 128u + (uint32_t)(((chunk & 7) ^ (cl & 7)) << 4);
        p2_split_store(v, x_hi + off, x_lo + off, sel);
    }
}

__global__ void __launch_bounds__(P2_THREADS, 1) pw_tc2_kernel(const __grid_constant__ CUtensorMap map_x, P2Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw0 = smem_u32(smem_raw);
    const uint32_t stage_addr = (raw0 + 1023u) & ~1023u;                    // P2_STAGES x [Xh | Xl]
    const uint32_t raw_addr = stage_addr + P2_STAGES * P2_STAGE;            // P2_RAW x [64 channels][128 points] f32
    const uint32_t tr_addr = raw_addr + P2_RAW * P2_RAW_BYTES;              // 4 epilogue warps x [32 channels][36] f32 (store transpose)
    const uint32_t bar_base = tr_addr + 4 * P2_TR_BYTES;
    const uint32_t bar_x_full = bar_base;                      // [STAGES] the 4 producer warps of the box's group arrive
    const uint32_t bar_x_empty = bar_x_full + 8 * P2_STAGES;   // [STAGES] MMA commit
    const uint32_t bar_d_full = bar_x_empty + 8 * P2_STAGES;   // [2] MMA commit
    const uint32_t bar_d_empty = bar_d_full + 16;              // [2] 128 epilogue threads
    const uint32_t bar_w_full = bar_d_empty + 16;              // 128 epilogue threads have written the weights
    const uint32_t bar_r_full = bar_w_full + 8;                // [RAW] TMA bytes
    const uint32_t bar_r_empty = bar_r_full + 8 * P2_RAW;      // [RAW] the 4 producer warps of the box's group have read it
    const uint32_t tmem_slot = bar_r_empty + 8 * P2_RAW;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw0));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int co0 = blockIdx.y * P2_M;
    const int nk = p.kpad / P2_KC;
    const int total_tiles = p.B * p.tiles_per_cloud;
    const int G = p.ctas_per_group;
    const int my_tiles = ((int)blockIdx.x < total_tiles) ? (total_tiles - 1 - (int)blockIdx.x) / G + 1 : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < P2_STAGES; ++s) { mbar_init(bar_x_full + 8 * s, 4); mbar_init(bar_x_empty + 8 * s, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(bar_d_full + 8 * i, 1); mbar_init(bar_d_empty + 8 * i, 128); }
        mbar_init(bar_w_full, 128);
        for (int s = 0; s < P2_RAW; ++s) { mbar_init(bar_r_full + 8 * s, 1); mbar_init(bar_r_empty + 8 * s, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tmem_slot) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot_ptr;

    if (warp == 0) {
        // ============================================================ TMA: raw FP32 activation boxes of every (tile, chunk)
        if (my_tiles > 0 && elect_one()) {
            int it = 0;
            for (int t = 0; t < my_tiles; ++t) {
                const int tile = (int)blockIdx.x + t * G;
                const int b = tile / p.tiles_per_cloud, n0 = (tile % p.tiles_per_cloud) * P2_N;
                for (int kc = 0; kc < nk; ++kc, ++it) {
                    const int rs = it % P2_RAW;
                    if (it >= P2_RAW) mbar_wait(bar_r_empty + 8 * rs, ((it / P2_RAW) - 1) & 1);
                    mbar_expect_tx(bar_r_full + 8 * rs, P2_RAW_BYTES);
                    tma_load_3d(raw_addr + rs * P2_RAW_BYTES, &map_x, bar_r_full + 8 * rs, n0, kc * P2_KC, b);   // OOB reads as 0
                }
            }
        }
    } else if (warp == 1) {
        // ============================================================ MMA issuer
        if (my_tiles > 0 && elect_one()) {
            constexpr uint32_t IDESC = make_idesc_n(1, P2_N);     // B (activations) MN-major, N = 128
            const uint32_t w_hi = tmem, w_lo = tmem + (uint32_t)(p.kpad >> 1);
            mbar_wait(bar_w_full, 0);
            tc_fence_after();
            int it = 0;                                            // running chunk counter over all my tiles
            for (int t = 0; t < my_tiles; ++t) {
                const int buf = t & 1;
                if (t >= 2) { mbar_wait(bar_d_empty + 8 * buf, ((t >> 1) - 1) & 1); tc_fence_after(); }
                const uint32_t d = tmem + P2_COL_D + (uint32_t)buf * P2_N;
                for (int kc = 0; kc < nk; ++kc, ++it) {
                    const int s = it % P2_STAGES;
                    mbar_wait(bar_x_full + 8 * s, (it / P2_STAGES) & 1);
                    tc_fence_after();
                    const uint32_t x_hi = stage_addr + s * P2_STAGE, x_lo = x_hi + P2_PART;
#pragma unroll
                    for (int term = 0; term < 3; ++term) {
                        const uint32_t wa = ((term == 2) ? w_lo : w_hi) + (uint32_t)kc * (P2_KC / 2);   // Wh, Wh, Wl
                        const uint32_t xb = (term == 1) ? x_lo : x_hi;                                  // Xh, Xl, Xh
#pragma unroll
                        for (int ks = 0; ks < P2_KC / 16; ++ks)
                            // B: MN-major, 16 channels = 16 rows of 128 B; the two halves of 64 points are P2_PART / 2 apart
                            umma_ts(d, wa + ks * 8, make_desc(xb + ks * 2048, P2_PART / 2), IDESC,
                                    (kc > 0 || term > 0 || ks > 0) ? 1u : 0u);
                    }
                    tc_commit(bar_x_empty + 8 * s);
                }
                tc_commit(bar_d_full + 8 * buf);
            }
        }
    } else if (warp >= 2 && warp < 10) {
        // ============================================================ producers: activations -> FP16 hi/lo MN-major image
        // Two groups of four warps take alternate boxes (group g: boxes it = g, g + 2, ...), so that two boxes are always
        // in conversion: one group's waits (raw box full, scale / shift loads, operand stage free) overlap the other's math.
        // Within a box a warp owns 16 channel rows, a thread 8 of them x 8 consecutive points.
        const int pw = warp - 2, group = pw >> 2, sub = pw & 3;
        const int chunk = lane & 15;                             // 8 consecutive points
        const int sel = p2_sel(chunk);                           // which of its two raw pieces this lane reads first
        const uint32_t so = (uint32_t)sel * 16u;
        const int total = my_tiles * nk;
        for (int it = group; it < total; it += 2) {
            const int t = it / nk, kc = it - t * nk;
            const int tile = (int)blockIdx.x + t * G;
            const int b = tile / p.tiles_per_cloud, n0 = (tile % p.tiles_per_cloud) * P2_N;
            const int pt = n0 + chunk * 8;
            const float* ia = p.in_a ? p.in_a + (long long)b * p.Cin : nullptr;
            const float* is = p.in_s ? p.in_s + (long long)b * p.Cin : nullptr;
            const int s = it % P2_STAGES, rs = it % P2_RAW;
            const int cl0 = sub * 16 + (lane >> 4);              // this thread's rows of the box: cl0 + 2 q, q = 0..7
            // Operand stage first, raw box second.  A parity wait is only meaningful on a barrier that is at most one phase
            // behind, and the previous box of this raw slot (it - 3) belonged to the OTHER group: nothing this warp has seen
            // so far proves that box landed (TMA boxes may complete out of order), and a wait for phase k on a barrier still
            // in phase k - 1 returns at once.  The stage wait closes that: stage s free => the MMA consumed box it - 3 => both
            // groups read every box <= it - 3 (their r_empty arrivals precede their x_full arrivals) => r_full[rs] is in
            // phase k.  (Boxes 0..2 wait for phase 0 of fresh barriers.)
            static_assert(P2_RAW >= P2_STAGES, "the stage wait must cover the raw slot's previous box");
            if (it >= P2_STAGES) mbar_wait(bar_x_empty + 8 * s, ((it / P2_STAGES) - 1) & 1);
            // raw box: [channel][128 points] f32, 512-B rows
            float cur[8][8];
            mbar_wait(bar_r_full + 8 * rs, (it / P2_RAW) & 1);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const uint32_t ra = raw_addr + rs * P2_RAW_BYTES + (uint32_t)(cl0 + q * 2) * 512u + (uint32_t)chunk * 32u;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(cur[q][0]), "=f"(cur[q][1]), "=f"(cur[q][2]), "=f"(cur[q][3]) : "r"(ra + so));
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(cur[q][4]), "=f"(cur[q][5]), "=f"(cur[q][6]), "=f"(cur[q][7]) : "r"(ra + 16u - so));
            }
            const uint32_t x_hi = stage_addr + s * P2_STAGE, x_lo = x_hi + P2_PART;
            if (n0 + P2_N <= p.N && kc * P2_KC + P2_KC <= p.Cin) {                   // interior box (CTA-uniform)
                const int c0 = kc * P2_KC + cl0;
                if (!ia) p2_convert_interior<-1>(cur, ia, is, c0, cl0, chunk, x_hi, x_lo, sel);
                else if (p.in_act == 1) p2_convert_interior<1>(cur, ia, is, c0, cl0, chunk, x_hi, x_lo, sel);
                else if (p.in_act == 2) p2_convert_interior<2>(cur, ia, is, c0, cl0, chunk, x_hi, x_lo, sel);
                else p2_convert_interior<0>(cur, ia, is, c0, cl0, chunk, x_hi, x_lo, sel);
            } else {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int cl = cl0 + q * 2;                  // channel within the chunk = row of the image
                    const int c = kc * P2_KC + cl;
                    float (&v)[8] = cur[q];
                    if (c >= p.Cin) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) v[e] = 0.f;
                    } else if (ia) {
                        const float av = __ldg(ia + c), sv = __ldg(is + c);
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int pe = pt + ((e < 4) ? e + 4 * sel : e - 4 * sel);       // the point v[e] holds
                            v[e] = (pe < p.N) ? p2_act(fmaf(av, v[e], sv), p.in_act) : 0.f;
                        }
                    }
                    // image: [half = point / 64][row = channel][128 B = 64 points], 16-byte chunks XOR-swizzled by (row & 7)
                    const uint32_t off = (uint32_t)(chunk >> 3) * (P2_PART / 2) + (uint32_t)cl * 128u +
                                         (uint32_t)(((chunk & 7) ^ (cl & 7)) << 4);
                    p2_split_store(v, x_hi + off, x_lo + off, sel);
                }
            }
            // generic-proxy writes of the image -> tensor-core reads; generic-proxy reads of the raw box -> TMA's next write
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) { mbar_arrive(bar_r_empty + 8 * rs); mbar_arrive(bar_x_full + 8 * s); }   // in this order (see above)
        }
    } else if (warp >= 10) {
        // ============================================================ epilogue warps: thread = output channel
        const int quarter = warp & 3;                           // TMEM lane quarter this warp may access
        const int row = quarter * 32 + lane;
        const int co = co0 + row;
        const bool cvalid = co < p.Cout;
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        {   // ---- this thread's weight row -> TMEM: hi at columns [0, kpad/2), lo at [kpad/2, kpad)
#pragma unroll 1
            for (int part = 0; part < 2; ++part) {
                const uint4* g4 = reinterpret_cast<const uint4*>((part ? p.Wl : p.Wh) + (long long)min(co, p.Cout - 1) * p.kpad);
                const uint32_t wb = tmem + lane_addr + (uint32_t)part * (uint32_t)(p.kpad >> 1);
#pragma unroll 1
                for (int c = 0; c < p.kpad / 32; ++c) {          // 32 halves = 16 columns per store
                    uint32_t w[16];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint4 v = make_uint4(0u, 0u, 0u, 0u);
                        if (cvalid) v = __ldg(g4 + c * 4 + i);
                        w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
                    }
                    tmem_st16(wb + c * 16, w);
                }
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(bar_w_full);
        }
        const float rs = cvalid ? __ldg(p.rscale + co) : 0.f;
        const int nblk = (p.Cout + 31) / 32;
        for (int t = 0; t < my_tiles; ++t) {
            const int tile = (int)blockIdx.x + t * G;
            const int b = tile / p.tiles_per_cloud, tp = tile % p.tiles_per_cloud, n0 = tp * P2_N;
            const int buf = t & 1;
            const float bv = (cvalid && p.bias) ? __ldg(p.bias + (long long)b * p.bias_bstride + co) : 0.f;
            mbar_wait(bar_d_full + 8 * buf, (t >> 1) & 1);
            tc_fence_after();
            const uint32_t db = tmem + lane_addr + P2_COL_D + (uint32_t)buf * P2_N;
            float s1 = 0.f, s2 = 0.f, mx = -INFINITY, mn = INFINITY;
            double ds1 = 0.0, ds2 = 0.0;
#pragma unroll 1
            for (int cc = 0; cc < 4; ++cc) {
                uint32_t v[32];
                tmem_ld32(db + cc * 32, v);
                tmem_ld_wait();
                if (cc == 3) {                                  // every tcgen05.ld of this accumulator has completed
                    tc_fence_before();
                    mbar_arrive(bar_d_empty + 8 * buf);
                }
                const int nb = n0 + cc * 32;
                float y[32];
                s1 = 0.f; s2 = 0.f;
#pragma unroll
                for (int i = 0; i < 32; ++i) y[i] = __fadd_rn(__uint_as_float(v[i]) * rs, bv);
                if (nb + 32 <= p.N) {                           // interior chunk: no per-point tests
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        s1 += y[i]; s2 = fmaf(y[i], y[i], s2);
                        mx = fmaxf(mx, y[i]); mn = fminf(mn, y[i]);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        if (nb + i < p.N) {
                            s1 += y[i]; s2 = fmaf(y[i], y[i], s2);
                            mx = fmaxf(mx, y[i]); mn = fminf(mn, y[i]);
                        }
                    }
                }
                ds1 += (double)s1; ds2 += (double)s2;
                if (p.Y && cvalid && p.y_point_major) {
                    float* Yb = p.Y + (long long)b * p.y_bstride;
                    {
                        float* o = Yb + (long long)nb * p.ldy + co;                                 // lanes = consecutive channels
                        if (nb + 32 <= p.N) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) o[(long long)i * p.ldy] = y[i];
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; ++i)
                                if (nb + i < p.N) o[(long long)i * p.ldy] = y[i];
                        }
                    }
                }
                if (p.Y && !p.y_point_major) {
                    // channel-major output: a thread owns one channel row, so a direct store scatters 32 x 16 B per warp
                    // instruction over 32 rows (32 LSU wavefronts each: the 4096 store cycles per tile were what bounded the
                    // layers that write Y).  Interior chunks go through a 32 x 36 shared-memory transpose instead: every store
                    // instruction then writes 4 rows x 128 contiguous bytes.
                    float* Yb = p.Y + (long long)b * p.y_bstride;
                    const int cb = co0 + quarter * 32;                                   // the warp's first channel
                    float* ob = Yb + (long long)cb * p.ldy + nb;
                    const bool fast = (nb + 32 <= p.N) && (cb + 32 <= p.Cout) && ((p.ldy & 3) == 0) &&
                                      ((reinterpret_cast<uintptr_t>(ob) & 15) == 0);
                    if (fast) {                                                          // warp-uniform
                        const uint32_t sc = tr_addr + (uint32_t)(warp - 10) * P2_TR_BYTES;
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sc + (uint32_t)(lane * P2_TR_PITCH + 4 * i) * 4u),
                                         "f"(y[4 * i]), "f"(y[4 * i + 1]), "f"(y[4 * i + 2]), "f"(y[4 * i + 3]));
                        __syncwarp();
#pragma unroll
                        for (int r4 = 0; r4 < 8; ++r4) {
                            const int rrow = r4 * 4 + (lane >> 3), col = (lane & 7) * 4;
                            float4 o4;
                            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(o4.x), "=f"(o4.y), "=f"(o4.z), "=f"(o4.w)
                                         : "r"(sc + (uint32_t)(rrow * P2_TR_PITCH + col) * 4u));
                            *reinterpret_cast<float4*>(ob + (long long)rrow * p.ldy + col) = o4;
                        }
                        __syncwarp();
                    } else if (cvalid) {
                        float* o = Yb + (long long)co * p.ldy + nb;
#pragma unroll
                        for (int i = 0; i < 32; ++i) if (nb + i < p.N) o[i] = y[i];
                    }
                }
            }
            if (p.stats) {                                      // one 32-channel block per warp: fixed-order butterfly
                const double t1 = warp_sum_d(cvalid ? ds1 : 0.0), t2 = warp_sum_d(cvalid ? ds2 : 0.0);
                const int blk = (co0 >> 5) + quarter;
                if (lane == 0 && blk < nblk) {
                    double* o = p.stats + (((long long)b * p.tiles_per_cloud + tp) * nblk + blk) * 2;
                    o[0] = t1; o[1] = t2;
                }
            }
            if (p.mm && cvalid) {
                float* o = p.mm + (((long long)b * p.tiles_per_cloud + tp) * p.Cout + co) * 2;
                o[0] = mx; o[1] = mn;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

// pointwise_tc.cu: hi / lo FP16 split of w * 2^k (k per row) into rows of cin_pad halves, rscale[row] = 2^-k
int pw_prep_weights(const float* W, int ldw, int Cout, int Cin, int cin_pad, __half* Wh, __half* Wl, float* rscale,
                    cudaStream_t st);

// Weight-stationary implementation of pw_gemm (same arguments).  SED_ERR_UNSUPPORTED for shapes it does not cover:
// Cin < 32 (FFMA kernel), Cin > 256 (the weight tile must fit 256 TMEM columns: pw_tc_kernel), Cout < 64.
int pw_gemm_tc2(const float* X, long long x_bstride, int ldx, const float* Wt, int ldw, const float* bias,
                long long bias_bstride, const float* in_a, const float* in_s, int in_act, float* Y, long long y_bstride,
                int ldy, int y_point_major, double* stats, float* mm, int B, int Cin, int Cout, int N, cudaStream_t st) {
    if (Cin < 32 || Cin > P2_KMAX || Cout < 64) return SED_ERR_UNSUPPORTED;
    // TMA needs 16-byte aligned rows
    if ((ldx & 3) || (x_bstride & 3) || (reinterpret_cast<uintptr_t>(X) & 15)) return SED_ERR_UNSUPPORTED;
    const int kpad = (Cin + P2_KC - 1) / P2_KC * P2_KC;
    const size_t wbytes = (size_t)Cout * kpad * sizeof(__half);
    char* buf = nullptr;
    bool temporary = true;
    SED_TRY(pw_prepared(Wt, ldw, Cout, Cin, kpad, st, &buf, &temporary));    // split once per matrix when a cache is bound
    __half* Wh = (__half*)buf;
    __half* Wl = (__half*)(buf + align_up(wbytes));
    float* rscale = (float*)(buf + 2 * align_up(wbytes));
    int rc = SED_OK;
    const int tiles_per_cloud = (N + P2_N - 1) / P2_N;
    const int groups = (Cout + P2_M - 1) / P2_M;
    const int G = std::max(1, std::min(kNumSMs / groups, B * tiles_per_cloud));
    // raw activations by TMA: (N, Cin, B) f32 with strides (ldx, x_bstride) elements; box 128 points x 64 channels
    CUtensorMap mx;
    if (rc == SED_OK) {
        EncodeTiledFn fn = get_encode_fn();
        const cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)Cin, (cuuint64_t)B};
        const cuuint64_t strides[2] = {(cuuint64_t)ldx * 4, (cuuint64_t)x_bstride * 4};
        const cuuint32_t box[3] = {P2_N, P2_KC, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        if (!fn || fn(&mx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)X, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            rc = SED_ERR_UNSUPPORTED;
    }
    P2Params p{X, x_bstride, ldx, Wh, Wl, rscale, bias, bias_bstride, in_a, in_s, in_act, Y, y_bstride, ldy, y_point_major,
               stats, mm, Cin, Cout, N, kpad, B, tiles_per_cloud, G};
    constexpr size_t smem = (size_t)P2_STAGES * P2_STAGE + (size_t)P2_RAW * P2_RAW_BYTES + 4 * P2_TR_BYTES + 1024 + 256;
    static_assert(smem <= 227 * 1024, "shared memory budget");
    cudaError_t e = cudaFuncSetAttribute(pw_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) rc = SED_ERR_CUDA_BASE - (int)e;
    if (rc == SED_OK) {
        pw_tc2_kernel<<<dim3(G, groups), P2_THREADS, smem, st>>>(mx, p);
        ++g_sed_launches;
        e = cudaGetLastError();
        if (e != cudaSuccess) rc = SED_ERR_CUDA_BASE - (int)e;
    }
    if (temporary) cudaFreeAsync(buf, st);
    return rc;
}

}  // namespace sed
