// Fused pairwise-score + exact top-k selection on the tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
// Serves   knn                 src/PointNet.py:62-87    (negative squared L2 in Gram form, k nearest, sorted)
//          knn_points_normals  src/PointNet.py:90-137   (point x normal metric of the first EdgeConv layer)
//          compute_bandwidth   src/mean_shift.py:130-135 (K-th smallest cosine distance of every row)
//          nms membership/labels src/mean_shift.py:146-149,177-178 (nearest centre of every point)
//
// The N x N score matrix is never written anywhere: a CTA owns 128 query rows, the 128-candidate tiles stream through
// shared memory by TMA, tcgen05.mma writes the 128 x 128 Gram tile into TMEM (double-buffered), and 128 threads --
// ONE THREAD PER QUERY ROW, the natural TMEM access pattern -- read their row with tcgen05.ld, turn each dot product
// into the reference's FP32 score and feed an exact radix select: 4 passes of 8 bits over the order-preserving integer
// image of the score, each pass a thread-private histogram in shared memory (no atomics, no inter-thread traffic),
// each pass simply re-running the (cheap) tensor-core Gram.  A fifth pass collects the k winners, which one warp per
// row then sorts with a register bitonic network.  Ties at the k-th score go to the lowest candidate index.
//
// Operands are FP16 hi/lo splits (22 significant bits, FP32 accumulation, Q_h X_h + Q_h X_l + Q_l X_h) of the inputs
// scaled per cloud by a power of two that puts max|x| in [1024, 2048), so the Gram is FP32-faithful for any input
// range; squared norms are FP32 from the unsplit values.
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "tc_common.cuh"

namespace sed {

enum { SEL_L2 = 0, SEL_PN = 1, SEL_COS = 2 };
enum { OUT_IDX = 0, OUT_KTH = 1, OUT_TOP1 = 2 };

constexpr int ST_THREADS = 320;   // warp 0 TMA, warp 1 MMA, warps 2-9 select (thread per row, two groups)
constexpr int ST_M = 128, ST_NC = 128;
constexpr int ST_KMAX = 64;       // list capacity of the OUT_IDX mode

struct SelParams {
    const float* xxq; const float* xxc;   // squared norms (B, Npad) for L2 / PN (xxq == xxc for self-kNN)
    const float* maxabs_q; const float* maxabs_c;  // per-cloud max|x| (operand scale); null -> fixed scale
    float fixed_scale;
    int Nq, Nc, npad, k;
    const int* nc_ptr;                    // optional per-cloud candidate count
    float W;
    void* out_idx; int idx64;             // OUT_IDX: (B,Nq,k); OUT_TOP1: (B,Nq)
    float* out_kth;                       // OUT_KTH: (B,Nq)
    int unsorted;                         // streaming kNN: the k neighbours may be written in any order (EdgeConv's max / sum)
    int win, soft;                        // streaming kernels: prune window / soft mark (tuning; defaults SS_WIN / KS_WIN, SS_SOFT)
};

static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

__host__ __device__ __forceinline__ float scale_from_maxabs(float m) {
    // power of two s with s * m in [1024, 2048); 1 for empty / degenerate input
    if (!(m > 0.f) || !(m < 3.0e38f)) return 1.0f;
    int e;
    frexpf(m, &e);            // m = f * 2^e, f in [0.5, 1)
    e = 11 - e;
    e = e < -60 ? -60 : (e > 60 ? 60 : e);
    return ldexpf(1.0f, e);
}

template <int MODE>
__device__ __forceinline__ float sel_score(float acc0, float acc1, float inv_s2, float xq, float xc, float W) {
    if (MODE == SEL_L2) {
        // src/PointNet.py:76-78: inner = -2 x.x' ; pd = (-xx_j - inner) - xx_i
        const float dot = acc0 * inv_s2;
        return __fsub_rn(fmaf(2.0f, dot, -xc), xq);
    } else if (MODE == SEL_PN) {
        // src/PointNet.py:112-120,128: p = (xx_j - 2 p.p') + xx_i ; n = 2 - 2 n.n' ; -(p * (1 + n * W))
        const float pd = __fadd_rn(fmaf(-2.0f, acc0 * inv_s2, xc), xq);
        const float nd = fmaf(-2.0f, acc1 * inv_s2, 2.0f);
        return -__fmul_rn(pd, __fadd_rn(1.0f, __fmul_rn(nd, W)));
    } else {
        // src/mean_shift.py:130,146: dist = 2 - 2 x.y ; score = -dist
        return -fmaf(-2.0f, acc0 * inv_s2, 2.0f);
    }
}

// KB: 64-channel boxes per operand row (1: K <= 64, 2: K = 128).  RBITS: radix bits per pass.  CB: bytes per histogram
// counter -- 1 (saturating at 255: ranks k <= 255) or 2 (ranks up to 65 535: the K-th-neighbour select of a guard retry
// with a large quantile, src/mean_shift.py:81-96; one candidate stage less to make room for the wider histograms).
template <int KB, int CB>
struct SelStages { static constexpr int value = (CB == 2) ? (KB == 1 ? 3 : 1) : ((KB == 1) ? 3 : (KB == 2 ? 2 : 1)); };

template <int MODE, int OUT, int KB, int RBITS, int CB = 1>
__global__ void __launch_bounds__(ST_THREADS, 1)
select_tc_kernel(const __grid_constant__ CUtensorMap map_qh, const __grid_constant__ CUtensorMap map_ql,
                 const __grid_constant__ CUtensorMap map_xh, const __grid_constant__ CUtensorMap map_xl, SelParams p) {
    constexpr uint32_t PART_BYTES = KB * BOX_BYTES;
    constexpr uint32_t TILE2 = 2 * PART_BYTES;                     // hi + lo
    constexpr int STAGES = SelStages<KB, CB>::value;             // KB = 3 (192-wide rows): one 96 KB candidate stage
    typedef typename std::conditional<CB == 1, uint8_t, uint16_t>::type CT;
    constexpr uint32_t CMAX = (CB == 1) ? 255u : 65535u;
    constexpr int NBINS = 1 << RBITS;
    constexpr int NPASS_RADIX = (OUT == OUT_TOP1) ? 0 : (32 + RBITS - 1) / RBITS;
    constexpr int NPASS = (OUT == OUT_TOP1) ? 1 : NPASS_RADIX + (OUT == OUT_IDX ? 1 : 0);
    constexpr uint32_t HIST_BYTES = NBINS * 128 * CB;               // one group's histogram
    constexpr uint32_t SEL_BYTES = (OUT == OUT_TOP1) ? 1024 : 2 * HIST_BYTES;   // lists (48 KB) alias the histograms
    static_assert(OUT != OUT_IDX || 2 * HIST_BYTES >= ST_KMAX * 128 * 6, "lists must fit in the histogram area");
    constexpr int NACC = 2;   // PN: points / normals Grams; L2, COS: hi.hi and the cross terms (summed in FP32 RN by the
                              // selection threads: the tensor core truncates when it accumulates, so the small terms
                              // must not ride on the large accumulator)
    constexpr uint32_t BUF_COLS = 128 * NACC;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw0 = smem_u32(smem_raw);
    const uint32_t smem_base = (raw0 + 1023u) & ~1023u;
    const uint32_t q_addr = smem_base;
    const uint32_t x_addr = q_addr + TILE2;
    const uint32_t sel_addr = x_addr + STAGES * TILE2;
    const uint32_t bar_base = sel_addr + SEL_BYTES;
    const uint32_t bar_q_full = bar_base;
    const uint32_t bar_x_full = bar_base + 8;
    const uint32_t bar_x_empty = bar_x_full + 8 * STAGES;
    const uint32_t bar_s_full = bar_x_empty + 8 * STAGES;   // [2]
    const uint32_t bar_s_empty = bar_s_full + 16;            // [2]
    const uint32_t tmem_slot = bar_s_empty + 16;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw0));
    uint8_t* sel_ptr = smem_raw + (sel_addr - raw0);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y, q0 = blockIdx.x * ST_M;
    const int Nc = p.nc_ptr ? min(max(p.nc_ptr[b], 0), p.Nc) : p.Nc;
    const int T = (Nc + ST_NC - 1) / ST_NC;
    const int J = T * NPASS;

    if (threadIdx.x == 0) {
        mbar_init(bar_q_full, 1);
        for (int s = 0; s < STAGES; ++s) { mbar_init(bar_x_full + 8 * s, 1); mbar_init(bar_x_empty + 8 * s, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(bar_s_full + 8 * i, 1); mbar_init(bar_s_empty + 8 * i, 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tmem_slot) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot_ptr;

    if (warp == 0) {
        // ============================================================ TMA producer
        if (J > 0 && elect_one()) {
            mbar_expect_tx(bar_q_full, TILE2);
#pragma unroll
            for (int kb = 0; kb < KB; ++kb) {
                tma_load_3d(q_addr + kb * BOX_BYTES, &map_qh, bar_q_full, kb * 64, q0, b);
                tma_load_3d(q_addr + PART_BYTES + kb * BOX_BYTES, &map_ql, bar_q_full, kb * 64, q0, b);
            }
            for (int j = 0; j < J; ++j) {
                const int s = j % STAGES, tile = j % T;
                if (j >= STAGES) mbar_wait(bar_x_empty + 8 * s, ((j / STAGES) - 1) & 1);
                const uint32_t dst = x_addr + s * TILE2, bar = bar_x_full + 8 * s;
                mbar_expect_tx(bar, TILE2);
#pragma unroll
                for (int kb = 0; kb < KB; ++kb) {
                    tma_load_3d(dst + kb * BOX_BYTES, &map_xh, bar, kb * 64, tile * ST_NC, b);
                    tma_load_3d(dst + PART_BYTES + kb * BOX_BYTES, &map_xl, bar, kb * 64, tile * ST_NC, b);
                }
            }
        }
    } else if (warp == 1) {
        // ============================================================ MMA issuer
        if (J > 0 && elect_one()) {
            constexpr uint32_t IDESC = make_idesc(0);
            mbar_wait(bar_q_full, 0);
            for (int j = 0; j < J; ++j) {
                const int s = j % STAGES;
                mbar_wait(bar_x_full + 8 * s, (j / STAGES) & 1);
                if (j >= 2) mbar_wait(bar_s_empty + 8 * (j & 1), ((j >> 1) - 1) & 1);
                tc_fence_after();
                const uint32_t xs = x_addr + s * TILE2;
                const uint32_t d = tmem + (uint32_t)(j & 1) * BUF_COLS;
                if (MODE == SEL_PN) {
#pragma unroll
                    for (int a = 0; a < 2; ++a) {
                        // channels [0,16): points (3 used), [16,32): normals (3 used): one K-step each
#pragma unroll
                        for (int term = 0; term < 3; ++term) {
                            const uint32_t qa = q_addr + ((term == 2) ? PART_BYTES : 0);   // Qh, Qh, Ql
                            const uint32_t xb = xs + ((term == 1) ? PART_BYTES : 0);       // Xh, Xl, Xh
                            umma_ss(d + a * 128, make_desc(qa + a * 32, 16), make_desc(xb + a * 32, 16), IDESC, term > 0);
                        }
                    }
                } else {
#pragma unroll
                    for (int term = 0; term < 3; ++term) {
                        const uint32_t qa = q_addr + ((term == 2) ? PART_BYTES : 0);
                        const uint32_t xb = xs + ((term == 1) ? PART_BYTES : 0);
                        const uint32_t dd = d + (term == 0 ? 0u : 128u);                   // hi.hi | cross terms
#pragma unroll
                        for (int ks = 0; ks < KB * 4; ++ks) {
                            const uint32_t off = (ks >> 2) * BOX_BYTES + (ks & 3) * 32;
                            umma_ss(dd, make_desc(qa + off, 16), make_desc(xb + off, 16), IDESC,
                                    (ks > 0 || term == 2) ? 1u : 0u);
                        }
                    }
                }
                tc_commit(bar_s_full + 8 * (j & 1));
                tc_commit(bar_x_empty + 8 * s);
            }
        }
    } else {
        // ============================================================ selection: one thread per query row.
        // Two groups of four warps; group g drains TMEM buffer g (tiles with j & 1 == g), each with its own histogram.
        const int group = (warp - 2) >> 2;
        const int quarter = warp & 3;                                   // TMEM lane quarter this warp may access
        const int row = quarter * 32 + lane;
        const int q = q0 + row;
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        const float sq = p.maxabs_q ? scale_from_maxabs(p.maxabs_q[b]) : p.fixed_scale;
        const float sc = p.maxabs_c ? scale_from_maxabs(p.maxabs_c[b]) : p.fixed_scale;
        const float inv_s2 = (1.0f / sq) * (1.0f / sc);
        const float xq = (MODE == SEL_COS) ? 0.f : p.xxq[(long long)b * p.npad + min(q, p.npad - 1)];
        const float* xxc = (MODE == SEL_COS) ? nullptr : p.xxc + (long long)b * p.npad;
        // u8 saturating counters hist[g][bin][slot], slot = lane*4 + quarter: every lane of a warp on its own bank.
        // Saturation at CMAX is exact: bins ABOVE the one holding rank k_rem (<= CMAX) hold fewer than k_rem.
        CT* hist = reinterpret_cast<CT*>(sel_ptr + group * HIST_BYTES) + lane * 4 + quarter;
        const CT* hist_o = reinterpret_cast<const CT*>(sel_ptr + (group ^ 1) * HIST_BYTES) + lane * 4 + quarter;
        uint32_t* lkey = reinterpret_cast<uint32_t*>(sel_ptr) + row;                                   // lkey[pos * 128]
        uint16_t* lidx = reinterpret_cast<uint16_t*>(sel_ptr + ST_KMAX * 128 * 4) + row;               // lidx[pos * 128]

        uint32_t pref = 0;          // value of the key bits decided so far
        int k_rem = p.k;            // rank still to be resolved inside the current prefix
        int ties = 0, cnt = 0;      // collection pass
        uint32_t best_key = 0; int best_idx = 0;   // OUT_TOP1
        if (OUT != OUT_TOP1)
            for (int i = 0; i < NBINS; ++i) hist[i * 128] = 0;

        for (int j = 0; j < J; ++j) {
            const int pass = j / T, tile = j - pass * T;
            const bool radix = pass < NPASS_RADIX;
            const bool mine = (OUT == OUT_IDX && !radix) ? (group == 0) : ((j & 1) == group);
            const int known = pass * RBITS;                                   // key bits decided before this pass
            const int shift = max(32 - known - RBITS, 0);
            const int width = 32 - known - shift;
            const uint32_t bmask = (1u << width) - 1u;
            if (mine) {
                const uint32_t sb = tmem + lane_addr + (uint32_t)(j & 1) * BUF_COLS;
                const int nvalid = Nc - tile * ST_NC;                         // >= 128 except in the last tile
                mbar_wait(bar_s_full + 8 * (j & 1), (j >> 1) & 1);
                tc_fence_after();
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    uint32_t v0[32], v1[32];
                    tmem_ld32(sb + c * 32, v0);
                    tmem_ld32(sb + 128 + c * 32, v1);
                    tmem_ld_wait();
                    const int cbase = tile * ST_NC + c * 32;
                    // ---- keys of the 32 candidates (independent: pipelines well)
#pragma unroll
                    for (int i4 = 0; i4 < 8; ++i4) {
                        float4 xc4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (MODE != SEL_COS) xc4 = __ldg(reinterpret_cast<const float4*>(xxc + cbase + i4 * 4));
                        const float xcs[4] = {xc4.x, xc4.y, xc4.z, xc4.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int i = i4 * 4 + e;
                            const float a0 = __uint_as_float(v0[i]), a1 = __uint_as_float(v1[i]);
                            const float s = (MODE == SEL_PN) ? sel_score<MODE>(a0, a1, inv_s2, xq, xcs[e], p.W)
                                                             : sel_score<MODE>(__fadd_rn(a0, a1), 0.f, inv_s2, xq, xcs[e], p.W);
                            v0[i] = f2ord(s);
                        }
                    }
                    const int nv = nvalid - c * 32;                           // candidates i < nv of this chunk exist
                    if (OUT == OUT_TOP1) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const bool gt = (i < nv) && (v0[i] > best_key);
                            best_key = gt ? v0[i] : best_key;
                            best_idx = gt ? cbase + i : best_idx;
                        }
                    } else if (radix && pass < 2) {
                        // dense passes: nearly every candidate shares the leading bits -> unconditional read-modify-write,
                        // four candidates at a time with duplicates of a bin resolved in registers
#pragma unroll
                        for (int i4 = 0; i4 < 8; ++i4) {
                            uint32_t bn[4], m[4], cn[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int i = i4 * 4 + e;
                                const uint32_t key = v0[i];
                                bn[e] = (key >> shift) & bmask;
                                m[e] = ((i < nv) && (known == 0 || (key >> (32 - known)) == pref)) ? 1u : 0u;
                                cn[e] = hist[bn[e] * 128];
                            }
                            const uint32_t n0 = min(cn[0] + m[0], CMAX);
                            cn[1] = (bn[1] == bn[0]) ? n0 : cn[1];
                            const uint32_t n1 = min(cn[1] + m[1], CMAX);
                            cn[2] = (bn[2] == bn[1]) ? n1 : ((bn[2] == bn[0]) ? n0 : cn[2]);
                            const uint32_t n2 = min(cn[2] + m[2], CMAX);
                            cn[3] = (bn[3] == bn[2]) ? n2 : ((bn[3] == bn[1]) ? n1 : ((bn[3] == bn[0]) ? n0 : cn[3]));
                            const uint32_t n3 = min(cn[3] + m[3], CMAX);
                            hist[bn[0] * 128] = (CT)n0;
                            hist[bn[1] * 128] = (CT)n1;
                            hist[bn[2] * 128] = (CT)n2;
                            hist[bn[3] * 128] = (CT)n3;
                        }
                    } else if (radix) {
                        // sparse passes: few candidates still match the prefix
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const uint32_t key = v0[i];
                            if ((i < nv) && (key >> (32 - known)) == pref) {
                                const uint32_t bin = (key >> shift) & bmask;
                                hist[bin * 128] = (CT)min((uint32_t)hist[bin * 128] + 1u, CMAX);
                            }
                        }
                    } else {
                        // collection: everything above the k-th key, then ties in index order
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const uint32_t key = v0[i];
                            if ((i < nv) && key >= pref) {
                                const bool take = key > pref || ties < k_rem;
                                if (take) {
                                    if (key == pref) ++ties;
                                    if (cnt < ST_KMAX) { lkey[cnt * 128] = key; lidx[cnt * 128] = (uint16_t)(cbase + i); }
                                    ++cnt;
                                }
                            }
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(bar_s_empty + 8 * (j & 1));
            }
            if (OUT != OUT_TOP1 && radix && tile == T - 1) {
                // end of a radix pass: both groups' histograms are complete -> locate the bin holding rank k_rem
                asm volatile("bar.sync 1, 256;" ::: "memory");
                int cum = 0, sel = 0;
                bool found = false;
                for (int bin = (int)bmask; bin >= 0; --bin) {
                    const int cbin = (int)hist[bin * 128] + (int)hist_o[bin * 128];
                    if (!found) {
                        if (cum + cbin >= k_rem) { sel = bin; found = true; }
                        else cum += cbin;
                    }
                }
                pref = (pref << width) | (uint32_t)sel;
                k_rem -= cum;
                asm volatile("bar.sync 1, 256;" ::: "memory");   // everyone has read both histograms
                // (the collection lists alias the histograms: nobody appends before this point)
                if (pass + 1 < NPASS_RADIX)
                    for (int bin = 0; bin < NBINS; ++bin) hist[bin * 128] = 0;
            }
        }

        // ---- results
        if (OUT == OUT_KTH) {
            if (group == 0 && q < p.Nq) p.out_kth[(long long)b * p.Nq + q] = ord2f(pref);
        } else if (OUT == OUT_TOP1) {
            // merge the two groups' candidates (lower index wins ties)
            unsigned long long* slot = reinterpret_cast<unsigned long long*>(sel_ptr) + row;
            const unsigned long long mineK = ((unsigned long long)best_key << 32) | (0xFFFFFFFFu - (uint32_t)best_idx);
            if (group == 1) *slot = mineK;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (group == 0 && q < p.Nq) {
                const unsigned long long o = *slot;
                const unsigned long long best = o > mineK ? o : mineK;
                const int bi = (best >> 32) ? (int)(0xFFFFFFFFu - (uint32_t)(best & 0xFFFFFFFFull)) : 0;
                if (p.idx64) reinterpret_cast<long long*>(p.out_idx)[(long long)b * p.Nq + q] = bi;
                else reinterpret_cast<int*>(p.out_idx)[(long long)b * p.Nq + q] = bi;
            }
        } else {
            // one warp per row: bitonic sort (descending score, ascending index) of the k collected entries;
            // the lists were written by group 0, both groups sort (16 rows per warp)
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const int k = p.k;
            for (int r = group * 16; r < group * 16 + 16; ++r) {
                const int rr = quarter * 32 + r;
                const int qq = q0 + rr;
                if (qq >= p.Nq) break;
                const uint32_t* rk = reinterpret_cast<const uint32_t*>(sel_ptr) + rr;
                const uint16_t* ri = reinterpret_cast<const uint16_t*>(sel_ptr + ST_KMAX * 128 * 4) + rr;
                unsigned long long e[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int pos = lane + 32 * h;
                    e[h] = (pos < k) ? (((unsigned long long)rk[pos * 128] << 32) | (0xFFFFFFFFu - (uint32_t)ri[pos * 128])) : 0ull;
                }
#pragma unroll
                for (int k2 = 2; k2 <= 64; k2 <<= 1) {
#pragma unroll
                    for (int jj = k2 >> 1; jj > 0; jj >>= 1) {
                        if (jj == 32) {
                            // partner is the other element of this lane; final merge: descending
                            if (e[0] < e[1]) { const unsigned long long t = e[0]; e[0] = e[1]; e[1] = t; }
                        } else {
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const int i = lane + 32 * h;
                                const unsigned long long o = shfl_xor_u64(e[h], jj);
                                const bool desc = ((i & k2) == 0);          // block direction (k2 == 64: all descending)
                                const bool lower = ((i & jj) == 0);         // this element sits at the lower index
                                const bool keep_max = (desc == lower);
                                e[h] = keep_max ? (e[h] > o ? e[h] : o) : (e[h] < o ? e[h] : o);
                            }
                        }
                    }
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int pos = lane + 32 * h;
                    if (pos < k) {
                        const int idx = (int)(0xFFFFFFFFu - (uint32_t)(e[h] & 0xFFFFFFFFull));
                        const long long o = ((long long)b * p.Nq + qq) * k + pos;
                        if (p.idx64) reinterpret_cast<long long*>(p.out_idx)[o] = idx;
                        else reinterpret_cast<int*>(p.out_idx)[o] = idx;
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------- streaming top-k
// Single-pass variant for k <= 64 (the kNN graph of the encoder): the Gram runs ONCE.  Each query row is owned by
// one thread that keeps, in shared memory, a private candidate buffer of SS_CAP (score, index) entries and a
// threshold register: a candidate is appended iff its score beats the threshold (predicated stores, no divergence).
// When any buffer of the warp is about to overflow, every lane prunes its own buffer: an interpolation / bisection
// search over the order-preserving integer image of the scores finds a pivot that keeps between k and k + SS_WIN
// entries, the buffer is compacted in place (index order preserved, so ties keep going to the lowest index) and the
// pivot becomes the new threshold.  A threshold only has to guarantee that >= k earlier candidates are at least as
// good, so the prune need not be exact; the final prune (window 0) is, and the k survivors of a row are sorted by one
// warp.  Expected appends per row ~ k ln(N/k) (a few hundred of 10 000 candidates); ~6 prunes per row.
// Candidate tiles are 64 wide (two 32-candidate chunks per thread), four TMEM buffers deep; padding candidates carry a
// squared norm of +inf, so their score is -inf (or NaN) and never passes the threshold test.
constexpr int SS_THREADS = 192;   // warp 0 TMA, warp 1 MMA, warps 2-5 select (thread per row)
constexpr int SS_NC = 64;         // candidates per tile
constexpr int SS_CAP = 228;       // entries per row buffer (row-major; 912-B / 456-B row strides keep LDS.128 / LDS.64 conflict-free)
constexpr int SS_WIN = 12;        // a prune leaves between k and k + SS_WIN entries (swept on the GPU: tools/sweep_prune*.py)
constexpr int SS_SOFT = 164;      // soft mark: book a CTA-wide prune
constexpr int SS_LAG = 3;         // ... this many tiles ahead
constexpr int SS_STAGES = 3;
constexpr int SS_NBUF = 3;        // TMEM buffers of 128 columns (two 64-column accumulators); columns [384,448): the query tile
constexpr int SS_XCRING = 8;      // candidate-norm slices in flight: >= SS_STAGES + SS_NBUF (producer's maximum lead)
constexpr uint32_t SS_XPART = 64 * 128;                  // one candidate tile part: 64 rows x 64 channels fp16
constexpr uint32_t SS_KEY_ROW = SS_CAP * 4, SS_IDX_ROW = SS_CAP * 2;
constexpr uint32_t SS_KEY_BYTES = SS_KEY_ROW * 128, SS_IDX_BYTES = SS_IDX_ROW * 128;
constexpr uint32_t SS_NEG_INF = 0xff800000u;

__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
    uint32_t v;
    asm volatile("{ .reg .u16 t; ld.shared.u16 t, [%1]; cvt.u32.u16 %0, t; }" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v)); }
__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v) {
    asm volatile("{ .reg .u16 t; cvt.u16.u32 t, %1; st.shared.u16 [%0], t; }" ::"r"(a), "r"(v));
}
// The prune-booking words are shared between warps without a barrier on purpose (a hint that arrives late only moves a
// prune to a later tile).  They are read and written with shared-memory atomics -- one lane per warp, the value broadcast
// by a shuffle -- so that concurrent access is well defined (and compute-sanitizer's racecheck has nothing to report).
__device__ __forceinline__ uint32_t hint_load(uint32_t a) {
    uint32_t v = 0;
    if ((threadIdx.x & 31) == 0) asm volatile("atom.shared.or.b32 %0, [%1], 0;" : "=r"(v) : "r"(a) : "memory");
    return __shfl_sync(0xffffffffu, v, 0);
}
__device__ __forceinline__ void hint_store(uint32_t a, uint32_t v) {
    uint32_t old;
    asm volatile("atom.shared.exch.b32 %0, [%1], %2;" : "=r"(old) : "r"(a), "r"(v) : "memory");
}
template <int NS>
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {   // producer-side wait: back off
    uint32_t ok;
    for (;;) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) break;
        __nanosleep(NS);
    }
}

__device__ __forceinline__ void lds_v4(uint32_t a, float (&v)[4]) {
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(a));
}

// number of entries of a row buffer (uniform bound cmax, a multiple of 4; padding entries are -inf) with
// key >= pivot / key > pivot.  Rows are contiguous, so one LDS.128 brings four entries.
template <bool STRICT>
__device__ __forceinline__ int ss_count(uint32_t key_row, int cmax, float pivot) {
    int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
#pragma unroll 4
    for (int i = 0; i < cmax; i += 4) {
        float v[4];
        lds_v4(key_row + (uint32_t)i * 4u, v);
        c0 += (STRICT ? (v[0] > pivot) : (v[0] >= pivot)) ? 1 : 0;
        c1 += (STRICT ? (v[1] > pivot) : (v[1] >= pivot)) ? 1 : 0;
        c2 += (STRICT ? (v[2] > pivot) : (v[2] >= pivot)) ? 1 : 0;
        c3 += (STRICT ? (v[3] > pivot) : (v[3] >= pivot)) ? 1 : 0;
    }
    return (c0 + c1) + (c2 + c3);
}

// Warp-wide prune of the 32 private buffers (all lanes call; lanes with cnt <= k + win keep everything).
// On return every lane holds between k and k + win entries (exactly min(cnt, k) for win == 0), in index order,
// slots >= cnt are -inf, and thr is a valid threshold (>= k kept entries are >= thr).
template <bool HAS_IDX>
__device__ __noinline__ void ss_prune(uint32_t key_row, uint32_t idx_row, int cap, int k, int win, int& cnt_io, float& thr_io) {
    int cnt = cnt_io;
    float thr = thr_io;
    const int cmax = min((__reduce_max_sync(0xffffffffu, cnt) + 3) & ~3, cap);
    const bool active = cnt > k + win;
    float mx;
    {
        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll 4
        for (int i = 0; i < cmax; i += 4) {
            float v[4];
            lds_v4(key_row + (uint32_t)i * 4u, v);
            m0 = fmaxf(m0, v[0]); m1 = fmaxf(m1, v[1]); m2 = fmaxf(m2, v[2]); m3 = fmaxf(m3, v[3]);
        }
        mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
    }
    float lo_f = thr;
    if (!(thr > -INFINITY)) {   // first prune of the row: every entry so far was accepted; start from the minimum
        float mn = INFINITY;
        for (int i = 0; i < cnt; ++i) mn = fminf(mn, __uint_as_float(lds_u32(key_row + (uint32_t)i * 4u)));
        lo_f = mn;
    }
    // invariant: count(>= lo) = c_lo >= k, count(>= hi) = c_hi < k (hi starts one step above the maximum)
    uint32_t lo_u = f2ord(lo_f), hi_u = f2ord(mx) + 1u;
    int c_lo = cnt, c_hi = 0;
    const float target = (float)k + 0.5f * (float)win + 0.5f;
    for (int it = 0; it < 80; ++it) {
        const bool need = active && c_lo > k + win && hi_u - lo_u > 1u;
        if (!__any_sync(0xffffffffu, need)) break;
        uint32_t mid_u = lo_u + ((hi_u - lo_u) >> 1);
        if (it < 3 || (it & 1)) {   // interpolate on the counts (the tail of the score distribution is near-linear)
            const float lf = ord2f(lo_u), hf = ord2f(hi_u - 1u);
            const float frac = ((float)c_lo - target) / (float)(c_lo - c_hi);
            const uint32_t g_u = f2ord(fmaf(hf - lf, frac, lf));
            if (need && g_u > lo_u && g_u < hi_u) mid_u = g_u;
        }
        const float mid_f = ord2f(mid_u);
        const int c = ss_count<false>(key_row, cmax, mid_f);
        if (need) {
            if (c >= k) { lo_u = mid_u; lo_f = mid_f; c_lo = c; }
            else { hi_u = mid_u; c_hi = c; }
        }
    }
    // compaction: keep key > lo, and ties at lo in index order up to the quota (unlimited unless the search ended on a
    // tie plateau wider than the window)
    const bool plateau = active && c_lo > k + win;
    int quota = 0x7fffffff;
    if (__any_sync(0xffffffffu, plateau)) {
        const int g = ss_count<true>(key_row, cmax, lo_f);
        if (plateau) quota = max(k - g, 0);
    }
    if (active) {
        int w = 0, t = 0;
        const int cend = (cnt + 3) & ~3;
        for (int r = 0; r < cend; r += 4) {
            float v[4];
            uint32_t i01 = 0, i23 = 0;
            lds_v4(key_row + (uint32_t)r * 4u, v);
            if (HAS_IDX)
                asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(i01), "=r"(i23) : "r"(idx_row + (uint32_t)r * 2u));
            asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(key_row + (uint32_t)r * 4u), "r"(SS_NEG_INF));
            const uint32_t iv[4] = {i01 & 0xffffu, i01 >> 16, i23 & 0xffffu, i23 >> 16};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const bool tie = (v[e] == lo_f);
                const bool keep = (v[e] > lo_f) || (tie && t < quota);   // padding (-inf) never kept
                t += tie ? 1 : 0;
                if (keep) {
                    sts_u32(key_row + (uint32_t)w * 4u, __float_as_uint(v[e]));
                    if (HAS_IDX) sts_u16(idx_row + (uint32_t)w * 2u, iv[e]);
                    ++w;
                }
            }
        }
        cnt = w;
        thr = lo_f;
    }
    __syncwarp();
    cnt_io = cnt;
    thr_io = thr;
}

// The query tile (FP16 hi and lo, 64 channels) lives in TMEM for the whole CTA (A operand from TMEM, as in the K-th
// kernel below): the 32 KB of shared memory it would take go to the row buffers instead.
template <int MODE>
__global__ void __launch_bounds__(SS_THREADS, 1)
select_stream_kernel(const __grid_constant__ CUtensorMap map_xh, const __grid_constant__ CUtensorMap map_xl,
                     const __half* __restrict__ qh, const __half* __restrict__ ql, SelParams p) {
    constexpr uint32_t XSTAGE = 2 * SS_XPART;                      // hi + lo
    constexpr uint32_t BUF_COLS = 128;
    constexpr uint32_t Q_COL = SS_NBUF * BUF_COLS;                 // 384: hi in [384,416), lo in [416,448)

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw0 = smem_u32(smem_raw);
    const uint32_t smem_base = (raw0 + 1023u) & ~1023u;
    const uint32_t x_addr = smem_base;
    const uint32_t list_addr = x_addr + SS_STAGES * XSTAGE;
    const uint32_t xc_addr = list_addr + SS_KEY_BYTES + SS_IDX_BYTES;       // [SS_XCRING][64] candidate squared norms
    const uint32_t rc_addr = xc_addr + SS_XCRING * SS_NC * 4;               // [128] final per-row counts
    const uint32_t bar_base = rc_addr + ST_M * 4;
    const uint32_t bar_q_full = bar_base;
    const uint32_t bar_x_full = bar_base + 8;
    const uint32_t bar_x_empty = bar_x_full + 8 * SS_STAGES;
    const uint32_t bar_s_full = bar_x_empty + 8 * SS_STAGES;   // [SS_NBUF]
    const uint32_t bar_s_empty = bar_s_full + 8 * SS_NBUF;     // [SS_NBUF]
    const uint32_t tmem_slot = bar_s_empty + 8 * SS_NBUF;
    const uint32_t sched_addr = tmem_slot + 8;                 // [16] tile index of a scheduled CTA-wide prune
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw0));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y, q0 = blockIdx.x * ST_M;
    const int Nc = p.Nc;
    const int T = (Nc + SS_NC - 1) / SS_NC;
    if (threadIdx.x < 16) sts_u32(sched_addr + threadIdx.x * 4, 0xffffffffu);

    if (threadIdx.x == 0) {
        mbar_init(bar_q_full, 128);
        for (int s = 0; s < SS_STAGES; ++s) { mbar_init(bar_x_full + 8 * s, 1); mbar_init(bar_x_empty + 8 * s, 1); }
        for (int i = 0; i < SS_NBUF; ++i) { mbar_init(bar_s_full + 8 * i, 1); mbar_init(bar_s_empty + 8 * i, 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tmem_slot) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot_ptr;

    if (warp == 0) {
        // ============================================================ TMA producer
        if (elect_one()) {
            for (int j = 0; j < T; ++j) {
                const int s = j % SS_STAGES;
                if (j >= SS_STAGES) mbar_wait_relaxed<256>(bar_x_empty + 8 * s, ((j / SS_STAGES) - 1) & 1);
                const uint32_t dst = x_addr + s * XSTAGE, bar = bar_x_full + 8 * s;
                mbar_expect_tx(bar, XSTAGE + SS_NC * 4);
                tma_load_3d(dst, &map_xh, bar, 0, j * SS_NC, b);
                tma_load_3d(dst + SS_XPART, &map_xl, bar, 0, j * SS_NC, b);
                // the tile's squared norms ride on the same barrier; their ring slot outlives the operand stage
                // (the MMA frees the stage, the selection threads read the norms up to SS_NBUF tiles later)
                bulk_load(xc_addr + (uint32_t)(j % SS_XCRING) * (SS_NC * 4), p.xxc + (long long)b * p.npad + j * SS_NC,
                          SS_NC * 4, bar);
            }
        }
    } else if (warp == 1) {
        // ============================================================ MMA issuer
        if (elect_one()) {
            constexpr uint32_t IDESC = make_idesc_n(0, SS_NC);
            mbar_wait_relaxed<32>(bar_q_full, 0);
            for (int j = 0; j < T; ++j) {
                const int s = j % SS_STAGES, buf = j % SS_NBUF;
                mbar_wait_relaxed<32>(bar_x_full + 8 * s, (j / SS_STAGES) & 1);
                if (j >= SS_NBUF) mbar_wait_relaxed<32>(bar_s_empty + 8 * buf, ((j / SS_NBUF) - 1) & 1);
                tc_fence_after();
                const uint32_t xs = x_addr + s * XSTAGE;
                const uint32_t d = tmem + (uint32_t)buf * BUF_COLS;
                if (MODE == SEL_PN) {
#pragma unroll
                    for (int a = 0; a < 2; ++a) {
                        // channels [0,16): points (3 used), [16,32): normals (3 used): one K-step (8 TMEM columns) each
#pragma unroll
                        for (int term = 0; term < 3; ++term) {
                            const uint32_t qa = tmem + Q_COL + ((term == 2) ? 32u : 0u);    // Qh, Qh, Ql   (TMEM)
                            const uint32_t xb = xs + ((term == 1) ? SS_XPART : 0);          // Xh, Xl, Xh
                            umma_ts(d + a * 64, qa + a * 8, make_desc(xb + a * 32, 16), IDESC, term > 0);
                        }
                    }
                } else {
#pragma unroll
                    for (int term = 0; term < 3; ++term) {
                        const uint32_t qa = tmem + Q_COL + ((term == 2) ? 32u : 0u);
                        const uint32_t xb = xs + ((term == 1) ? SS_XPART : 0);
                        const uint32_t dd = d + (term == 0 ? 0u : 64u);                     // hi.hi | cross terms
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            umma_ts(dd, qa + ks * 8, make_desc(xb + ks * 32, 16), IDESC, (ks > 0 || term == 2) ? 1u : 0u);
                    }
                }
                tc_commit(bar_s_full + 8 * buf);
                tc_commit(bar_x_empty + 8 * s);
            }
        }
    } else {
        // ============================================================ selection: one thread per query row
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const int q = q0 + row;
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        {   // this thread's query row -> TMEM columns [384,416) (hi) and [416,448) (lo)
            const long long ro = ((long long)b * p.Nq + min(q, p.Nq - 1)) * 64;
#pragma unroll
            for (int part = 0; part < 2; ++part) {
                const uint4* g4 = reinterpret_cast<const uint4*>((part ? ql : qh) + ro);
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t w[16];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint4 v = make_uint4(0u, 0u, 0u, 0u);
                        if (q < p.Nq) v = __ldg(g4 + c * 4 + i);
                        w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
                    }
                    tmem_st16(tmem + lane_addr + Q_COL + part * 32u + c * 16u, w);
                }
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(bar_q_full);
        }
        const float sq = scale_from_maxabs(p.maxabs_q[b]);
        const float sc = scale_from_maxabs(p.maxabs_c[b]);
        const float inv2 = 2.0f * ((1.0f / sq) * (1.0f / sc));          // power of two: products below are exact
        const float xq = p.xxq[(long long)b * p.npad + min(q, p.npad - 1)];
        const float W = p.W;
        const int k = p.k;
        const uint32_t key_addr = list_addr + (uint32_t)row * SS_KEY_ROW;                  // key[pos] at + pos * 4
        const uint32_t idx_addr = list_addr + SS_KEY_BYTES + (uint32_t)row * SS_IDX_ROW;   // idx[pos] at + pos * 2
        float thr = -INFINITY;
        int cnt = 0;
        for (int i = 0; i < SS_CAP; i += 4)
            asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(key_addr + (uint32_t)i * 4u), "r"(SS_NEG_INF));

        auto maybe_prune = [&]() {
            if (__any_sync(0xffffffffu, cnt > SS_CAP - 32)) ss_prune<true>(key_addr, idx_addr, SS_CAP, k, p.win, cnt, thr);
        };
        // one 32-candidate chunk of this thread's row: v0 / v1 = the two accumulators
        auto process = [&](const uint32_t (&v0)[32], const uint32_t (&v1)[32], int cbase, uint32_t xc_slot) {
            float xc[32];   // the chunk's candidate norms, all loads up front
#pragma unroll
            for (int i4 = 0; i4 < 8; ++i4)
                asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                    : "=f"(xc[i4 * 4]), "=f"(xc[i4 * 4 + 1]), "=f"(xc[i4 * 4 + 2]), "=f"(xc[i4 * 4 + 3]) : "r"(xc_slot + i4 * 16));
#pragma unroll
            for (int i4 = 0; i4 < 8; ++i4) {
                const float xcs[4] = {xc[i4 * 4], xc[i4 * 4 + 1], xc[i4 * 4 + 2], xc[i4 * 4 + 3]};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int i = i4 * 4 + e;
                    const float a0 = __uint_as_float(v0[i]), a1 = __uint_as_float(v1[i]);
                    float s;
                    if (MODE == SEL_PN) {
                        // src/PointNet.py:112-120,128: p = (xx_j - 2 p.p') + xx_i ; n = 2 - 2 n.n' ; -(p * (1 + n * W))
                        const float pd = __fadd_rn(fmaf(-inv2, a0, xcs[e]), xq);
                        const float nd = fmaf(-inv2, a1, 2.0f);
                        s = -__fmul_rn(pd, __fadd_rn(1.0f, __fmul_rn(nd, W)));
                    } else {
                        // src/PointNet.py:76-78: inner = -2 x.x' ; pd = (-xx_j - inner) - xx_i
                        s = __fsub_rn(fmaf(inv2, __fadd_rn(a0, a1), -xcs[e]), xq);
                    }
                    if (s > thr) {   // addresses from cnt in fresh registers: no write-after-read wait on the stores
                        sts_u32(key_addr + (uint32_t)cnt * 4u, __float_as_uint(s));
                        sts_u16(idx_addr + (uint32_t)cnt * 2u, (uint32_t)(cbase + i));
                        ++cnt;
                    }
                }
            }
        };

        uint32_t a0[32], a1[32], b0[32], b1[32];
        mbar_wait(bar_s_full, 0);
        tc_fence_after();
        tmem_ld32(tmem + lane_addr, a0);
        tmem_ld32(tmem + lane_addr + 64, a1);
#pragma unroll 1
        for (int j = 0; j < T; ++j) {
            const int buf = j % SS_NBUF;
            const uint32_t sb = tmem + lane_addr + (uint32_t)buf * BUF_COLS;
            tmem_ld_wait();                         // chunk 0 of tile j has landed in (a0, a1)
            tmem_ld32(sb + 32, b0);                 // chunk 1 loads while chunk 0 is processed
            tmem_ld32(sb + 64 + 32, b1);
            const uint32_t xcs = xc_addr + (uint32_t)(j % SS_XCRING) * (SS_NC * 4);
            // CTA-wide prunes: the four warps share the TMEM ring, so a warp that prunes alone stalls the others after
            // ~2 tiles.  A warp whose buffers pass the soft mark books a prune SS_LAG tiles ahead (further than the
            // warps can drift apart); every warp prunes when it reaches the booked tile.  The private prune of
            // maybe_prune() stays as the overflow guard.  The booking words are hints written and polled without a barrier
            // on purpose, through shared-memory atomics (hint_load / hint_store): a missed hint only moves a warp's
            // prune to a later tile, and the result does not depend on when a row is pruned.
            if (hint_load(sched_addr + (uint32_t)(j & 15) * 4u) == (uint32_t)j) {
                ss_prune<true>(key_addr, idx_addr, SS_CAP, k, p.win, cnt, thr);
            } else if (__any_sync(0xffffffffu, cnt > p.soft)) {
                bool booked = false;
#pragma unroll
                for (int d = 1; d <= SS_LAG; ++d) booked |= hint_load(sched_addr + (uint32_t)((j + d) & 15) * 4u) == (uint32_t)(j + d);
                if (!booked && lane == 0) hint_store(sched_addr + (uint32_t)((j + SS_LAG) & 15) * 4u, (uint32_t)(j + SS_LAG));
            }
            maybe_prune();
            process(a0, a1, j * SS_NC, xcs);
            tmem_ld_wait();
            if (j + 1 < T) {                        // chunk 0 of tile j + 1 loads while chunk 1 is processed
                const int nb = (j + 1) % SS_NBUF;
                mbar_wait(bar_s_full + 8 * nb, ((j + 1) / SS_NBUF) & 1);
                tc_fence_after();
                const uint32_t sn = tmem + lane_addr + (uint32_t)nb * BUF_COLS;
                tmem_ld32(sn, a0);
                tmem_ld32(sn + 64, a1);
            }
            maybe_prune();
            process(b0, b1, j * SS_NC + 32, xcs + 128);
            tc_fence_before();
            mbar_arrive(bar_s_empty + 8 * buf);    // every tcgen05.ld of tile j has completed (both waits above)
        }

        // ---- exact top-k of the survivors; the per-row counts go to shared memory (the Q tile is dead: every MMA
        // completed before the last s_full) for the sort below
        ss_prune<true>(key_addr, idx_addr, SS_CAP, k, 0, cnt, thr);
        sts_u32(rc_addr + (uint32_t)row * 4u, (uint32_t)cnt);
    }
    // ================================================================ sort: one warp per row, all six warps
    __syncwarp();
    asm volatile("bar.sync 2, 192;" ::: "memory");
    {
        const int k = p.k;
        for (int rr = warp; rr < ST_M; rr += SS_THREADS / 32) {
            const int qq = q0 + rr;
            if (qq >= p.Nq) break;
            const int rcnt = (int)lds_u32(rc_addr + (uint32_t)rr * 4u);
            const uint32_t rk = list_addr + (uint32_t)rr * SS_KEY_ROW, ri = list_addr + SS_KEY_BYTES + (uint32_t)rr * SS_IDX_ROW;
            if (p.unsorted) {   // consumers that reduce over the neighbours do not need the order: skip the sort
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int pos = lane + 32 * h;
                    if (pos < k) {
                        const int idx = pos < rcnt ? (int)lds_u16(ri + (uint32_t)pos * 2u) : 0;
                        const long long o = ((long long)b * p.Nq + qq) * k + pos;
                        if (p.idx64) reinterpret_cast<long long*>(p.out_idx)[o] = idx;
                        else reinterpret_cast<int*>(p.out_idx)[o] = idx;
                    }
                }
                continue;
            }
            unsigned long long e[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int pos = lane + 32 * h;
                e[h] = (pos < rcnt) ? (((unsigned long long)f2ord(__uint_as_float(lds_u32(rk + (uint32_t)pos * 4u))) << 32) |
                                       (0xFFFFFFFFu - lds_u16(ri + (uint32_t)pos * 2u)))
                                    : 0ull;
            }
#pragma unroll
            for (int k2 = 2; k2 <= 64; k2 <<= 1) {
#pragma unroll
                for (int jj = k2 >> 1; jj > 0; jj >>= 1) {
                    if (jj == 32) {
                        if (e[0] < e[1]) { const unsigned long long t = e[0]; e[0] = e[1]; e[1] = t; }
                    } else {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int i = lane + 32 * h;
                            const unsigned long long o = shfl_xor_u64(e[h], jj);
                            const bool desc = ((i & k2) == 0);
                            const bool lower = ((i & jj) == 0);
                            const bool keep_max = (desc == lower);
                            e[h] = keep_max ? (e[h] > o ? e[h] : o) : (e[h] < o ? e[h] : o);
                        }
                    }
                }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int pos = lane + 32 * h;
                if (pos < k) {
                    const int idx = (e[h] >> 32) ? (int)(0xFFFFFFFFu - (uint32_t)(e[h] & 0xFFFFFFFFull)) : 0;
                    const long long o = ((long long)b * p.Nq + qq) * k + pos;
                    if (p.idx64) reinterpret_cast<long long*>(p.out_idx)[o] = idx;
                    else reinterpret_cast<int*>(p.out_idx)[o] = idx;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

template <int MODE>
static int launch_stream(const CUtensorMap& xh, const CUtensorMap& xl, const __half* qh, const __half* ql,
                         const SelParams& p, int B, cudaStream_t st) {
    constexpr size_t smem = (size_t)SS_STAGES * 2 * SS_XPART + SS_KEY_BYTES + SS_IDX_BYTES + SS_XCRING * SS_NC * 4 +
                            ST_M * 4 + 1024 + 256;
    static_assert(smem <= 227 * 1024, "shared memory budget");
    auto kern = select_stream_kernel<MODE>;
    SED_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((p.Nq + ST_M - 1) / ST_M, B);
    kern<<<grid, SS_THREADS, smem, st>>>(xh, xl, qh, ql, p);
    SED_CHECK_LAUNCH();
    return SED_OK;
}


// ---------------------------------------------------------------------------------------------- streaming K-th score
// compute_bandwidth (src/mean_shift.py:130-135): the K-th largest cosine score (= K-th smallest distance) of every
// row, K up to KS_KMAX.  Same single-pass scheme as the kNN kernel above, for 128-channel rows: the query tile (hi and
// lo parts) lives in TMEM for the whole CTA (A operand from TMEM), which leaves shared memory to three 32 KB candidate
// stages and 252-entry row buffers; only the scores are kept (no indices).
constexpr int KS_CAP = 316;       // 1264-B row stride: LDS.128 conflict-free (prunes dominate this kernel: buffers as large as shared memory allows)
constexpr int KS_KMAX = 200;
constexpr int KS_WIN = 12;
constexpr int KS_SOFT = 252;       // soft mark of the CTA-wide booked prunes (0 = off; SEDNET_B200_KS_SOFT; swept on the GPU)
constexpr int KS_STAGES = 2;       // the MMA (768 cycles per tile) paces the stream: two candidate stages suffice
constexpr int KS_NBUF = 3;        // TMEM: 3 x 128 accumulator columns + 128 columns of Q
constexpr uint32_t KS_XPART = 2 * SS_XPART;              // 64 rows x 128 channels fp16 (two 64-channel boxes)

__global__ void __launch_bounds__(SS_THREADS, 1)
kth_stream_kernel(const __grid_constant__ CUtensorMap map_xh, const __grid_constant__ CUtensorMap map_xl,
                  const __half* __restrict__ qh, const __half* __restrict__ ql, SelParams p) {
    constexpr uint32_t XSTAGE = 2 * KS_XPART;
    constexpr uint32_t BUF_COLS = 128;
    constexpr uint32_t KEY_ROW = KS_CAP * 4;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw0 = smem_u32(smem_raw);
    const uint32_t smem_base = (raw0 + 1023u) & ~1023u;
    const uint32_t x_addr = smem_base;
    const uint32_t list_addr = x_addr + KS_STAGES * XSTAGE;
    const uint32_t bar_base = list_addr + KEY_ROW * 128;
    const uint32_t bar_q_full = bar_base;
    const uint32_t bar_x_full = bar_base + 8;
    const uint32_t bar_x_empty = bar_x_full + 8 * KS_STAGES;
    const uint32_t bar_s_full = bar_x_empty + 8 * KS_STAGES;   // [KS_NBUF]
    const uint32_t bar_s_empty = bar_s_full + 8 * KS_NBUF;     // [KS_NBUF]
    const uint32_t tmem_slot = bar_s_empty + 8 * KS_NBUF;
    const uint32_t sched_addr = tmem_slot + 8;                 // [16] tile index of a booked CTA-wide prune (as in the kNN kernel)
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw0));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y, q0 = blockIdx.x * ST_M;
    const int Nc = p.Nc;
    const int T = (Nc + SS_NC - 1) / SS_NC;
    if (threadIdx.x < 16) sts_u32(sched_addr + threadIdx.x * 4, 0xffffffffu);

    if (threadIdx.x == 0) {
        mbar_init(bar_q_full, 128);
        for (int s = 0; s < KS_STAGES; ++s) { mbar_init(bar_x_full + 8 * s, 1); mbar_init(bar_x_empty + 8 * s, 1); }
        for (int i = 0; i < KS_NBUF; ++i) { mbar_init(bar_s_full + 8 * i, 1); mbar_init(bar_s_empty + 8 * i, 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tmem_slot) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot_ptr;

    if (warp == 0) {
        // ============================================================ TMA producer
        if (elect_one()) {
            for (int j = 0; j < T; ++j) {
                const int s = j % KS_STAGES;
                if (j >= KS_STAGES) mbar_wait_relaxed<256>(bar_x_empty + 8 * s, ((j / KS_STAGES) - 1) & 1);
                const uint32_t dst = x_addr + s * XSTAGE, bar = bar_x_full + 8 * s;
                mbar_expect_tx(bar, XSTAGE);
                tma_load_3d(dst, &map_xh, bar, 0, j * SS_NC, b);
                tma_load_3d(dst + SS_XPART, &map_xh, bar, 64, j * SS_NC, b);
                tma_load_3d(dst + KS_XPART, &map_xl, bar, 0, j * SS_NC, b);
                tma_load_3d(dst + KS_XPART + SS_XPART, &map_xl, bar, 64, j * SS_NC, b);
            }
        }
    } else if (warp == 1) {
        // ============================================================ MMA issuer
        if (elect_one()) {
            constexpr uint32_t IDESC = make_idesc_n(0, SS_NC);
            mbar_wait_relaxed<32>(bar_q_full, 0);
            for (int j = 0; j < T; ++j) {
                const int s = j % KS_STAGES, buf = j % KS_NBUF;
                mbar_wait_relaxed<32>(bar_x_full + 8 * s, (j / KS_STAGES) & 1);
                if (j >= KS_NBUF) mbar_wait_relaxed<32>(bar_s_empty + 8 * buf, ((j / KS_NBUF) - 1) & 1);
                tc_fence_after();
                const uint32_t xs = x_addr + s * XSTAGE;
                const uint32_t d = tmem + (uint32_t)buf * BUF_COLS;
#pragma unroll
                for (int term = 0; term < 3; ++term) {
                    const uint32_t qa = tmem + 384u + ((term == 2) ? 64u : 0u);         // Qh, Qh, Ql   (TMEM)
                    const uint32_t xb = xs + ((term == 1) ? KS_XPART : 0);              // Xh, Xl, Xh
                    const uint32_t dd = d + (term == 0 ? 0u : 64u);                     // hi.hi | cross terms
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks)
                        umma_ts(dd, qa + ks * 8, make_desc(xb + (ks >> 2) * SS_XPART + (ks & 3) * 32, 16), IDESC,
                                (ks > 0 || term == 2) ? 1u : 0u);
                }
                tc_commit(bar_s_full + 8 * buf);
                tc_commit(bar_x_empty + 8 * s);
            }
        }
    } else {
        // ============================================================ selection: one thread per query row
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const int q = q0 + row;
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        {   // this thread's query row -> TMEM columns [384,448) (hi) and [448,512) (lo)
            const long long ro = ((long long)b * p.Nq + min(q, p.Nq - 1)) * 128;
#pragma unroll
            for (int part = 0; part < 2; ++part) {
                const uint4* g4 = reinterpret_cast<const uint4*>((part ? ql : qh) + ro);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t w[16];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint4 v = make_uint4(0u, 0u, 0u, 0u);
                        if (q < p.Nq) v = __ldg(g4 + c * 4 + i);
                        w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
                    }
                    tmem_st16(tmem + lane_addr + 384u + part * 64u + c * 16u, w);
                }
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(bar_q_full);
        }
        const float inv2 = 2.0f / (p.fixed_scale * p.fixed_scale);       // power of two
        const int k = p.k;
        const uint32_t key_addr = list_addr + (uint32_t)row * KEY_ROW;
        float thr = -INFINITY;
        int cnt = 0;
        for (int i = 0; i < KS_CAP; i += 4)
            asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(key_addr + (uint32_t)i * 4u), "r"(SS_NEG_INF));

        auto maybe_prune = [&]() {
            if (__any_sync(0xffffffffu, cnt > KS_CAP - 32)) ss_prune<false>(key_addr, 0u, KS_CAP, k, p.win, cnt, thr);
        };
        // src/mean_shift.py:130: dist = 2 - 2 x.y ; score = -dist.  nv: valid candidates of the chunk (tail masking)
        auto process = [&](const uint32_t (&v0)[32], const uint32_t (&v1)[32], int nv) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float s = fmaf(inv2, __fadd_rn(__uint_as_float(v0[i]), __uint_as_float(v1[i])), -2.0f);
                if ((s > thr) && (i < nv)) {
                    sts_u32(key_addr + (uint32_t)cnt * 4u, __float_as_uint(s));
                    ++cnt;
                }
            }
        };

        uint32_t a0[32], a1[32], b0[32], b1[32];
        mbar_wait(bar_s_full, 0);
        tc_fence_after();
        tmem_ld32(tmem + lane_addr, a0);
        tmem_ld32(tmem + lane_addr + 64, a1);
#pragma unroll 1
        for (int j = 0; j < T; ++j) {
            const int buf = j % KS_NBUF;
            const uint32_t sb = tmem + lane_addr + (uint32_t)buf * BUF_COLS;
            const int nvalid = Nc - j * SS_NC;
            tmem_ld_wait();
            tmem_ld32(sb + 32, b0);
            tmem_ld32(sb + 64 + 32, b1);
            // CTA-wide booked prunes (see select_stream_kernel): the four warps share the TMEM ring, so they prune together
            if (p.soft > 0) {
                if (hint_load(sched_addr + (uint32_t)(j & 15) * 4u) == (uint32_t)j) {
                    ss_prune<false>(key_addr, 0u, KS_CAP, k, p.win, cnt, thr);
                } else if (__any_sync(0xffffffffu, cnt > p.soft)) {
                    bool booked = false;
#pragma unroll
                    for (int d = 1; d <= SS_LAG; ++d) booked |= hint_load(sched_addr + (uint32_t)((j + d) & 15) * 4u) == (uint32_t)(j + d);
                    if (!booked && lane == 0) hint_store(sched_addr + (uint32_t)((j + SS_LAG) & 15) * 4u, (uint32_t)(j + SS_LAG));
                }
            }
            maybe_prune();
            process(a0, a1, nvalid);
            tmem_ld_wait();
            if (j + 1 < T) {
                const int nb = (j + 1) % KS_NBUF;
                mbar_wait(bar_s_full + 8 * nb, ((j + 1) / KS_NBUF) & 1);
                tc_fence_after();
                const uint32_t sn = tmem + lane_addr + (uint32_t)nb * BUF_COLS;
                tmem_ld32(sn, a0);
                tmem_ld32(sn + 64, a1);
            }
            maybe_prune();
            process(b0, b1, nvalid - 32);
            tc_fence_before();
            mbar_arrive(bar_s_empty + 8 * buf);
        }
        // ---- exact K-th: prune to exactly K entries, their minimum is the answer
        ss_prune<false>(key_addr, 0u, KS_CAP, k, 0, cnt, thr);
        float mn = INFINITY;
        for (int i = 0; i < cnt; ++i) mn = fminf(mn, __uint_as_float(lds_u32(key_addr + (uint32_t)i * 4u)));
        if (q < p.Nq) p.out_kth[(long long)b * p.Nq + q] = mn;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

static int launch_kth_stream(const CUtensorMap& xh, const CUtensorMap& xl, const __half* qh, const __half* ql,
                             const SelParams& p, int B, cudaStream_t st) {
    constexpr size_t smem = (size_t)KS_STAGES * 2 * KS_XPART + (size_t)KS_CAP * 4 * 128 + 1024 + 256;
    static_assert(smem <= 227 * 1024, "shared memory budget");
    SED_CUDA(cudaFuncSetAttribute(kth_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((p.Nq + ST_M - 1) / ST_M, B);
    kth_stream_kernel<<<grid, SS_THREADS, smem, st>>>(xh, xl, qh, ql, p);
    SED_CHECK_LAUNCH();
    return SED_OK;
}

// ---------------------------------------------------------------------------------------------- operand packing
__global__ void maxabs_kernel(const float* __restrict__ x, long long bstride, long long n, float* __restrict__ out) {
    const int b = blockIdx.y;
    const float* xb = x + (long long)b * bstride;
    float m = 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        m = fmaxf(m, fabsf(xb[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<unsigned int*>(out + b), __float_as_uint(m));
}

// channel-major x (B,C,N) -> fp16 hi/lo rows [B][N][64] (scaled), FP32 squared norms (B,npad).
// pn: channels 0-2 -> slots 0-2 (points), 3-5 -> slots 16-18 (normals); norms over the points only.
__global__ void pack_cm_kernel(const float* __restrict__ x, long long bstride, int C, int N, int npad, int pn,
                               const float* __restrict__ maxabs, __half* __restrict__ hi, __half* __restrict__ lo,
                               float* __restrict__ xx) {
    const int b = blockIdx.y, n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= npad) return;
    if (n >= N) { xx[(long long)b * npad + n] = INFINITY; return; }   // padding: score -inf, never selected
    const float s = scale_from_maxabs(maxabs[b]);
    const float* xb = x + (long long)b * bstride + n;
    __half* ho = hi + ((long long)b * N + n) * 64;
    __half* lw = lo + ((long long)b * N + n) * 64;
    float nrm = 0.f;
#pragma unroll 1
    for (int g = 0; g < 8; ++g) {
        __align__(16) __half h[8];
        __align__(16) __half l[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int slot = g * 8 + e;
            int c = slot;
            if (pn) c = (slot < 3) ? slot : ((slot >= 16 && slot < 19) ? slot - 13 : -1);
            float v = (c >= 0 && c < C) ? xb[(long long)c * N] : 0.f;
            if (c >= 0 && c < C && (!pn || c < 3)) nrm = fmaf(v, v, nrm);
            v *= s;
            h[e] = __float2half_rn(v);
            l[e] = __float2half_rn(v - __half2float(h[e]));
        }
        *reinterpret_cast<uint4*>(ho + g * 8) = *reinterpret_cast<const uint4*>(h);
        *reinterpret_cast<uint4*>(lw + g * 8) = *reinterpret_cast<const uint4*>(l);
    }
    xx[(long long)b * npad + n] = nrm;
}

// row-major x (B,N,d) d <= W -> fp16 hi/lo rows [B][N][W] (W = 128 or 192, zero-padded) scaled by `scale`
__global__ void pack_rm_kernel(const float* __restrict__ x, long long rows, int d, int W, float scale, __half* __restrict__ hi,
                               __half* __restrict__ lo) {
    const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    for (int c = lane * 4; c < W; c += 128) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < d) v = *reinterpret_cast<const float4*>(x + row * d + c);
        const float a[4] = {v.x * scale, v.y * scale, v.z * scale, v.w * scale};
        __align__(8) __half h[4];
        __align__(8) __half l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            h[e] = __float2half_rn(a[e]);
            l[e] = __float2half_rn(a[e] - __half2float(h[e]));
        }
        *reinterpret_cast<uint2*>(hi + row * W + c) = *reinterpret_cast<const uint2*>(h);
        *reinterpret_cast<uint2*>(lo + row * W + c) = *reinterpret_cast<const uint2*>(l);
    }
}

template <int MODE, int OUT, int KB, int RBITS, int CB = 1>
static int launch_select(const CUtensorMap& qh, const CUtensorMap& ql, const CUtensorMap& xh, const CUtensorMap& xl,
                         const SelParams& p, int B, cudaStream_t st) {
    constexpr int STAGES = SelStages<KB, CB>::value;
    constexpr int NBINS = 1 << RBITS;
    constexpr size_t SEL = (OUT == OUT_TOP1) ? 1024 : (size_t)2 * NBINS * 128 * CB;
    constexpr size_t smem = (size_t)(STAGES + 1) * 2 * KB * BOX_BYTES + SEL + 1024 + 256;
    static_assert(smem <= 227 * 1024, "shared memory budget");
    auto kern = select_tc_kernel<MODE, OUT, KB, RBITS, CB>;
    SED_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((p.Nq + ST_M - 1) / ST_M, B);
    kern<<<grid, ST_THREADS, smem, st>>>(qh, ql, xh, xl, p);
    SED_CHECK_LAUNCH();
    return SED_OK;
}

// kNN over channel-major features: metric L2 (pn = 0) or point x normal (pn = 1).  k <= 64, C <= 64, N < 65536.
int knn_tc(const float* x, long long bstride, int B, int C, int N, int k, int pn, float W, void* idx, int idx64,
           cudaStream_t st, int sorted) {
    if (k > ST_KMAX || C > 64 || N >= 65536 || N < k || (pn && C != 6)) return SED_ERR_UNSUPPORTED;
    const int npad = (N + 127) / 128 * 128;
    const size_t rows = (size_t)B * N;
    ensure_pool_config();
    char* buf = nullptr;
    const size_t bytes_h = rows * 64 * sizeof(__half);
    SED_CUDA(cudaMallocAsync((void**)&buf, 2 * bytes_h + (size_t)B * npad * 4 + align_up((size_t)B * 4), st));   // + per-cloud max|x|
    __half* hi = (__half*)buf;
    __half* lo = (__half*)(buf + bytes_h);
    float* xx = (float*)(buf + 2 * bytes_h);
    float* mx = xx + (size_t)B * npad;
    int rc = SED_OK;
    if (cudaMemsetAsync(mx, 0, B * sizeof(float), st) != cudaSuccess) rc = SED_ERR_CUDA_BASE - 1;
    if (rc == SED_OK) {
        maxabs_kernel<<<dim3(64, B), 256, 0, st>>>(x, bstride, (long long)C * N, mx);
        pack_cm_kernel<<<dim3((npad + 127) / 128, B), 128, 0, st>>>(x, bstride, C, N, npad, pn, mx, hi, lo, xx);
        g_sed_launches += 2;
        CUtensorMap mh, ml;
        rc = make_map_f16(&mh, hi, B, N, 64);
        if (rc == SED_OK) rc = make_map_f16(&ml, lo, B, N, 64);
        static const int ss_win = env_int("SEDNET_B200_SS_WIN", SS_WIN), ss_soft = env_int("SEDNET_B200_SS_SOFT", SS_SOFT);
        SelParams p{xx, xx, mx, mx, 1.0f, N, N, npad, k, nullptr, W, idx, idx64, nullptr, sorted ? 0 : 1,
                    min(max(ss_win, 0), SS_CAP - 32 - k), min(max(ss_soft, k + 8), SS_CAP - 33)};
        // SEDNET_B200_KNN=radix selects the multi-pass radix kernel (A/B comparisons); default: single-pass streaming
        static const bool radix = [] { const char* e = getenv("SEDNET_B200_KNN"); return e && !strcmp(e, "radix"); }();
        if (rc == SED_OK && (radix || (pn && !(W >= 0.f)))) {
            rc = pn ? launch_select<SEL_PN, OUT_IDX, 1, 8>(mh, ml, mh, ml, p, B, st)
                    : launch_select<SEL_L2, OUT_IDX, 1, 8>(mh, ml, mh, ml, p, B, st);
        } else if (rc == SED_OK) {
            CUtensorMap xh, xl;   // candidate tiles of the streaming kernel are 64 rows
            rc = make_map_f16(&xh, hi, B, N, 64, SS_NC);
            if (rc == SED_OK) rc = make_map_f16(&xl, lo, B, N, 64, SS_NC);
            if (rc == SED_OK)
                rc = pn ? launch_stream<SEL_PN>(xh, xl, hi, lo, p, B, st) : launch_stream<SEL_L2>(xh, xl, hi, lo, p, B, st);
        }
    }
    cudaFreeAsync(buf, st);
    return rc;
}

// Cosine scores between unit rows: Q (B,Nq,d), Cand (B,Nc,d) row-major, d <= 128 (multiple of 4).
//   kth_out != null : K-th largest score (-dist) of every row (Q == Cand)           -> compute_bandwidth
//   idx_out != null : index of the best candidate of every row (first on ties)       -> nms membership / labels
int cos_select_tc(const float* Q, const float* Cand, int B, int Nq, int Nc, const int* nc_ptr, int d, int K,
                  float* kth_out, void* idx_out, int idx64, cudaStream_t st) {
    if (d > 192 || (d & 3) || Nc >= 65536 || (kth_out && (K > Nc || K <= 0))) return SED_ERR_UNSUPPORTED;
    const int W = d <= 128 ? 128 : 192;   // operand row width: two or three 64-channel boxes
    const bool same = (Q == Cand && Nq == Nc);
    const size_t rq = (size_t)B * Nq, rc_ = (size_t)B * Nc;
    const size_t bytes_q = rq * W * sizeof(__half), bytes_c = rc_ * W * sizeof(__half);
    ensure_pool_config();
    char* buf = nullptr;
    SED_CUDA(cudaMallocAsync((void**)&buf, 2 * bytes_q + (same ? 0 : 2 * bytes_c), st));
    __half *qh = (__half*)buf, *ql = (__half*)(buf + bytes_q);
    __half *ch = same ? qh : (__half*)(buf + 2 * bytes_q), *cl = same ? ql : (__half*)(buf + 2 * bytes_q + bytes_c);
    const float scale = 8.0f;
    pack_rm_kernel<<<(unsigned)((rq + 7) / 8), 256, 0, st>>>(Q, (long long)rq, d, W, scale, qh, ql);
    ++g_sed_launches;
    if (!same) {
        pack_rm_kernel<<<(unsigned)((rc_ + 7) / 8), 256, 0, st>>>(Cand, (long long)rc_, d, W, scale, ch, cl);
        ++g_sed_launches;
    }
    CUtensorMap mqh, mql, mch, mcl;
    int rc = make_map_f16(&mqh, qh, B, Nq, W);
    if (rc == SED_OK) rc = make_map_f16(&mql, ql, B, Nq, W);
    if (rc == SED_OK) rc = make_map_f16(&mch, ch, B, Nc, W);
    if (rc == SED_OK) rc = make_map_f16(&mcl, cl, B, Nc, W);
    static const int ks_win = env_int("SEDNET_B200_KS_WIN", KS_WIN), ks_soft = env_int("SEDNET_B200_KS_SOFT", KS_SOFT);
    SelParams p{nullptr, nullptr, nullptr, nullptr, scale, Nq, Nc, 0, K, nc_ptr, 0.f, idx_out, idx64, kth_out, 0,
                min(max(ks_win, 0), max(KS_CAP - 32 - K - 8, 0)), ks_soft > 0 ? min(max(ks_soft, K + 16), KS_CAP - 33) : 0};
    static const bool radix = [] { const char* e = getenv("SEDNET_B200_KNN"); return e && !strcmp(e, "radix"); }();
    if (rc == SED_OK && W == 192) {   // 129..192 columns (the hpnet embedding): the multi-pass radix kernel, three boxes per row
        rc = !kth_out  ? launch_select<SEL_COS, OUT_TOP1, 3, 8>(mqh, mql, mch, mcl, p, B, st)
             : K <= 255 ? launch_select<SEL_COS, OUT_KTH, 3, 7>(mqh, mql, mch, mcl, p, B, st)
                        : launch_select<SEL_COS, OUT_KTH, 3, 6, 2>(mqh, mql, mch, mcl, p, B, st);   // 16-bit counters
    } else if (rc == SED_OK && kth_out && K <= KS_KMAX && !radix) {
        CUtensorMap xh64, xl64;   // 64-row candidate tiles
        rc = make_map_f16(&xh64, ch, B, Nc, 128, SS_NC);
        if (rc == SED_OK) rc = make_map_f16(&xl64, cl, B, Nc, 128, SS_NC);
        if (rc == SED_OK) rc = launch_kth_stream(xh64, xl64, qh, ql, p, B, st);
    } else if (rc == SED_OK) {
        rc = !kth_out  ? launch_select<SEL_COS, OUT_TOP1, 2, 8>(mqh, mql, mch, mcl, p, B, st)
             : K <= 255 ? launch_select<SEL_COS, OUT_KTH, 2, 7>(mqh, mql, mch, mcl, p, B, st)
                        : launch_select<SEL_COS, OUT_KTH, 2, 7, 2>(mqh, mql, mch, mcl, p, B, st);   // 16-bit counters
    }
    cudaFreeAsync(buf, st);
    return rc;
}

}  // namespace sed
