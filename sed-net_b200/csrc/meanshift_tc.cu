// Mean-shift iteration on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
// One iteration (reference src/mean_shift.py:57-77) is   O = K(Q X^T) X,  new = O / ||O||   with
// K(s) = exp((s - 1) / b^2)  (= exp(-(2 - 2s)/b^2/2)); the 1/sum K factor of the reference cancels in the
// normalisation.  It is attention without a running max (the exponent is <= 0), so the kernel is a flash-style
// pass: a CTA owns 128 query rows, streams 128-key tiles of X through shared memory by TMA, and per tile issues
//     S  = Q X^T          tcgen05.mma  SS  (A = Q tile, B = X tile, both K-major, 128B swizzle)  -> TMEM
//     P  = K(S)           4 warps: tcgen05.ld -> ex2 -> f16x2 -> tcgen05.st (P overwrites S in TMEM)
//     O += P X            tcgen05.mma  TS  (A = P from TMEM, B = the SAME X tile read MN-major)    -> TMEM
// with S double-buffered in TMEM so that the tensor pipe computes S(j+1) while the exp warps work on S(j).
//
// Precision.  Operands are FP16 with FP32 accumulation.  FP16 has the 11-bit significand of TF32 at twice the
// rate; to stay FP32-faithful where it matters (the exponent s/b^2 amplifies errors of s by 1/b^2 ~ 100) each
// operand is split x*8 = hi + lo (hi = fp16(8x), lo = fp16(8x - hi): 22 significant bits, lo may be subnormal,
// which costs nothing in absolute error) and
//     prec_mode 1:  S = Qh Xh + Qh Xl + Ql Xh   (3 MMAs),   O = Ph Xh + Ph Xl   (2 MMAs)
//     prec_mode 2:  S = Qh Xh,                               O = Ph Xh           (fast)
// The scale 8 keeps hi/lo away from the bottom of the FP16 range; 64 = 8*8 is folded into the exp2 argument and
// the factor 8 on O vanishes in the normalisation.
#include <cuda.h>
#include <cuda_fp16.h>

#include "internal.h"

namespace sed {

constexpr int TC_M = 128;        // query rows per CTA  (UMMA M)
constexpr int TC_NK = 128;       // keys per tile        (UMMA N of S, K of PV)
constexpr int TC_D = 128;        // channels             (K of S, N of PV)
constexpr int TC_THREADS = 192;  // warp 0 TMA, warp 1 MMA, warps 2-5 exp/epilogue
constexpr int BOX_BYTES = 128 * 128;            // one TMA box: 128 rows x 64 fp16 (128 B, one swizzle atom wide)
constexpr int TILE_BYTES = 2 * BOX_BYTES;       // 128 rows x 128 fp16
constexpr float kOperandScale = 8.0f;

// ---------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
          "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Shared-memory matrix descriptor (tcgen05): 128B swizzle, SBO = 1024 B (8 rows x 128 B), version 1.
//   K-major  operand: LBO field unused by the swizzled layouts (set to 1);
//   MN-major operand: LBO = byte distance between the two 64-element halves of the MN extent.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);            // start address  bits [0,14)
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;   // leading byte offset bits [16,30)
    d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;        // stride byte offset bits [32,46)
    d |= (uint64_t)1 << 46;                             // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                             // SWIZZLE_128B
    return d;
}
// Instruction descriptor kind::f16: D = F32, A = B = F16, M = 128, N = 128; b_mn_major selects the B layout.
__host__ __device__ constexpr uint32_t make_idesc(int b_mn_major) {
    return (1u << 4) | (0u << 7) | (0u << 10) | (0u << 15) | ((uint32_t)b_mn_major << 16) | ((128u >> 3) << 17) |
           ((128u >> 4) << 24);
}

struct TcParams {
    const float* bw;        // (B)
    float* out_f32;         // (B,N,128) or null
    __half* q_next_hi;      // (B,N,128) operand of the next iteration
    __half* q_next_lo;      // or null when the mode has no lo part
    int N, kernel_type;
};

// NS: MMAs per S tile (1 or 3); NV: MMAs per PV tile (1 or 2)
template <int NS, int NV>
__global__ void __launch_bounds__(TC_THREADS, 1)
ms_shift_tc_kernel(const __grid_constant__ CUtensorMap map_qh, const __grid_constant__ CUtensorMap map_ql,
                   const __grid_constant__ CUtensorMap map_xh, const __grid_constant__ CUtensorMap map_xl, TcParams p) {
    constexpr bool HAS_LO = (NS > 1) || (NV > 1);
    constexpr int STAGES = HAS_LO ? 2 : 4;
    constexpr int PARTS = HAS_LO ? 2 : 1;
    constexpr uint32_t STAGE_BYTES = PARTS * TILE_BYTES;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // 128B swizzle needs 1024-B alignment
    const uint32_t q_addr = smem_base;                                  // [hi | lo] x 32 KB
    const uint32_t x_addr = q_addr + STAGE_BYTES;                       // STAGES x [hi | lo]
    const uint32_t bar_base = x_addr + STAGES * STAGE_BYTES;
    // barriers (8 B each)
    const uint32_t bar_q_full = bar_base;
    const uint32_t bar_x_full = bar_base + 8;                  // [STAGES]
    const uint32_t bar_x_empty = bar_x_full + 8 * STAGES;      // [STAGES]
    const uint32_t bar_s_full = bar_x_empty + 8 * STAGES;      // [2]
    const uint32_t bar_p_full = bar_s_full + 16;               // [2]
    const uint32_t bar_o_full = bar_p_full + 16;
    const uint32_t tmem_slot = bar_o_full + 8;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y, q0 = blockIdx.x * TC_M;
    const int N = p.N;
    const int T = (N + TC_NK - 1) / TC_NK;

    if (threadIdx.x == 0) {
        mbar_init(bar_q_full, 1);
        for (int s = 0; s < STAGES; ++s) { mbar_init(bar_x_full + 8 * s, 1); mbar_init(bar_x_empty + 8 * s, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(bar_s_full + 8 * i, 1); mbar_init(bar_p_full + 8 * i, 128); }
        mbar_init(bar_o_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {  // TMEM: S0 [0,128) S1 [128,256) O [256,384): allocate 512 columns (power of two)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tmem_slot) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot_ptr;

    if (warp == 0) {
        // ============================================================ TMA producer
        if (lane == 0) {
            mbar_expect_tx(bar_q_full, STAGE_BYTES);
            tma_load_3d(q_addr, &map_qh, bar_q_full, 0, q0, b);
            tma_load_3d(q_addr + BOX_BYTES, &map_qh, bar_q_full, 64, q0, b);
            if (HAS_LO) {
                tma_load_3d(q_addr + TILE_BYTES, &map_ql, bar_q_full, 0, q0, b);
                tma_load_3d(q_addr + TILE_BYTES + BOX_BYTES, &map_ql, bar_q_full, 64, q0, b);
            }
            for (int j = 0; j < T; ++j) {
                const int s = j % STAGES;
                if (j >= STAGES) mbar_wait(bar_x_empty + 8 * s, ((j / STAGES) - 1) & 1);
                const uint32_t dst = x_addr + s * STAGE_BYTES, bar = bar_x_full + 8 * s;
                mbar_expect_tx(bar, STAGE_BYTES);
                tma_load_3d(dst, &map_xh, bar, 0, j * TC_NK, b);
                tma_load_3d(dst + BOX_BYTES, &map_xh, bar, 64, j * TC_NK, b);
                if (HAS_LO) {
                    tma_load_3d(dst + TILE_BYTES, &map_xl, bar, 0, j * TC_NK, b);
                    tma_load_3d(dst + TILE_BYTES + BOX_BYTES, &map_xl, bar, 64, j * TC_NK, b);
                }
            }
        }
    } else if (warp == 1) {
        // ============================================================ MMA issuer (one thread)
        if (lane == 0) {
            constexpr uint32_t IDESC_S = make_idesc(0), IDESC_PV = make_idesc(1);
            const uint32_t tmem_o = tmem + 256;
            // S(j) = sum over terms (a_part, b_part) of Q[a_part] X[b_part]^T, K = 128 channels = 8 steps of 16
            auto issue_s = [&](int j) {
                const uint32_t xs = x_addr + (j % STAGES) * STAGE_BYTES;
                const uint32_t d = tmem + (uint32_t)(j & 1) * 128u;
                uint32_t acc = 0;
#pragma unroll
                for (int term = 0; term < NS; ++term) {
                    const uint32_t qa = q_addr + ((term == 2) ? TILE_BYTES : 0);   // Qh, Qh, Ql
                    const uint32_t xb = xs + ((term == 1) ? TILE_BYTES : 0);       // Xh, Xl, Xh
#pragma unroll
                    for (int ks = 0; ks < TC_D / 16; ++ks) {
                        const uint32_t off = (ks >> 2) * BOX_BYTES + (ks & 3) * 32;  // 4 K-steps per 128-B swizzle row
                        umma_ss(d, make_desc(qa + off, 16), make_desc(xb + off, 16), IDESC_S, acc);
                        acc = 1;
                    }
                }
            };
            mbar_wait(bar_q_full, 0);
            mbar_wait(bar_x_full, 0);
            tc_fence_after();
            issue_s(0);
            tc_commit(bar_s_full);
            for (int j = 0; j < T; ++j) {
                if (j + 1 < T) {
                    mbar_wait(bar_x_full + 8 * ((j + 1) % STAGES), ((j + 1) / STAGES) & 1);
                    tc_fence_after();
                    issue_s(j + 1);
                    tc_commit(bar_s_full + 8 * ((j + 1) & 1));
                }
                mbar_wait(bar_p_full + 8 * (j & 1), (j >> 1) & 1);
                tc_fence_after();
                // O += P(j) X(j): A = P in TMEM (fp16 pairs, 8 columns per K-step of 16 keys), B = X tile MN-major
                const uint32_t xs = x_addr + (j % STAGES) * STAGE_BYTES;
                const uint32_t pa = tmem + (uint32_t)(j & 1) * 128u;
#pragma unroll
                for (int term = 0; term < NV; ++term) {
                    const uint32_t xb = xs + term * TILE_BYTES;
#pragma unroll
                    for (int ks = 0; ks < TC_NK / 16; ++ks) {
                        umma_ts(tmem_o, pa + ks * 8, make_desc(xb + ks * 2048, BOX_BYTES), IDESC_PV,
                                (j > 0 || term > 0 || ks > 0) ? 1u : 0u);
                    }
                }
                tc_commit(bar_x_empty + 8 * (j % STAGES));
            }
            tc_commit(bar_o_full);
        }
    } else {
        // ============================================================ exp warps (4 x 32 rows) + epilogue
        const int quarter = warp & 3;                       // TMEM lane quarter this warp may access
        const int row = quarter * 32 + lane;                // query row within the tile
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        const float bwv = p.bw[b];
        const float inv_b2 = 1.0f / (bwv * bwv);
        // exp((s - 1)/b^2) = 2^(S_acc * c1 + c0),  S_acc = 64 s
        const float c1 = inv_b2 * 1.4426950408889634f / (kOperandScale * kOperandScale);
        const float c0 = -inv_b2 * 1.4426950408889634f;
        // epanechnikov: 0.75 * (1 - (2 - 2s)/b^2) = S_acc * e1 + e0
        const float e1 = 1.5f * inv_b2 / (kOperandScale * kOperandScale);
        const float e0 = 0.75f - 1.5f * inv_b2;
        for (int j = 0; j < T; ++j) {
            const uint32_t sb = tmem + lane_addr + (uint32_t)(j & 1) * 128u;
            mbar_wait(bar_s_full + 8 * (j & 1), (j >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t v[32];
                tmem_ld32(sb + c * 32, v);
                tmem_ld_wait();
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float p0, p1;
                    if (p.kernel_type == 0) {
                        p0 = ex2_approx(fmaf(__uint_as_float(v[2 * i]), c1, c0));
                        p1 = ex2_approx(fmaf(__uint_as_float(v[2 * i + 1]), c1, c0));
                    } else {
                        p0 = fmaxf(fmaf(__uint_as_float(v[2 * i]), e1, e0), 0.f);
                        p1 = fmaxf(fmaf(__uint_as_float(v[2 * i + 1]), e1, e0), 0.f);
                    }
                    const __half2 h = __floats2half2_rn(p0, p1);   // low half = even key, high half = odd key
                    pk[i] = *reinterpret_cast<const uint32_t*>(&h);
                }
                tmem_st16(sb + c * 16, pk);   // P (fp16) overwrites the S columns already consumed
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(bar_p_full + 8 * (j & 1));
        }
        // ---- epilogue: new = O / ||O||  (thread = one full row of 128 channels)
        mbar_wait(bar_o_full, 0);
        tc_fence_after();
        const uint32_t ob = tmem + lane_addr + 256u;
        float ss = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            uint32_t v[32];
            tmem_ld32(ob + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) { const float o = __uint_as_float(v[i]); ss = fmaf(o, o, ss); }
        }
        const float rn = 1.0f / sqrtf(ss);
        const int q = q0 + row;
        const long long rowoff = ((long long)b * N + q) * TC_D;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            uint32_t v[32];
            tmem_ld32(ob + c * 32, v);
            tmem_ld_wait();
            if (q < N) {
                float z[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) z[i] = __uint_as_float(v[i]) * rn;
                if (p.out_f32) {
                    float4* dst = reinterpret_cast<float4*>(p.out_f32 + rowoff + c * 32);
#pragma unroll
                    for (int i = 0; i < 8; ++i) dst[i] = make_float4(z[4 * i], z[4 * i + 1], z[4 * i + 2], z[4 * i + 3]);
                }
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float a0 = z[2 * i] * kOperandScale, a1 = z[2 * i + 1] * kOperandScale;
                    const __half2 h = __floats2half2_rn(a0, a1);
                    const float2 hf = __half22float2(h);
                    const __half2 l = __floats2half2_rn(a0 - hf.x, a1 - hf.y);
                    hi[i] = *reinterpret_cast<const uint32_t*>(&h);
                    lo[i] = *reinterpret_cast<const uint32_t*>(&l);
                }
                uint4* dh = reinterpret_cast<uint4*>(p.q_next_hi + rowoff + c * 32);
#pragma unroll
                for (int i = 0; i < 4; ++i) dh[i] = make_uint4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
                if (HAS_LO && p.q_next_lo) {
                    uint4* dl = reinterpret_cast<uint4*>(p.q_next_lo + rowoff + c * 32);
#pragma unroll
                    for (int i = 0; i < 4; ++i) dl[i] = make_uint4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
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

// x (rows, 128) f32 -> hi = fp16(8x), lo = fp16(8x - hi)
__global__ void split_f16_kernel(const float* __restrict__ x, long long n, __half* __restrict__ hi, __half* __restrict__ lo) {
    const long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 4;
    if (i >= n) return;
    const float4 v = *reinterpret_cast<const float4*>(x + i);
    const float a[4] = {v.x * kOperandScale, v.y * kOperandScale, v.z * kOperandScale, v.w * kOperandScale};
    __half h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        h[e] = __float2half_rn(a[e]);
        l[e] = __float2half_rn(a[e] - __half2float(h[e]));
    }
    *reinterpret_cast<uint2*>(hi + i) = *reinterpret_cast<const uint2*>(h);
    if (lo) *reinterpret_cast<uint2*>(lo + i) = *reinterpret_cast<const uint2*>(l);
}

// ---------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
    }
    return fn;
}

// (B, N, 128) fp16 row-major, box 128 rows x 64 channels, 128B swizzle, out-of-range rows read as zero
static int make_map(CUtensorMap* m, const __half* base, int B, int N) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return SED_ERR_UNSUPPORTED;
    const cuuint64_t dims[3] = {(cuuint64_t)TC_D, (cuuint64_t)N, (cuuint64_t)B};
    const cuuint64_t strides[2] = {(cuuint64_t)TC_D * 2, (cuuint64_t)N * TC_D * 2};
    const cuuint32_t box[3] = {64, 128, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, (void*)base, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? SED_OK : SED_ERR_CUDA_BASE - 1;
}

template <int NS, int NV>
static int launch_tc(const CUtensorMap& qh, const CUtensorMap& ql, const CUtensorMap& xh, const CUtensorMap& xl,
                     const TcParams& p, int B, cudaStream_t st) {
    constexpr bool HAS_LO = (NS > 1) || (NV > 1);
    constexpr int STAGES = HAS_LO ? 2 : 4;
    constexpr size_t smem = (size_t)(STAGES + 1) * (HAS_LO ? 2 : 1) * TILE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
    auto kern = ms_shift_tc_kernel<NS, NV>;
    SED_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((p.N + TC_M - 1) / TC_M, B);
    kern<<<grid, TC_THREADS, smem, st>>>(qh, ql, xh, xl, p);
    SED_CHECK_LAUNCH();
    return SED_OK;
}

int ms_shift_tc(const float* X, const float* bw, int B, int N, int d, int iterations, int kernel_type, int prec_mode,
                float* out, float* tmp, cudaStream_t st) {
    (void)tmp;
    if (d != TC_D) return SED_ERR_UNSUPPORTED;
    const bool has_lo = (prec_mode == 1);
    const size_t elems = (size_t)B * N * TC_D;
    // fp16 operands: X (hi, lo) and two ping-pong Q buffers (hi, lo)
    __half* buf = nullptr;
    SED_CUDA(cudaMallocAsync((void**)&buf, elems * sizeof(__half) * 6, st));
    __half *xh = buf, *xl = buf + elems, *qh[2] = {buf + 2 * elems, buf + 4 * elems},
           *ql[2] = {buf + 3 * elems, buf + 5 * elems};
    split_f16_kernel<<<(unsigned)((elems / 4 + 255) / 256), 256, 0, st>>>(X, (long long)elems, xh, xl);
    ++g_sed_launches;
    CUtensorMap mxh, mxl, mqh[2], mql[2];
    int rc = make_map(&mxh, xh, B, N);
    if (rc == SED_OK) rc = make_map(&mxl, xl, B, N);
    for (int i = 0; i < 2 && rc == SED_OK; ++i) {
        rc = make_map(&mqh[i], qh[i], B, N);
        if (rc == SED_OK) rc = make_map(&mql[i], ql[i], B, N);
    }
    for (int it = 0; it < iterations && rc == SED_OK; ++it) {
        // iteration 0 reads Q = X; iteration it > 0 reads ping-pong buffer (it-1)&1 and writes buffer it&1
        const CUtensorMap& cqh = it == 0 ? mxh : mqh[(it - 1) & 1];
        const CUtensorMap& cql = it == 0 ? mxl : mql[(it - 1) & 1];
        TcParams p{bw, it == iterations - 1 ? out : nullptr, qh[it & 1], has_lo ? ql[it & 1] : nullptr, N, kernel_type};
        rc = has_lo ? launch_tc<3, 2>(cqh, cql, mxh, mxl, p, B, st) : launch_tc<1, 1>(cqh, cql, mxh, mxl, p, B, st);
    }
    cudaFreeAsync(buf, st);
    return rc;
}

}  // namespace sed
