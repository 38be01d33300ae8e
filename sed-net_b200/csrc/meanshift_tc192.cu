// Mean-shift iteration on the tensor cores for embeddings of 129..192 columns (the 148-column hpnet embedding,
// reference src/smooth_normal_matrix.py:221-233 -> src/mean_shift.py:45-79).  Same scheme as meanshift_tc.cu --
// flash-style pass, FP16 hi/lo split operands, S = Qh Xh + Qh Xl + Ql Xh (FP32-faithful), P = K(S) in FP16 written over
// S in TMEM, O += Ph Xh -- re-laid-out for 192-wide rows:
//   TMEM   S0 [0,64)  S1 [64,128)   O [128,320)   Q_hi [320,416)   Q_lo [416,512)          (all 512 columns)
//   smem   4 stages x [X_hi: 3 boxes of 64 keys x 64 channels | X_lo: 3 boxes] = 4 x 48 KB
// O needs 192 accumulator columns and Q 2 x 96, which leaves 128 columns for S: the key tile is 64 wide (not 128) so that
// S stays double-buffered -- the S MMAs of tile j+1 overlap the exp work of tile j, exp group g serving the tiles of
// buffer g, exactly as in the 128-wide kernel.
#include "tc_common.cuh"

namespace sed {

constexpr int W_M = 128, W_NK = 64, W_D = 192, W_THREADS = 320, W_STAGES = 4;
constexpr uint32_t W_XBOX = W_NK * 128;               // one TMA box of the key tile: 64 keys x 64 channels fp16 = 8 KB
constexpr uint32_t W_PART = 3 * W_XBOX;               // one operand part (hi or lo) of a 64 x 192 tile: 24 KB
constexpr uint32_t W_STAGE = 2 * W_PART;
constexpr uint32_t W_COL_O = 128, W_COL_QH = 320, W_COL_QL = 416;
constexpr float kWScale = 8.0f;

struct WParams {
    const __half* q_hi; const __half* q_lo;   // (B,N,192) operand of this iteration
    const float* bw;                          // (B)
    float* out_f32;                           // (B,N,192) or null
    __half* q_next_hi; __half* q_next_lo;     // (B,N,192) operand of the next iteration
    int N, qt_per_cloud;
};

template <int KT>
__global__ void __launch_bounds__(W_THREADS, 1)
ms_shift_tc192_kernel(const __grid_constant__ CUtensorMap map_xh, const __grid_constant__ CUtensorMap map_xl, WParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t x_addr = smem_base;
    const uint32_t bar_base = x_addr + W_STAGES * W_STAGE;
    const uint32_t bar_q_full = bar_base;
    const uint32_t bar_x_full = bar_base + 8;                    // [W_STAGES]
    const uint32_t bar_x_empty = bar_x_full + 8 * W_STAGES;      // [W_STAGES]
    const uint32_t bar_s_full = bar_x_empty + 8 * W_STAGES;      // [2]
    const uint32_t bar_p_full = bar_s_full + 16;                 // [2]
    const uint32_t bar_o_full = bar_p_full + 16;
    const uint32_t tmem_slot = bar_o_full + 8;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int N = p.N;
    const int T = (N + W_NK - 1) / W_NK;
    const int b = blockIdx.x / p.qt_per_cloud, q0 = (blockIdx.x % p.qt_per_cloud) * W_M;

    if (threadIdx.x == 0) {
        mbar_init(bar_q_full, 256);
        for (int s = 0; s < W_STAGES; ++s) { mbar_init(bar_x_full + 8 * s, 1); mbar_init(bar_x_empty + 8 * s, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(bar_s_full + 8 * i, 1); mbar_init(bar_p_full + 8 * i, 128); }
        mbar_init(bar_o_full, 1);
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
                const int s = j % W_STAGES;
                if (j >= W_STAGES) mbar_wait(bar_x_empty + 8 * s, ((j / W_STAGES) - 1) & 1);
                const uint32_t dst = x_addr + s * W_STAGE, bar = bar_x_full + 8 * s;
                mbar_expect_tx(bar, W_STAGE);
#pragma unroll
                for (int box = 0; box < 3; ++box) {
                    tma_load_3d(dst + box * W_XBOX, &map_xh, bar, box * 64, j * W_NK, b);
                    tma_load_3d(dst + W_PART + box * W_XBOX, &map_xl, bar, box * 64, j * W_NK, b);
                }
            }
        }
    } else if (warp == 1) {
        // ============================================================ MMA issuer (one elected thread)
        if (elect_one()) {
            constexpr uint32_t IDESC_S = make_idesc_n(0, W_NK), IDESC_PV = make_idesc_n(1, W_D);
            // S(j) = Qh Xh + Qh Xl + Ql Xh over 192 channels = 12 K-steps of 16, into S buffer j & 1
            auto issue_s = [&](int j) {
                const uint32_t xs = x_addr + (j % W_STAGES) * W_STAGE;
                const uint32_t d = tmem + (uint32_t)(j & 1) * 64u;
                uint32_t acc = 0;
#pragma unroll
                for (int term = 0; term < 3; ++term) {
                    const uint32_t qa = tmem + (term == 2 ? W_COL_QL : W_COL_QH);
                    const uint32_t xb = xs + (term == 1 ? W_PART : 0);
#pragma unroll
                    for (int ks = 0; ks < W_D / 16; ++ks) {
                        const uint32_t off = (ks >> 2) * W_XBOX + (ks & 3) * 32;
                        umma_ts(d, qa + ks * 8, make_desc(xb + off, 16), IDESC_S, acc);
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
                    mbar_wait(bar_x_full + 8 * ((j + 1) % W_STAGES), ((j + 1) / W_STAGES) & 1);
                    tc_fence_after();
                    issue_s(j + 1);
                    tc_commit(bar_s_full + 8 * ((j + 1) & 1));
                }
                mbar_wait(bar_p_full + 8 * (j & 1), (j >> 1) & 1);
                tc_fence_after();
                // O += P(j) Xh(j): A = P in TMEM (8 columns per K-step of 16 keys), B = the X_hi tile MN-major, N = 192
                const uint32_t xs = x_addr + (j % W_STAGES) * W_STAGE;
                const uint32_t pa = tmem + (uint32_t)(j & 1) * 64u;
#pragma unroll
                for (int ks = 0; ks < W_NK / 16; ++ks)
                    umma_ts(tmem + W_COL_O, pa + ks * 8, make_desc(xs + ks * 2048, W_XBOX), IDESC_PV, (j > 0 || ks > 0) ? 1u : 0u);
                tc_commit(bar_x_empty + 8 * (j % W_STAGES));
            }
            tc_commit(bar_o_full);
        }
    } else {
        // ============================================================ exp warps + epilogue: two groups of 4 x 32 rows
        const int group = (warp - 2) >> 2;
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        const float bwv = p.bw[b];
        const float inv_b2 = 1.0f / (bwv * bwv);
        const float c1 = inv_b2 * 1.4426950408889634f / (kWScale * kWScale), c0 = -inv_b2 * 1.4426950408889634f;
        const float e1 = 1.5f * inv_b2 / (kWScale * kWScale), e0 = 0.75f - 1.5f * inv_b2;
        {   // this thread's query row -> TMEM: group 0 the hi part, group 1 the lo part (96 columns each)
            const int qrow = q0 + row;
            const __half* src = group == 1 ? p.q_lo : p.q_hi;
            const uint4* g4 = reinterpret_cast<const uint4*>(src + ((long long)b * N + min(qrow, N - 1)) * W_D);
            const uint32_t qb = tmem + lane_addr + (group == 1 ? W_COL_QL : W_COL_QH);
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                uint32_t w[16];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    uint4 v = make_uint4(0u, 0u, 0u, 0u);
                    if (qrow < N) v = __ldg(g4 + c * 4 + i);
                    w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
                }
                tmem_st16(qb + c * 16, w);
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(bar_q_full);
        }
        for (int j = group; j < T; j += 2) {
            const uint32_t sb = tmem + lane_addr + (uint32_t)(j & 1) * 64u;
            mbar_wait(bar_s_full + 8 * (j & 1), (j >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t v[32];
                tmem_ld32(sb + c * 32, v);
                tmem_ld_wait();
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float s0 = __uint_as_float(v[2 * i]), s1 = __uint_as_float(v[2 * i + 1]);
                    float p0, p1;
                    if (KT == 0) {
                        p0 = ex2_approx(fmaf(s0, c1, c0));
                        p1 = ex2_approx(fmaf(s1, c1, c0));
                    } else {
                        p0 = fmaxf(fmaf(s0, e1, e0), 0.f);
                        p1 = fmaxf(fmaf(s1, e1, e0), 0.f);
                    }
                    const __half2 h = __floats2half2_rn(p0, p1);
                    pk[i] = *reinterpret_cast<const uint32_t*>(&h);
                }
                tmem_st16(sb + c * 16, pk);   // P (fp16) overwrites the S columns this thread has already consumed
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(bar_p_full + 8 * (j & 1));
        }
        // ---- epilogue: new = O / ||O||.  Thread (row, group) owns three of the row's six 32-channel chunks.
        mbar_wait(bar_o_full, 0);
        tc_fence_after();
        const uint32_t ob = tmem + lane_addr + W_COL_O;
        float* ssx = reinterpret_cast<float*>(smem_raw + (x_addr - smem_u32(smem_raw)));   // the X stages are dead by now
        float ss = 0.f;
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
            uint32_t v[32];
            tmem_ld32(ob + (group * 3 + cc) * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) ss = fmaf(__uint_as_float(v[i]), __uint_as_float(v[i]), ss);
        }
        ssx[group * 128 + row] = ss;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        // a row whose every weight underflows FP16 keeps its position instead of becoming NaN (see meanshift_tc.cu)
        const float sst = ssx[row] + ssx[128 + row];
        const bool dead = !(sst > 0.f);
        const bool any_dead = __any_sync(0xffffffffu, dead);
        const float rn = dead ? 1.0f / kWScale : 1.0f / sqrtf(sst);
        const int q = q0 + row;
        const long long rowoff = ((long long)b * N + q) * W_D;
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
            const int c = group * 3 + cc;
            uint32_t v[32];
            tmem_ld32(ob + c * 32, v);
            tmem_ld_wait();
            if (any_dead) {
                uint32_t qh16[16], ql16[16];
                tmem_ld16(tmem + lane_addr + W_COL_QH + (uint32_t)c * 16u, qh16);
                tmem_ld16(tmem + lane_addr + W_COL_QL + (uint32_t)c * 16u, ql16);
                tmem_ld_wait();
                if (dead) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&qh16[i]));
                        const float2 l = __half22float2(*reinterpret_cast<const __half2*>(&ql16[i]));
                        v[2 * i] = __float_as_uint(h.x + l.x);
                        v[2 * i + 1] = __float_as_uint(h.y + l.y);
                    }
                }
            }
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
                    const float a0 = z[2 * i] * kWScale, a1 = z[2 * i + 1] * kWScale;
                    const __half2 h = __floats2half2_rn(a0, a1);
                    const float2 hf = __half22float2(h);
                    const __half2 l = __floats2half2_rn(a0 - hf.x, a1 - hf.y);
                    hi[i] = *reinterpret_cast<const uint32_t*>(&h);
                    lo[i] = *reinterpret_cast<const uint32_t*>(&l);
                }
                uint4* dh = reinterpret_cast<uint4*>(p.q_next_hi + rowoff + c * 32);
                uint4* dl = reinterpret_cast<uint4*>(p.q_next_lo + rowoff + c * 32);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    dh[i] = make_uint4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
                    dl[i] = make_uint4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
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

// x (rows, d) f32, d <= 192 a multiple of 4 -> (rows, 192) hi = fp16(8x), lo = fp16(8x - hi), columns >= d zero
__global__ void split192_kernel(const float* __restrict__ x, long long n, int d, __half* __restrict__ hi, __half* __restrict__ lo) {
    const long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 4;
    if (i >= n) return;
    const long long row = i / W_D;
    const int col = (int)(i - row * W_D);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col < d) v = *reinterpret_cast<const float4*>(x + row * d + col);
    const float a[4] = {v.x * kWScale, v.y * kWScale, v.z * kWScale, v.w * kWScale};
    __half h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        h[e] = __float2half_rn(a[e]);
        l[e] = __float2half_rn(a[e] - __half2float(h[e]));
    }
    *reinterpret_cast<uint2*>(hi + i) = *reinterpret_cast<const uint2*>(h);
    *reinterpret_cast<uint2*>(lo + i) = *reinterpret_cast<const uint2*>(l);
}

// Embeddings of 129..192 columns; SED_ERR_UNSUPPORTED outside that range.
int ms_shift_tc192(const float* X, const float* bw, int B, int N, int d, int iterations, int kernel_type, float* out,
                   cudaStream_t st) {
    if (d <= 128 || d > W_D || (d & 3)) return SED_ERR_UNSUPPORTED;
    const size_t elems = (size_t)B * N * W_D;
    ensure_pool_config();
    __half* buf = nullptr;
    SED_CUDA(cudaMallocAsync((void**)&buf, elems * sizeof(__half) * 6 + elems * sizeof(float), st));
    __half *xh = buf, *xl = buf + elems, *qh[2] = {buf + 2 * elems, buf + 4 * elems}, *ql[2] = {buf + 3 * elems, buf + 5 * elems};
    float* out192 = reinterpret_cast<float*>(buf + 6 * elems);
    split192_kernel<<<(unsigned)((elems / 4 + 255) / 256), 256, 0, st>>>(X, (long long)elems, d, xh, xl);
    ++g_sed_launches;
    CUtensorMap mxh, mxl;
    int rc = make_map_f16(&mxh, xh, B, N, W_D, W_NK);
    if (rc == SED_OK) rc = make_map_f16(&mxl, xl, B, N, W_D, W_NK);
    constexpr size_t smem = (size_t)W_STAGES * W_STAGE + 1024 + 256;
    const int qtc = (N + W_M - 1) / W_M;
    if (rc == SED_OK) {
        if (cudaFuncSetAttribute(ms_shift_tc192_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
            cudaFuncSetAttribute(ms_shift_tc192_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            rc = SED_ERR_CUDA_BASE - 1;
    }
    for (int it = 0; it < iterations && rc == SED_OK; ++it) {
        const __half* cqh = it == 0 ? xh : qh[(it - 1) & 1];
        const __half* cql = it == 0 ? xl : ql[(it - 1) & 1];
        WParams p{cqh, cql, bw, it == iterations - 1 ? out192 : nullptr, qh[it & 1], ql[it & 1], N, qtc};
        if (kernel_type == 0) ms_shift_tc192_kernel<0><<<B * qtc, W_THREADS, smem, st>>>(mxh, mxl, p);
        else ms_shift_tc192_kernel<1><<<B * qtc, W_THREADS, smem, st>>>(mxh, mxl, p);
        ++g_sed_launches;
        if (cudaGetLastError() != cudaSuccess) rc = SED_ERR_CUDA_BASE - 1;
    }
    if (rc == SED_OK &&
        cudaMemcpy2DAsync(out, (size_t)d * sizeof(float), out192, (size_t)W_D * sizeof(float), (size_t)d * sizeof(float),
                          (size_t)B * N, cudaMemcpyDeviceToDevice, st) != cudaSuccess)
        rc = SED_ERR_CUDA_BASE - 1;
    cudaFreeAsync(buf, st);
    return rc;
}

}  // namespace sed
