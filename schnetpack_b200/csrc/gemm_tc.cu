// Dense layer on the 5th-generation tensor cores (tcgen05, TMEM accumulator) with 3xTF32 error compensation:
//
//     Y = act( (A .* act'(a_pre)) * W^T + bias ) + addend            A [M,K] fp32, W [N,K] fp32 (K contiguous)
//
// The reference computes these layers with true-fp32 cuBLAS SGEMM (torch matmul precision "highest",
// /root/reference/src/schnetpack/cli.py:95-97, nn/base.py:52-55); single-pass TF32 (10-bit mantissa) would miss the
// 1e-5 parity bar by two orders of magnitude, so every operand is split  x = hi + lo  (hi = round-to-nearest TF32,
// lo = x - hi, exact in fp32) and three tensor-core products  Ah*Bh + Al*Bh + Ah*Bl  are accumulated in fp32 in TMEM.
// The dropped Al*Bl term and the TF32 rounding of lo are O(2^-22) relative.
//
// Structure (one 128-thread CTA per 128 x BN output tile, cta_group::1, UMMA M=128, N=BN<=128, K=8 per instruction):
//   * all four warps stage the next K-tile (32 floats) of A and W: coalesced 128-bit global loads, optional backward
//     prologue (x act'(a_pre)), hi/lo split in registers, st.shared into the canonical K-major no-swizzle core-matrix
//     layout (8 rows x 16 B cores; LBO = plane stride between 16 B K-chunks, SBO = 128 B between 8-row groups);
//   * fence.proxy.async + barrier, then ONE thread issues 12 tcgen05.mma.kind::tf32 per K-tile (4 k-steps x 3 products)
//     and tcgen05.commit's them to the stage's mbarrier; two smem stages let the loads of tile t+1 overlap the MMAs of t;
//   * epilogue: each warp reads its 32 TMEM lanes (= 32 output rows) with tcgen05.ld.32x32b.x32, applies
//     bias / activation / addend, optionally saves the pre-activation, and stores 128-bit vectors.
#include "common.cuh"

namespace {

constexpr int TM = 128;       // rows per CTA tile (UMMA M)
constexpr int TN = 128;       // max columns per CTA tile (UMMA N)
constexpr int TK = 32;        // floats per K-tile (4 UMMA k-steps of 8)
constexpr int NSTAGE = 2;
constexpr int GROUPS = TM / 8;                 // 8-row groups
constexpr int PLANE = GROUPS * 128 + 16;       // bytes between consecutive 16 B K-chunks (LBO), padded vs bank conflicts
constexpr int OPER_BYTES = 8 * PLANE;          // one operand tile (8 K-chunks of 16 B per row)
constexpr int STAGE_BYTES = 4 * OPER_BYTES;    // A_hi, A_lo, B_hi, B_lo
constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    // K-major, SWIZZLE_NONE: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout 0 [61,64)
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(PLANE >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) |
           (1ull << 46);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}

__device__ __forceinline__ float tf32_rn(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// 32 consecutive accumulator columns of this thread's TMEM lane -> registers
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct TcArgs {
    const float* A;
    const float* a_pre;
    const float* Wh;   // [N, K] tf32-rounded weights
    const float* Wl;   // [N, K] residual
    const float* bias;
    const float* addend;
    float* Y;
    float* y_pre;
    int64_t M, lda, ld_add, ldy;
    int K, N, a_act, act;
};

// byte offset of (row r, 16 B chunk c) inside an operand tile
__device__ __forceinline__ int tile_off(int r, int c) { return c * PLANE + (r >> 3) * 128 + (r & 7) * 16; }

__global__ void __launch_bounds__(128, 1) k_dense_tc(TcArgs g) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar_mma[NSTAGE];
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t m0 = (int64_t)blockIdx.x * TM;
    const int n0 = blockIdx.y * TN;
    const int bn_real = min(TN, g.N - n0);            // valid columns of this tile
    const int BN = (bn_real + 15) & ~15;              // UMMA N (multiple of 16)

    if (tid == 0) {
        mbar_init(&bar_mma[0], 1);
        mbar_init(&bar_mma[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(2 * TN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = s_tmem;

    // instruction descriptor: D=F32 (1<<4), A=B=TF32 (2<<7, 2<<10), K-major both, N>>3 at [17,23), M>>4 at [24,29)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

    const int nk = (g.K + TK - 1) / TK;
    // coalesced staging map: lane -> (row sub-index = lane/8, 16 B chunk = lane%8); 8 passes cover 32 rows per warp
    const int chunk = lane & 7, rsub = lane >> 3;
    const bool a_vec = ((g.lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.A) & 15) == 0) &&
                       (!g.a_pre || (reinterpret_cast<uintptr_t>(g.a_pre) & 15) == 0);
    const bool w_vec = ((g.K & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.Wh) & 15) == 0) &&
                       ((reinterpret_cast<uintptr_t>(g.Wl) & 15) == 0);

    // fp32 running sum of the per-k-tile main accumulators.  The tensor core adds into its TMEM accumulator with
    // truncation (biased toward zero); keeping every hi*hi partial sum to ONE k-tile (3 truncating adds) and summing the
    // tiles here with IEEE round-to-nearest keeps the layer at fp32-grade error.
    float accr[TN];
#pragma unroll
    for (int i = 0; i < TN; ++i) accr[i] = 0.f;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    auto drain = [&](int tile) {
        mbar_wait(&bar_mma[tile & 1], (tile >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int c0 = 0; c0 < TN; c0 += 32) {
            if (c0 < BN) {
                uint32_t r[32];
                tmem_ld32(lane_addr + (uint32_t)c0, r);
#pragma unroll
                for (int j = 0; j < 32; ++j) accr[c0 + j] += __uint_as_float(r[j]);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    };

    for (int kt = 0; kt < nk; ++kt) {
        const int s = kt & 1;
        uint8_t* st = smem + s * STAGE_BYTES;
        if (kt >= NSTAGE) mbar_wait(&bar_mma[s], ((kt - NSTAGE) >> 1) & 1);   // MMAs that read this stage retired
        const int k = kt * TK + chunk * 4;
        // ---- A tile: rows warp*32 + p*4 + rsub ------------------------------------------------------------------
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            const int r = warp * 32 + p * 4 + rsub;
            const int64_t m = m0 + r;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (m < g.M && k < g.K) {
                const float* src = g.A + m * g.lda + k;
                if (a_vec && k + 3 < g.K) {
                    const float4 t = *reinterpret_cast<const float4*>(src);
                    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
                    if (g.a_pre) {
                        const float4 q = *reinterpret_cast<const float4*>(g.a_pre + m * g.lda + k);
                        v[0] *= spk_act_grad(q.x, g.a_act);
                        v[1] *= spk_act_grad(q.y, g.a_act);
                        v[2] *= spk_act_grad(q.z, g.a_act);
                        v[3] *= spk_act_grad(q.w, g.a_act);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (k + i < g.K) {
                            v[i] = src[i];
                            if (g.a_pre) v[i] *= spk_act_grad(g.a_pre[m * g.lda + k + i], g.a_act);
                        }
                }
            }
            float4 hi, lo;
            hi.x = tf32_rn(v[0]); hi.y = tf32_rn(v[1]); hi.z = tf32_rn(v[2]); hi.w = tf32_rn(v[3]);
            lo.x = v[0] - hi.x; lo.y = v[1] - hi.y; lo.z = v[2] - hi.z; lo.w = v[3] - hi.w;
            const int off = tile_off(r, chunk);
            *reinterpret_cast<float4*>(st + off) = hi;
            *reinterpret_cast<float4*>(st + OPER_BYTES + off) = lo;
        }
        // ---- W tile: rows (output features) n0 + r ---------------------------------------------------------------
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            const int r = warp * 32 + p * 4 + rsub;
            if (r >= BN) continue;
            const int n = n0 + r;
            float4 hi = make_float4(0.f, 0.f, 0.f, 0.f), lo = hi;
            if (n < g.N && k < g.K) {
                const int64_t o = (int64_t)n * g.K + k;
                if (w_vec && k + 3 < g.K) {
                    hi = *reinterpret_cast<const float4*>(g.Wh + o);
                    lo = *reinterpret_cast<const float4*>(g.Wl + o);
                } else {
                    float h[4] = {0.f, 0.f, 0.f, 0.f}, l[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (k + i < g.K) {
                            h[i] = g.Wh[o + i];
                            l[i] = g.Wl[o + i];
                        }
                    hi = make_float4(h[0], h[1], h[2], h[3]);
                    lo = make_float4(l[0], l[1], l[2], l[3]);
                }
            }
            const int off = tile_off(r, chunk);
            *reinterpret_cast<float4*>(st + 2 * OPER_BYTES + off) = hi;
            *reinterpret_cast<float4*>(st + 3 * OPER_BYTES + off) = lo;
        }
        // make the generic-proxy stores visible to the tensor core (async proxy), then hand over to the issuer
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (kt >= 1) {               // main accumulator of the previous k-tile -> registers, before it is overwritten
            drain(kt - 1);
            __syncthreads();
        }
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t sa = smem_u32(st);
#pragma unroll
            for (int ks = 0; ks < TK / 8; ++ks) {
                // k-step ks covers 16 B chunks 2ks, 2ks+1  -> advance the start address by 2 planes
                const uint64_t ah = make_desc(sa + 2 * ks * PLANE);
                const uint64_t al = make_desc(sa + OPER_BYTES + 2 * ks * PLANE);
                const uint64_t bh = make_desc(sa + 2 * OPER_BYTES + 2 * ks * PLANE);
                const uint64_t bl = make_desc(sa + 3 * OPER_BYTES + 2 * ks * PLANE);
                umma_tf32(tmem_base + TN, al, bh, idesc, (kt | ks) ? 1u : 0u);   // small terms: own accumulator
                umma_tf32(tmem_base + TN, ah, bl, idesc, 1u);
                umma_tf32(tmem_base, ah, bh, idesc, ks ? 1u : 0u);               // main term: fresh every k-tile
            }
            // arrives on the stage barrier when every MMA issued so far has completed (implies before_thread_sync)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                smem_u32(&bar_mma[s])) : "memory");
        }
    }
    // ---- last k-tile, then the correction accumulator (lo terms of all k-tiles), then the epilogue -----------------
    drain(nk - 1);
    const int64_t m = m0 + tid;                     // TMEM lane == output row
#pragma unroll
    for (int c0 = 0; c0 < TN; c0 += 32) {
        if (c0 >= BN) break;
        uint32_t r[32];
        tmem_ld32(lane_addr + (uint32_t)(TN + c0), r);
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(accr[c0 + j] + __uint_as_float(r[j]));
        if (m < g.M) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const int n = n0 + c0 + j;
                if (n >= g.N) break;
                float v[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[j + i]) + ((g.bias && n + i < g.N) ? g.bias[n + i] : 0.f);
                const bool full = (n + 3 < g.N);
                if (g.y_pre) {
                    float* p = g.y_pre + m * g.ldy + n;
                    if (full && ((g.ldy & 3) == 0)) *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
                    else
                        for (int i = 0; i < 4; ++i)
                            if (n + i < g.N) p[i] = v[i];
                }
                if (g.act != SPK_ACT_NONE) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) v[i] = spk_act(v[i], g.act);
                }
                if (g.addend) {
                    const float* a = g.addend + m * g.ld_add + n;
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (n + i < g.N) v[i] += a[i];
                }
                float* y = g.Y + m * g.ldy + n;
                if (full && ((g.ldy & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.Y) & 15) == 0))
                    *reinterpret_cast<float4*>(y) = make_float4(v[0], v[1], v[2], v[3]);
                else
                    for (int i = 0; i < 4; ++i)
                        if (n + i < g.N) y[i] = v[i];
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * TN));
    }
}

}  // namespace

extern "C" int spk_dense_tc(const float* A, int64_t M, int K, int64_t lda, const float* a_pre, int a_act,
                            const float* W_hi, const float* W_lo, int N, const float* bias, int act, const float* addend,
                            int64_t ld_add, float* Y, int64_t ldy, float* y_pre, spk_stream_t stream) {
    if (M < 0 || K <= 0 || N <= 0 || lda < K || ldy < N) return SPK_ERR_ARG;
    if (act < 0 || act > 2 || a_act < 0 || a_act > 2) return SPK_ERR_ARG;
    if (M == 0) return SPK_OK;
    if (!A || !W_hi || !W_lo || !Y) return SPK_ERR_ARG;
    if (addend && ld_add < N) return SPK_ERR_ARG;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_dense_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return SPK_CUDA_ERR(e);
        attr_set = true;
    }
    TcArgs g;
    g.A = A; g.a_pre = a_pre; g.Wh = W_hi; g.Wl = W_lo; g.bias = bias; g.addend = addend; g.Y = Y; g.y_pre = y_pre;
    g.M = M; g.lda = lda; g.ld_add = ld_add; g.ldy = ldy; g.K = K; g.N = N; g.a_act = a_act; g.act = act;
    dim3 grid((unsigned)spk_cdiv(M, TM), (unsigned)spk_cdiv(N, TN));
    k_dense_tc<<<grid, 128, SMEM_BYTES, spk_st(stream)>>>(g);
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}
