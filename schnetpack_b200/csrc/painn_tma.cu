// PaiNN fused edge kernels, TMA + mbarrier pipeline variant (default).
//
// Why: ncu (profiles/r1_ncu_edge_kernels_v1.csv and the source page) showed the LDG variant stalled ~45 % on
// long_scoreboard: the sender rows x[j] / mu[j] sit in L2 (~600+ cycles) and register prefetching does not help because
// the six scoreboard slots alias -- waiting for ANY load drains the prefetched ones too.  Here the per-edge operands are
// moved by the TMA unit (cp.async.bulk global->shared, SASS UBLKCP) and tracked by mbarriers, not scoreboards:
//
//   * one PRODUCER warp per CTA walks the CTA's edge range up to DEPTH edges ahead.  Per edge it issues one bulk copy per
//     operand row (x[j] 3F floats, mu[j] 3F floats, the radial-basis row, the geometry record) into a ring stage and
//     arms the stage's `full` mbarrier with the byte count; it waits on the stage's `empty` mbarrier before reuse;
//   * F CONSUMER threads (thread c = feature channel c of all three filter thirds, filter weights in registers) wait on
//     `full`, read their operands with conflict-free LDS, do the filter (packed FFMA2), message and register-accumulated
//     segmented reduction, and release the stage (`empty`, one elected arrive per warp).  No block-wide barriers in the
//     forward kernel at all; the reverse kernel keeps one consumer-only named barrier per 32 edges for the cross-warp
//     reduction of the four per-edge scalars.
#include "painn_common.cuh"
#include "tcgen05.cuh"

namespace {

constexpr int DEPTH = 8;     // ring stages = edges in flight per CTA
constexpr int RCH = 32;      // reverse kernel: edges per cross-warp reduction chunk
constexpr int FCH = 32;      // forward kernel: slots per radial-basis/geometry chunk stage

__device__ __forceinline__ void consumer_bar(int nthreads) {
    asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
}

// ------------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------------
template <int F, bool HAS_MU>
struct FwdStage {                 // per-edge ring stage: the sender's rows
    float x[3 * F];
    float mu[HAS_MU ? 3 * F : 4];
};
template <int NRB>
struct FwdChunk {                 // per-chunk stage: radial basis and geometry records of FCH consecutive slots
    float phi[FCH * NRB];
    float geo[FCH * SPK_GEO_STRIDE];
};

template <int NW, int NRB, bool HAS_MU>
__global__ void __launch_bounds__((NW + 1) * 32) k_painn_edge_fwd_tma(
    const float* __restrict__ x, const float* __restrict__ mu, const float* __restrict__ q,
    const float* __restrict__ phi, const float* __restrict__ geo, const int* __restrict__ rowptr,
    const int* __restrict__ slot_j, const float* __restrict__ wf, const float* __restrict__ bf, int n_atoms,
    int n_edges, int n_rbf, float* __restrict__ q_out, float* __restrict__ mu_out) {
    SPK_PDL_ENTER();
    constexpr int F = NW * 32;
    using Stage = FwdStage<F, HAS_MU>;
    using Chunk = FwdChunk<NRB>;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    Stage* stages = reinterpret_cast<Stage*>(smem_raw);
    Chunk* chunks = reinterpret_cast<Chunk*>(smem_raw + DEPTH * sizeof(Stage));
    __shared__ __align__(8) uint64_t full_bar[DEPTH];
    __shared__ __align__(8) uint64_t empty_bar[DEPTH];
    __shared__ __align__(8) uint64_t cfull_bar[2];
    __shared__ __align__(8) uint64_t cempty_bar[2];

    const int tid = threadIdx.x;
    const int nb = gridDim.x, b = blockIdx.x;
    const int row_lo = spk_block_row_begin(rowptr, n_atoms, n_edges, nb, b);
    const int row_hi = spk_block_row_begin(rowptr, n_atoms, n_edges, nb, b + 1);
    if (row_lo >= row_hi) return;
    const int s_begin = rowptr[row_lo], s_end = rowptr[row_hi];
    const int n_e = s_end - s_begin;
    const int KP = spk_kp(n_rbf);

    if (tid == 0) {
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
            mbar_init(&full_bar[d], 1);
            mbar_init(&empty_bar[d], NW);
        }
        mbar_init(&cfull_bar[0], 1);
        mbar_init(&cfull_bar[1], 1);
        mbar_init(&cempty_bar[0], NW);
        mbar_init(&cempty_bar[1], NW);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (tid >= F) {
        // ===================== producer warp =====================
        // 2 bulk copies per edge (x[j], mu[j]) + 2 per chunk of FCH slots (phi rows, geo records; KP == NRB required)
        const int lane = tid - F;
        const uint32_t x_bytes = (HAS_MU ? 3 : 2) * F * 4;
        const uint32_t e_bytes = x_bytes + (HAS_MU ? 3 * F * 4 : 0);
        for (int e0 = 0; e0 < n_e; e0 += 32) {
            const int jv = (e0 + lane < n_e) ? slot_j[s_begin + e0 + lane] : 0;
            const int cnt = min(32, n_e - e0);
            for (int l = 0; l < cnt; ++l) {
                const int j = __shfl_sync(0xffffffffu, jv, l);
                if (lane == 0) {
                    const int e = e0 + l;
                    if (e % FCH == 0) {   // stage the records of chunk e / FCH
                        const int ck = e / FCH, cb = ck & 1;
                        if (ck >= 2) mbar_wait(&cempty_bar[cb], ((ck >> 1) - 1) & 1);
                        const int cn = min(FCH, n_e - e);
                        const int64_t s0 = s_begin + e;
                        mbar_expect_tx(&cfull_bar[cb], (uint32_t)cn * (NRB + SPK_GEO_STRIDE) * 4);
                        tma_load(chunks[cb].phi, phi + s0 * KP, (uint32_t)cn * NRB * 4, &cfull_bar[cb]);
                        tma_load(chunks[cb].geo, geo + s0 * SPK_GEO_STRIDE, (uint32_t)cn * SPK_GEO_STRIDE * 4, &cfull_bar[cb]);
                    }
                    const int st = e % DEPTH, k = e / DEPTH;
                    if (k >= 1) mbar_wait(&empty_bar[st], (k - 1) & 1);
                    Stage& S = stages[st];
                    mbar_expect_tx(&full_bar[st], e_bytes);
                    tma_load(S.x, x + (size_t)j * (3 * F), x_bytes, &full_bar[st]);
                    if (HAS_MU) tma_load(S.mu, mu + (size_t)j * (3 * F), 3 * F * 4, &full_bar[st]);
                }
            }
        }
        return;
    }

    // ===================== consumers =====================
    const int c = tid, lane = tid & 31;
    FilterRegs<NRB> w;
    load_filter<NRB>(w, wf, bf, F, n_rbf, c);

    int i = row_lo;
    int next_boundary = rowptr[i + 1];
    int boundary2 = rowptr[min(i + 2, n_atoms)];
    float dq = 0.f, dm0 = 0.f, dm1 = 0.f, dm2 = 0.f;
    float rq, rm0 = 0.f, rm1 = 0.f, rm2 = 0.f;
    auto load_res = [&](int row) {
        rq = q[(size_t)row * F + c];
        if (HAS_MU) {
            const float* __restrict__ mr = mu + (size_t)row * (3 * F) + c;
            rm0 = mr[0];
            rm1 = mr[F];
            rm2 = mr[2 * F];
        }
    };
    load_res(i);
    auto flush_advance = [&]() {
        q_out[(size_t)i * F + c] = rq + dq;
        float* __restrict__ mo = mu_out + (size_t)i * (3 * F) + c;
        mo[0] = rm0 + dm0;
        mo[F] = rm1 + dm1;
        mo[2 * F] = rm2 + dm2;
        dq = dm0 = dm1 = dm2 = 0.f;
        ++i;
        next_boundary = boundary2;
        boundary2 = rowptr[min(i + 2, n_atoms)];
        if (i < row_hi) load_res(i);
    };

    for (int e = 0; e < n_e; ++e) {
        const int s = s_begin + e;
        const int st = e % DEPTH;
        const int ck = e / FCH, cb = ck & 1, t = e - ck * FCH;
        while (s >= next_boundary) flush_advance();
        if (t == 0) mbar_wait(&cfull_bar[cb], (ck >> 1) & 1);
        mbar_wait(&full_bar[st], (e / DEPTH) & 1);
        const Stage& S = stages[st];
        const Chunk& C = chunks[cb];
        const float xa = S.x[c], xb = S.x[F + c];
        float xc = 0.f, m0 = 0.f, m1 = 0.f, m2 = 0.f;
        if (HAS_MU) {
            xc = S.x[2 * F + c];
            m0 = S.mu[c];
            m1 = S.mu[F + c];
            m2 = S.mu[2 * F + c];
        }
        // the sender rows are in registers: hand the ring stage back to the producer
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[st]);
        const float4 g0 = *reinterpret_cast<const float4*>(C.geo + t * SPK_GEO_STRIDE);      // ux uy uz d
        const float fc = C.geo[t * SPK_GEO_STRIDE + 4];
        float2 pa2 = make_float2(w.ba, 0.f), pb2 = make_float2(w.bb, 0.f), pc2 = make_float2(w.bc, 0.f);
        const float4* __restrict__ ph = reinterpret_cast<const float4*>(C.phi + t * NRB);
#pragma unroll
        for (int k4 = 0; k4 < NRB / 4; ++k4) {
            const float4 p = ph[k4];
            const float2 p01 = make_float2(p.x, p.y), p23 = make_float2(p.z, p.w);
            pa2 = __ffma2_rn(p01, w.a[2 * k4], pa2);
            pb2 = __ffma2_rn(p01, w.b[2 * k4], pb2);
            pa2 = __ffma2_rn(p23, w.a[2 * k4 + 1], pa2);
            pb2 = __ffma2_rn(p23, w.b[2 * k4 + 1], pb2);
            if (HAS_MU) {
                pc2 = __ffma2_rn(p01, w.c[2 * k4], pc2);
                pc2 = __ffma2_rn(p23, w.c[2 * k4 + 1], pc2);
            }
        }
        if (t == FCH - 1 || e == n_e - 1) {     // last edge of the chunk: release the chunk stage
            __syncwarp();
            if (lane == 0) mbar_arrive(&cempty_bar[cb]);
        }
        const float pa = pa2.x + pa2.y, pb = pb2.x + pb2.y, pc = pc2.x + pc2.y;
        dq = fmaf(fc * pa, xa, dq);
        const float tb = fc * pb * xb;
        dm0 = fmaf(tb, g0.x, dm0);
        dm1 = fmaf(tb, g0.y, dm1);
        dm2 = fmaf(tb, g0.z, dm2);
        if (HAS_MU) {
            const float tc = fc * pc * xc;
            dm0 = fmaf(tc, m0, dm0);
            dm1 = fmaf(tc, m1, dm1);
            dm2 = fmaf(tc, m2, dm2);
        }
    }
    while (i < row_hi) flush_advance();
}

// ------------------------------------------------------------------------------------------------------------------
// reverse (grouped by sender)
// ------------------------------------------------------------------------------------------------------------------
template <int F, int NRB>
struct BwdStage {
    float gq[F];
    float gmu[3 * F];
    float phi[NRB];                  // the combined per-slot record [phi | dphi | geo] lands here with ONE bulk copy
    float dphi[NRB];
    float geo[SPK_GEO_STRIDE];
};

template <int NW, int NRB, bool HAS_MU>
__global__ void __launch_bounds__((NW + 1) * 32) k_painn_edge_bwd_tma(
    const float* __restrict__ x, const float* __restrict__ mu, const float* __restrict__ g_q,
    const float* __restrict__ g_mu, const float* __restrict__ erec, const int* __restrict__ sptr,
    const int* __restrict__ pos_slot, const int* __restrict__ pos_i, const int* __restrict__ slot_eid,
    const float* __restrict__ wf, const float* __restrict__ bf, int n_atoms, int n_edges, int n_rbf,
    float* __restrict__ g_x, float* __restrict__ g_mu_in, float* __restrict__ g_rij, int accumulate) {
    SPK_PDL_ENTER();
    constexpr int F = NW * 32;
    constexpr int REC = 2 * NRB + SPK_GEO_STRIDE;
    using Stage = BwdStage<F, NRB>;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    Stage* stages = reinterpret_cast<Stage*>(smem_raw);
    __shared__ __align__(8) uint64_t full_bar[DEPTH];
    __shared__ __align__(8) uint64_t empty_bar[DEPTH];
    __shared__ float s_red[RCH][NW][4];
    __shared__ float s_fin[RCH][4];      // ux uy uz 1/d of the chunk's edges

    const int tid = threadIdx.x;
    const int nb = gridDim.x, b = blockIdx.x;
    const int j_lo = spk_block_row_begin(sptr, n_atoms, n_edges, nb, b);
    const int j_hi = spk_block_row_begin(sptr, n_atoms, n_edges, nb, b + 1);
    if (j_lo >= j_hi) return;
    const int p_begin = sptr[j_lo], p_end = sptr[j_hi];
    const int n_e = p_end - p_begin;

    if (tid == 0) {
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
            mbar_init(&full_bar[d], 1);
            mbar_init(&empty_bar[d], NW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (tid >= F) {
        // ===================== producer warp =====================
        const int lane = tid - F;
        const uint32_t bytes = F * 4 + 3 * F * 4 + REC * 4;
        for (int e0 = 0; e0 < n_e; e0 += 32) {
            const bool ok = e0 + lane < n_e;
            const int iv = ok ? pos_i[p_begin + e0 + lane] : 0;
            const int sv = ok ? pos_slot[p_begin + e0 + lane] : 0;
            const int cnt = min(32, n_e - e0);
            for (int l = 0; l < cnt; ++l) {
                const int i = __shfl_sync(0xffffffffu, iv, l);
                const int64_t s = __shfl_sync(0xffffffffu, sv, l);
                if (lane == 0) {
                    const int e = e0 + l, st = e % DEPTH, k = e / DEPTH;
                    if (k >= 1) mbar_wait(&empty_bar[st], (k - 1) & 1);
                    Stage& S = stages[st];
                    mbar_expect_tx(&full_bar[st], bytes);
                    tma_load(S.gq, g_q + (size_t)i * F, F * 4, &full_bar[st]);
                    tma_load(S.gmu, g_mu + (size_t)i * (3 * F), 3 * F * 4, &full_bar[st]);
                    tma_load(S.phi, erec + s * REC, REC * 4, &full_bar[st]);
                }
            }
        }
        return;
    }

    // ===================== consumers =====================
    const int c = tid, lane = tid & 31, warp = tid >> 5;
    FilterRegs<NRB> w;
    load_filter<NRB>(w, wf, bf, F, n_rbf, c);

    int j = j_lo;
    int next_boundary = sptr[j + 1];
    int boundary2 = sptr[min(j + 2, n_atoms)];
    float xa, xb, xc = 0.f, m0 = 0.f, m1 = 0.f, m2 = 0.f, r0 = 0.f, r1 = 0.f, r2 = 0.f;
    auto load_own = [&](int row) {
        const float* __restrict__ xr = x + (size_t)row * (3 * F) + c;
        xa = xr[0];
        xb = xr[F];
        if (HAS_MU) {
            xc = xr[2 * F];
            const float* __restrict__ mr = mu + (size_t)row * (3 * F) + c;
            m0 = mr[0];
            m1 = mr[F];
            m2 = mr[2 * F];
            const float* __restrict__ gr = g_mu + (size_t)row * (3 * F) + c;
            r0 = gr[0];
            r1 = gr[F];
            r2 = gr[2 * F];
        }
    };
    load_own(j);
    float gxa = 0.f, gxb = 0.f, gxc = 0.f, gm0 = 0.f, gm1 = 0.f, gm2 = 0.f;
    auto flush_advance = [&]() {
        float* __restrict__ gx = g_x + (size_t)j * (3 * F) + c;
        gx[0] = gxa;
        gx[F] = gxb;
        gx[2 * F] = gxc;
        if (HAS_MU) {
            float* __restrict__ gm = g_mu_in + (size_t)j * (3 * F) + c;
            gm[0] = r0 + gm0;
            gm[F] = r1 + gm1;
            gm[2 * F] = r2 + gm2;
        }
        gxa = gxb = gxc = gm0 = gm1 = gm2 = 0.f;
        ++j;
        next_boundary = boundary2;
        boundary2 = sptr[min(j + 2, n_atoms)];
        if (j < j_hi) load_own(j);
    };

    for (int cs = 0; cs < n_e; cs += RCH) {
        const int n = min(RCH, n_e - cs);
        for (int t = 0; t < n; ++t) {
            const int e = cs + t;
            const int p = p_begin + e;
            const int st = e % DEPTH;
            while (p >= next_boundary) flush_advance();
            mbar_wait(&full_bar[st], (e / DEPTH) & 1);
            const Stage& S = stages[st];
            const float gq = S.gq[c];
            const float g0 = S.gmu[c], g1 = S.gmu[F + c], g2 = S.gmu[2 * F + c];
            const float4 ge = *reinterpret_cast<const float4*>(S.geo);   // ux uy uz d
            const float fc = S.geo[4], dfc = S.geo[5], inv = S.geo[6];
            float2 pa2 = make_float2(w.ba, 0.f), pb2 = make_float2(w.bb, 0.f), pc2 = make_float2(w.bc, 0.f);
            float2 da2 = make_float2(0.f, 0.f), db2 = da2, dc2 = da2;
            const float4* __restrict__ ph = reinterpret_cast<const float4*>(S.phi);
            const float4* __restrict__ dh = reinterpret_cast<const float4*>(S.dphi);
#pragma unroll
            for (int k4 = 0; k4 < NRB / 4; ++k4) {
                const float4 p4 = ph[k4];
                const float4 d4 = dh[k4];
                const float2 p01 = make_float2(p4.x, p4.y), p23 = make_float2(p4.z, p4.w);
                const float2 d01 = make_float2(d4.x, d4.y), d23 = make_float2(d4.z, d4.w);
                pa2 = __ffma2_rn(p01, w.a[2 * k4], pa2);
                da2 = __ffma2_rn(d01, w.a[2 * k4], da2);
                pb2 = __ffma2_rn(p01, w.b[2 * k4], pb2);
                db2 = __ffma2_rn(d01, w.b[2 * k4], db2);
                pa2 = __ffma2_rn(p23, w.a[2 * k4 + 1], pa2);
                da2 = __ffma2_rn(d23, w.a[2 * k4 + 1], da2);
                pb2 = __ffma2_rn(p23, w.b[2 * k4 + 1], pb2);
                db2 = __ffma2_rn(d23, w.b[2 * k4 + 1], db2);
                if (HAS_MU) {
                    pc2 = __ffma2_rn(p01, w.c[2 * k4], pc2);
                    dc2 = __ffma2_rn(d01, w.c[2 * k4], dc2);
                    pc2 = __ffma2_rn(p23, w.c[2 * k4 + 1], pc2);
                    dc2 = __ffma2_rn(d23, w.c[2 * k4 + 1], dc2);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[st]);
            if (c == 0) *reinterpret_cast<float4*>(s_fin[t]) = make_float4(ge.x, ge.y, ge.z, inv);
            const float pa = pa2.x + pa2.y, pb = pb2.x + pb2.y, pc = pc2.x + pc2.y;
            const float da = da2.x + da2.y, db = db2.x + db2.y, dc = dc2.x + dc2.y;
            const float Wa = fc * pa, Wb = fc * pb;
            const float dWa = fmaf(dfc, pa, fc * da), dWb = fmaf(dfc, pb, fc * db);
            const float gu = g0 * ge.x + g1 * ge.y + g2 * ge.z;   // sum_d g_mu[i,d] u_d
            gxa = fmaf(Wa, gq, gxa);
            gxb = fmaf(Wb, gu, gxb);
            float part_d = gq * xa * dWa + gu * xb * dWb;
            const float wbx = Wb * xb;
            float pu0 = g0 * wbx, pu1 = g1 * wbx, pu2 = g2 * wbx;
            if (HAS_MU) {
                const float Wc = fc * pc;
                const float dWc = fmaf(dfc, pc, fc * dc);
                const float gm = g0 * m0 + g1 * m1 + g2 * m2;     // sum_d g_mu[i,d] mu[j,d]
                gxc = fmaf(Wc, gm, gxc);
                const float wcx = Wc * xc;
                gm0 = fmaf(wcx, g0, gm0);
                gm1 = fmaf(wcx, g1, gm1);
                gm2 = fmaf(wcx, g2, gm2);
                part_d = fmaf(gm * xc, dWc, part_d);
            }
            part_d = spk_warp_sum(part_d);
            pu0 = spk_warp_sum(pu0);
            pu1 = spk_warp_sum(pu1);
            pu2 = spk_warp_sum(pu2);
            if (lane == 0) *reinterpret_cast<float4*>(s_red[t][warp]) = make_float4(part_d, pu0, pu1, pu2);
        }
        consumer_bar(F);
        if (tid < n) {
            const int t = tid;
            float gd = 0.f, u0 = 0.f, u1 = 0.f, u2 = 0.f;
#pragma unroll
            for (int wv = 0; wv < NW; ++wv) {
                const float4 r = *reinterpret_cast<const float4*>(s_red[t][wv]);
                gd += r.x;
                u0 += r.y;
                u1 += r.z;
                u2 += r.w;
            }
            const float4 fin = *reinterpret_cast<const float4*>(s_fin[t]);
            const float dot = u0 * fin.x + u1 * fin.y + u2 * fin.z;
            float o0 = gd * fin.x + (u0 - dot * fin.x) * fin.w;
            float o1 = gd * fin.y + (u1 - dot * fin.y) * fin.w;
            float o2 = gd * fin.z + (u2 - dot * fin.z) * fin.w;
            const int eid = slot_eid[pos_slot[p_begin + cs + t]];
            float* out = g_rij + (int64_t)eid * 3;
            if (accumulate) {
                o0 += out[0];
                o1 += out[1];
                o2 += out[2];
            }
            out[0] = o0;
            out[1] = o1;
            out[2] = o2;
        }
        consumer_bar(F);
    }
    while (j < j_hi) flush_advance();
}

template <typename K>
int single_wave_grid(K kernel, int threads, size_t smem, int n_atoms, int n_edges, int* cache) {
    if (!*cache) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int occ = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem);
        *cache = occ < 1 ? 1 : occ;
    }
    int64_t nb = (int64_t)spk_num_sms() * (*cache);
    if (nb > spk_cdiv((int64_t)n_edges, 32)) nb = spk_cdiv((int64_t)n_edges, 32);
    if (nb < 1) nb = 1;
    if (nb > n_atoms) nb = n_atoms;
    return (int)nb;
}

}  // namespace

template <int NW, int NRB>
int spk_launch_edge_fwd_tma(const float* x, const float* mu, const float* q, const float* phi, const float* geo,
                            const int* rowptr, const int* slot_j, const float* wf, const float* bf, int n_atoms,
                            int n_edges, int n_rbf, float* q_out, float* mu_out, cudaStream_t st) {
    constexpr int F = NW * 32, T = (NW + 1) * 32;
    static int occ_mu = 0, occ_nomu = 0;
    if (spk_kp(n_rbf) != NRB) return -1;   // chunked phi rows need KP == NRB (n_rbf in 17..20 or 29..32): caller uses LDG
    if (mu) {
        const size_t sm = DEPTH * sizeof(FwdStage<F, true>) + 2 * sizeof(FwdChunk<NRB>);
        int nb = single_wave_grid(k_painn_edge_fwd_tma<NW, NRB, true>, T, sm, n_atoms, n_edges, &occ_mu);
        spk_launch(k_painn_edge_fwd_tma<NW, NRB, true>, nb, T, sm, st, x, mu, q, phi, geo, rowptr, slot_j, wf, bf, n_atoms,
                                                               n_edges, n_rbf, q_out, mu_out);
    } else {
        const size_t sm = DEPTH * sizeof(FwdStage<F, false>) + 2 * sizeof(FwdChunk<NRB>);
        int nb = single_wave_grid(k_painn_edge_fwd_tma<NW, NRB, false>, T, sm, n_atoms, n_edges, &occ_nomu);
        spk_launch(k_painn_edge_fwd_tma<NW, NRB, false>, nb, T, sm, st, x, mu, q, phi, geo, rowptr, slot_j, wf, bf, n_atoms,
                                                                n_edges, n_rbf, q_out, mu_out);
    }
    return 0;
}

template <int NW, int NRB>
int spk_launch_edge_bwd_tma(const float* x, const float* mu, const float* g_q, const float* g_mu, const float* erec,
                            const int* sptr, const int* pos_slot, const int* pos_i,
                            const int* slot_eid, const float* wf, const float* bf, int n_atoms, int n_edges, int n_rbf,
                            float* g_x, float* g_mu_in, float* g_rij, int accumulate, cudaStream_t st) {
    constexpr int F = NW * 32, T = (NW + 1) * 32;
    static int occ_mu = 0, occ_nomu = 0;
    const size_t sm = DEPTH * sizeof(BwdStage<F, NRB>);
    if (mu) {
        int nb = single_wave_grid(k_painn_edge_bwd_tma<NW, NRB, true>, T, sm, n_atoms, n_edges, &occ_mu);
        spk_launch(k_painn_edge_bwd_tma<NW, NRB, true>, nb, T, sm, st, x, mu, g_q, g_mu, erec, sptr, pos_slot, pos_i,
                                                               slot_eid, wf, bf, n_atoms, n_edges, n_rbf, g_x, g_mu_in,
                                                               g_rij, accumulate);
    } else {
        int nb = single_wave_grid(k_painn_edge_bwd_tma<NW, NRB, false>, T, sm, n_atoms, n_edges, &occ_nomu);
        spk_launch(k_painn_edge_bwd_tma<NW, NRB, false>, nb, T, sm, st, x, mu, g_q, g_mu, erec, sptr, pos_slot, pos_i,
                                                                slot_eid, wf, bf, n_atoms, n_edges, n_rbf, g_x, g_mu_in,
                                                                g_rij, accumulate);
    }
    return 0;
}

// explicit instantiations used by the dispatcher in painn.cu
#define INST(NW, NRB)                                                                                                   \
    template int spk_launch_edge_fwd_tma<NW, NRB>(const float*, const float*, const float*, const float*, const float*, \
                                                  const int*, const int*, const float*, const float*, int, int, int,    \
                                                  float*, float*, cudaStream_t);                                        \
    template int spk_launch_edge_bwd_tma<NW, NRB>(const float*, const float*, const float*, const float*, const float*, \
                                                  const int*, const int*, const int*,                                   \
                                                  const int*, const float*, const float*, int, int, int, float*, float*, \
                                                  float*, int, cudaStream_t);
INST(1, 20) INST(2, 20) INST(4, 20) INST(8, 20) INST(1, 32) INST(2, 32) INST(4, 32) INST(8, 32)
#undef INST
