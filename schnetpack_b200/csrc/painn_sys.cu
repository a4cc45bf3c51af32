// PaiNN fused edge kernels, "system-resident" variant for batches of small systems (molecules).
//
// Edges never cross systems (the reference's collate builds one flat graph out of independent systems,
// data/loader.py:35-46), so all senders of a system's edges are the system's own atoms.  When a system's rows fit in
// shared memory (MD17/QM9-sized molecules: 21 atoms x 3 KB = 63 KB), one CTA (or `parts` CTAs) per system stages the
// whole sender table x[a0:a1], mu[a0:a1] (reverse: g_q, g_mu) ONCE with coalesced 128-bit loads and every per-edge
// gather becomes a conflict-free LDS: the ~600-cycle L2 latency per edge that bounds the streaming variants
// (profiles/r1_ncu_edge_kernels_*.csv: 45 % long-scoreboard stalls, one edge in flight per CTA) disappears and each row
// leaves L2 once per CTA instead of once per edge.  Systems larger than the shared-memory capacity chosen at launch fall
// back to global gathers inside the same kernel (uniform branch per CTA), so results never depend on the capacity.
#include "painn_common.cuh"

namespace {

constexpr int CH = 32;  // edges whose radial/geometry records are staged per chunk

template <int NRB, int NTHR>
__device__ __forceinline__ void stage_rows(float* __restrict__ dst, const float* __restrict__ src, int n, int KP) {
    const int kq = KP >> 2;
    for (int t = threadIdx.x; t < n * (NRB / 4); t += NTHR) {
        int r = t / (NRB / 4), q = t - r * (NRB / 4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q < kq) v = *reinterpret_cast<const float4*>(src + (int64_t)r * KP + q * 4);
        *reinterpret_cast<float4*>(dst + r * NRB + q * 4) = v;
    }
}

template <int NW, int NRB, bool HAS_MU>
__global__ void __launch_bounds__(NW * 32) k_painn_edge_fwd_sys(
    const float* __restrict__ x, const float* __restrict__ mu, const float* __restrict__ q,
    const float* __restrict__ phi, const float* __restrict__ geo, const int* __restrict__ rowptr,
    const int* __restrict__ slot_j, const float* __restrict__ wf, const float* __restrict__ bf,
    const int* __restrict__ mol_ptr, int parts, int cap_atoms, int n_rbf, float* __restrict__ q_out,
    float* __restrict__ mu_out) {
    SPK_PDL_ENTER();
    constexpr int F = NW * 32;
    constexpr int NTHR = NW * 32;
    extern __shared__ __align__(16) float s_tab[];           // [cap][3F] x rows, then [cap][3F] mu rows
    __shared__ __align__(16) float s_phi[CH * NRB];
    __shared__ __align__(16) float s_geo[CH * SPK_GEO_STRIDE];
    __shared__ int s_j[CH];

    const int c = threadIdx.x;
    const int m = blockIdx.x / parts, part = blockIdx.x - m * parts;
    const int a0 = mol_ptr[m], a1 = mol_ptr[m + 1];
    const int nA = a1 - a0;
    if (nA <= 0) return;
    const bool resident = nA <= cap_atoms;
    float* s_x = s_tab;
    float* s_mu = s_tab + (size_t)cap_atoms * 3 * F;
    if (resident) {
        const float4* gx = reinterpret_cast<const float4*>(x + (size_t)a0 * 3 * F);
        for (int t = threadIdx.x; t < nA * (3 * F / 4); t += NTHR) reinterpret_cast<float4*>(s_x)[t] = gx[t];
        if (HAS_MU) {
            const float4* gm = reinterpret_cast<const float4*>(mu + (size_t)a0 * 3 * F);
            for (int t = threadIdx.x; t < nA * (3 * F / 4); t += NTHR) reinterpret_cast<float4*>(s_mu)[t] = gm[t];
        }
    }
    const int row_lo = a0 + (int)(((long long)nA * part) / parts);
    const int row_hi = a0 + (int)(((long long)nA * (part + 1)) / parts);
    if (row_lo >= row_hi) return;        // (whole CTA: uniform)

    FilterRegs<NRB> w;
    load_filter<NRB>(w, wf, bf, F, n_rbf, c);
    const int KP = spk_kp(n_rbf);

    const int s_begin = rowptr[row_lo], s_end = rowptr[row_hi];
    int i = row_lo;
    int next_boundary = rowptr[i + 1];
    float dq = 0.f, dm0 = 0.f, dm1 = 0.f, dm2 = 0.f;
    auto flush = [&](int row) {
        const size_t o = (size_t)row * F + c;
        q_out[o] = q[o] + dq;
        const size_t om = (size_t)row * 3 * F + c;
        if (HAS_MU) {
            // the receiver's own mu row: from the staged table when resident
            const float* __restrict__ mr = resident ? (s_mu + (size_t)(row - a0) * 3 * F + c) : (mu + om);
            mu_out[om] = mr[0] + dm0;
            mu_out[om + F] = mr[F] + dm1;
            mu_out[om + 2 * F] = mr[2 * F] + dm2;
        } else {
            mu_out[om] = dm0;
            mu_out[om + F] = dm1;
            mu_out[om + 2 * F] = dm2;
        }
        dq = dm0 = dm1 = dm2 = 0.f;
    };

    for (int cs = s_begin; cs < s_end; cs += CH) {
        const int n = min(CH, s_end - cs);
        __syncthreads();                                   // also orders the table staging before the first use
        stage_rows<NRB, NTHR>(s_phi, phi + (int64_t)cs * KP, n, KP);
        for (int t = threadIdx.x; t < n * 2; t += NTHR)
            reinterpret_cast<float4*>(s_geo)[t] = reinterpret_cast<const float4*>(geo + (int64_t)cs * SPK_GEO_STRIDE)[t];
        for (int t = threadIdx.x; t < n; t += NTHR) s_j[t] = slot_j[cs + t];
        __syncthreads();

#pragma unroll 2
        for (int t = 0; t < n; ++t) {
            const int s = cs + t;
            while (s >= next_boundary) {
                flush(i);
                ++i;
                next_boundary = rowptr[i + 1];
            }
            const int j = s_j[t];
            float xa, xb, xc = 0.f, m0 = 0.f, m1 = 0.f, m2 = 0.f;
            if (resident) {
                const float* __restrict__ xj = s_x + (size_t)(j - a0) * 3 * F + c;
                xa = xj[0];
                xb = xj[F];
                if (HAS_MU) {
                    xc = xj[2 * F];
                    const float* __restrict__ mj = s_mu + (size_t)(j - a0) * 3 * F + c;
                    m0 = mj[0];
                    m1 = mj[F];
                    m2 = mj[2 * F];
                }
            } else {
                const float* __restrict__ xj = x + (size_t)j * 3 * F + c;
                xa = xj[0];
                xb = xj[F];
                if (HAS_MU) {
                    xc = xj[2 * F];
                    const float* __restrict__ mj = mu + (size_t)j * 3 * F + c;
                    m0 = mj[0];
                    m1 = mj[F];
                    m2 = mj[2 * F];
                }
            }
            const float4 g0 = *reinterpret_cast<const float4*>(s_geo + t * SPK_GEO_STRIDE);      // ux uy uz d
            const float fc = s_geo[t * SPK_GEO_STRIDE + 4];
            float2 pa2 = make_float2(w.ba, 0.f), pb2 = make_float2(w.bb, 0.f), pc2 = make_float2(w.bc, 0.f);
            const float4* __restrict__ ph = reinterpret_cast<const float4*>(s_phi + t * NRB);
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
    }
    if (s_begin == s_end) __syncthreads();               // table staging visible before the flushes read it
    for (; i < row_hi; ++i) flush(i);
}

// ------------------------------------------------------------------------------------------------------------------
// reverse, grouped by sender: the system's receiver-gradient rows g_q[a0:a1], g_mu[a0:a1] are staged
// ------------------------------------------------------------------------------------------------------------------
template <int NW, int NRB, bool HAS_MU>
__global__ void __launch_bounds__(NW * 32) k_painn_edge_bwd_sys(
    const float* __restrict__ x, const float* __restrict__ mu, const float* __restrict__ g_q,
    const float* __restrict__ g_mu, const float* __restrict__ phi, const float* __restrict__ dphi,
    const float* __restrict__ geo, const int* __restrict__ sptr, const int* __restrict__ pos_slot,
    const int* __restrict__ pos_i, const int* __restrict__ slot_eid, const float* __restrict__ wf,
    const float* __restrict__ bf, const int* __restrict__ mol_ptr, int parts, int cap_atoms, int n_rbf,
    float* __restrict__ g_x, float* __restrict__ g_mu_in, float* __restrict__ g_rij, int accumulate) {
    SPK_PDL_ENTER();
    constexpr int F = NW * 32;
    constexpr int NTHR = NW * 32;
    extern __shared__ __align__(16) float s_tab[];           // [cap][F] g_q rows, then [cap][3F] g_mu rows
    __shared__ __align__(16) float s_phi[CH * NRB];
    __shared__ __align__(16) float s_dphi[CH * NRB];
    __shared__ __align__(16) float s_geo[CH * SPK_GEO_STRIDE];
    __shared__ int s_i[CH];
    __shared__ int s_eid[CH];
    __shared__ float s_red[CH][NW][4];

    const int c = threadIdx.x;
    const int lane = c & 31, warp = c >> 5;
    const int m = blockIdx.x / parts, part = blockIdx.x - m * parts;
    const int a0 = mol_ptr[m], a1 = mol_ptr[m + 1];
    const int nA = a1 - a0;
    if (nA <= 0) return;
    const bool resident = nA <= cap_atoms;
    float* s_gq = s_tab;
    float* s_gmu = s_tab + (size_t)cap_atoms * F;
    if (resident) {
        const float4* gq4 = reinterpret_cast<const float4*>(g_q + (size_t)a0 * F);
        for (int t = threadIdx.x; t < nA * (F / 4); t += NTHR) reinterpret_cast<float4*>(s_gq)[t] = gq4[t];
        const float4* gm4 = reinterpret_cast<const float4*>(g_mu + (size_t)a0 * 3 * F);
        for (int t = threadIdx.x; t < nA * (3 * F / 4); t += NTHR) reinterpret_cast<float4*>(s_gmu)[t] = gm4[t];
    }
    const int j_lo = a0 + (int)(((long long)nA * part) / parts);
    const int j_hi = a0 + (int)(((long long)nA * (part + 1)) / parts);
    if (j_lo >= j_hi) return;

    FilterRegs<NRB> w;
    load_filter<NRB>(w, wf, bf, F, n_rbf, c);
    const int KP = spk_kp(n_rbf);
    const int kq = KP >> 2;

    const int p_begin = sptr[j_lo], p_end = sptr[j_hi];
    int j = j_lo;
    int next_boundary = sptr[j + 1];
    float xa, xb, xc = 0.f, m0 = 0.f, m1 = 0.f, m2 = 0.f;
    float gxa = 0.f, gxb = 0.f, gxc = 0.f, gm0 = 0.f, gm1 = 0.f, gm2 = 0.f;
    auto load_own = [&](int row) {
        const float* __restrict__ xr = x + (size_t)row * 3 * F + c;
        xa = xr[0];
        xb = xr[F];
        if (HAS_MU) {
            xc = xr[2 * F];
            const float* __restrict__ mr = mu + (size_t)row * 3 * F + c;
            m0 = mr[0];
            m1 = mr[F];
            m2 = mr[2 * F];
        }
    };
    auto flush = [&](int row) {
        const size_t o = (size_t)row * 3 * F + c;
        g_x[o] = gxa;
        g_x[o + F] = gxb;
        g_x[o + 2 * F] = gxc;
        if (HAS_MU) {
            const float* __restrict__ gr = resident ? (s_gmu + (size_t)(row - a0) * 3 * F + c) : (g_mu + o);
            g_mu_in[o] = gr[0] + gm0;
            g_mu_in[o + F] = gr[F] + gm1;
            g_mu_in[o + 2 * F] = gr[2 * F] + gm2;
        }
        gxa = gxb = gxc = gm0 = gm1 = gm2 = 0.f;
    };
    load_own(j);

    for (int cs = p_begin; cs < p_end; cs += CH) {
        const int n = min(CH, p_end - cs);
        __syncthreads();
        for (int t = threadIdx.x; t < n * (NRB / 4); t += NTHR) {
            int r = t / (NRB / 4), qd = t - r * (NRB / 4);
            int s = pos_slot[cs + r];
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f), dv = v;
            if (qd < kq) {
                v = *reinterpret_cast<const float4*>(phi + (int64_t)s * KP + qd * 4);
                dv = *reinterpret_cast<const float4*>(dphi + (int64_t)s * KP + qd * 4);
            }
            *reinterpret_cast<float4*>(s_phi + r * NRB + qd * 4) = v;
            *reinterpret_cast<float4*>(s_dphi + r * NRB + qd * 4) = dv;
        }
        for (int t = threadIdx.x; t < n * 2; t += NTHR) {
            int r = t >> 1;
            int s = pos_slot[cs + r];
            reinterpret_cast<float4*>(s_geo)[t] = reinterpret_cast<const float4*>(geo + (int64_t)s * SPK_GEO_STRIDE)[t & 1];
        }
        for (int t = threadIdx.x; t < n; t += NTHR) {
            s_i[t] = pos_i[cs + t];
            s_eid[t] = slot_eid[pos_slot[cs + t]];
        }
        __syncthreads();

        for (int t = 0; t < n; ++t) {
            const int p = cs + t;
            if (p >= next_boundary) {
                do {
                    flush(j);
                    ++j;
                    next_boundary = sptr[j + 1];
                } while (p >= next_boundary);
                load_own(j);
            }
            const int i = s_i[t];
            float gq, g0, g1, g2;
            if (resident) {
                gq = s_gq[(size_t)(i - a0) * F + c];
                const float* __restrict__ gmi = s_gmu + (size_t)(i - a0) * 3 * F + c;
                g0 = gmi[0];
                g1 = gmi[F];
                g2 = gmi[2 * F];
            } else {
                gq = g_q[(size_t)i * F + c];
                const float* __restrict__ gmi = g_mu + (size_t)i * 3 * F + c;
                g0 = gmi[0];
                g1 = gmi[F];
                g2 = gmi[2 * F];
            }
            const float4 ge = *reinterpret_cast<const float4*>(s_geo + t * SPK_GEO_STRIDE);   // ux uy uz d
            const float fc = s_geo[t * SPK_GEO_STRIDE + 4], dfc = s_geo[t * SPK_GEO_STRIDE + 5];
            float2 pa2 = make_float2(w.ba, 0.f), pb2 = make_float2(w.bb, 0.f), pc2 = make_float2(w.bc, 0.f);
            float2 da2 = make_float2(0.f, 0.f), db2 = da2, dc2 = da2;
            const float4* __restrict__ ph = reinterpret_cast<const float4*>(s_phi + t * NRB);
            const float4* __restrict__ dh = reinterpret_cast<const float4*>(s_dphi + t * NRB);
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
        __syncthreads();
        if (threadIdx.x < n) {
            const int t = threadIdx.x;
            float gd = 0.f, u0 = 0.f, u1 = 0.f, u2 = 0.f;
#pragma unroll
            for (int wv = 0; wv < NW; ++wv) {
                const float4 r = *reinterpret_cast<const float4*>(s_red[t][wv]);
                gd += r.x;
                u0 += r.y;
                u1 += r.z;
                u2 += r.w;
            }
            const float ux = s_geo[t * SPK_GEO_STRIDE + 0], uy = s_geo[t * SPK_GEO_STRIDE + 1],
                        uz = s_geo[t * SPK_GEO_STRIDE + 2], inv = s_geo[t * SPK_GEO_STRIDE + 6];
            const float dot = u0 * ux + u1 * uy + u2 * uz;
            float r0 = gd * ux + (u0 - dot * ux) * inv;
            float r1 = gd * uy + (u1 - dot * uy) * inv;
            float r2 = gd * uz + (u2 - dot * uz) * inv;
            float* out = g_rij + (int64_t)s_eid[t] * 3;
            if (accumulate) {
                r0 += out[0];
                r1 += out[1];
                r2 += out[2];
            }
            out[0] = r0;
            out[1] = r1;
            out[2] = r2;
        }
    }
    if (p_begin == p_end) __syncthreads();
    for (; j < j_hi; ++j) flush(j);
}

// raise the kernel's dynamic shared-memory limit when needed (remembered per kernel: no API call on the steady path)
// (`cur` is the caller's per-kernel-instantiation cache; the default 48 KB limit counts static + dynamic shared memory,
// so the attribute is always raised at least once)
template <typename Kern>
static int set_smem(Kern k, size_t bytes, size_t* cur) {
    if (bytes <= *cur) return 0;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return SPK_CUDA_ERR(e);
    *cur = bytes;
    return 0;
}

template <int NW, int NRB>
int launch_fwd_sys(const float* x, const float* mu, const float* q, const float* phi, const float* geo,
                   const int* rowptr, const int* slot_j, const float* wf, const float* bf, const int* mol_ptr,
                   int n_mol, int parts, int cap, int n_rbf, float* q_out, float* mu_out, cudaStream_t st) {
    constexpr int F = NW * 32;
    const size_t sm = (size_t)cap * 3 * F * 4 * (mu ? 2 : 1);
    static size_t cur_mu = 0, cur_nomu = 0;
    int rc;
    if (mu) {
        if ((rc = set_smem(k_painn_edge_fwd_sys<NW, NRB, true>, sm, &cur_mu))) return rc;
        spk_launch(k_painn_edge_fwd_sys<NW, NRB, true>, n_mol * parts, NW * 32, sm, st, x, mu, q, phi, geo, rowptr, slot_j, wf, bf,
                                                                              mol_ptr, parts, cap, n_rbf, q_out, mu_out);
    } else {
        if ((rc = set_smem(k_painn_edge_fwd_sys<NW, NRB, false>, sm, &cur_nomu))) return rc;
        spk_launch(k_painn_edge_fwd_sys<NW, NRB, false>, n_mol * parts, NW * 32, sm, st, x, mu, q, phi, geo, rowptr, slot_j, wf,
                                                                               bf, mol_ptr, parts, cap, n_rbf, q_out, mu_out);
    }
    return 0;
}

template <int NW, int NRB>
int launch_bwd_sys(const float* x, const float* mu, const float* g_q, const float* g_mu, const float* phi,
                   const float* dphi, const float* geo, const int* sptr, const int* pos_slot, const int* pos_i,
                   const int* slot_eid, const float* wf, const float* bf, const int* mol_ptr, int n_mol, int parts,
                   int cap, int n_rbf, float* g_x, float* g_mu_in, float* g_rij, int accumulate, cudaStream_t st) {
    constexpr int F = NW * 32;
    const size_t sm = (size_t)cap * 4 * F * 4;
    static size_t cur_mu = 0, cur_nomu = 0;
    int rc;
    if (mu) {
        if ((rc = set_smem(k_painn_edge_bwd_sys<NW, NRB, true>, sm, &cur_mu))) return rc;
        spk_launch(k_painn_edge_bwd_sys<NW, NRB, true>, n_mol * parts, NW * 32, sm, st, 
            x, mu, g_q, g_mu, phi, dphi, geo, sptr, pos_slot, pos_i, slot_eid, wf, bf, mol_ptr, parts, cap, n_rbf, g_x,
            g_mu_in, g_rij, accumulate);
    } else {
        if ((rc = set_smem(k_painn_edge_bwd_sys<NW, NRB, false>, sm, &cur_nomu))) return rc;
        spk_launch(k_painn_edge_bwd_sys<NW, NRB, false>, n_mol * parts, NW * 32, sm, st, 
            x, mu, g_q, g_mu, phi, dphi, geo, sptr, pos_slot, pos_i, slot_eid, wf, bf, mol_ptr, parts, cap, n_rbf, g_x,
            g_mu_in, g_rij, accumulate);
    }
    return 0;
}

// shared-memory capacity (atoms) and CTAs per system for a batch of n_mol systems with n_atoms atoms in total
void plan(int64_t n_atoms, int64_t n_mol, int F, int bytes_per_atom, int* cap, int* parts) {
    int64_t avg = (n_atoms + n_mol - 1) / n_mol;
    int64_t c = (avg * 9 + 7) / 8;                       // 12.5 % head-room over the average system size
    c = (c + 3) & ~3ll;
    const int64_t max_c = (200 * 1024) / bytes_per_atom; // at most ~200 KB of shared memory per CTA
    if (c > max_c) c = max_c;
    if (c < 4) c = 4;
    *cap = (int)c;
    int64_t p = (3ll * spk_num_sms()) / n_mol;           // fill ~3 CTAs per SM without spilling into a second wave
    if (p < 1) p = 1;
    if (p > 4) p = 4;
    *parts = (int)p;
}

}  // namespace

#define DISPATCH_SYS(CALL)                                                     \
    do {                                                                       \
        const int nw_ = F / 32;                                                \
        if (n_rbf <= 20) {                                                     \
            if (nw_ == 1) rc = CALL(1, 20); else if (nw_ == 2) rc = CALL(2, 20); \
            else if (nw_ == 4) rc = CALL(4, 20); else if (nw_ == 8) rc = CALL(8, 20); \
            else return SPK_ERR_UNSUPPORTED;                                   \
        } else {                                                               \
            if (nw_ == 1) rc = CALL(1, 32); else if (nw_ == 2) rc = CALL(2, 32); \
            else if (nw_ == 4) rc = CALL(4, 32); else if (nw_ == 8) rc = CALL(8, 32); \
            else return SPK_ERR_UNSUPPORTED;                                   \
        }                                                                      \
    } while (0)

extern "C" int spk_painn_edge_fwd_sys(const float* x, const float* mu, const float* q, const float* phi,
                                      const float* geo, const int32_t* rowptr, const int32_t* slot_j, const float* wf,
                                      const float* bf, const int32_t* mol_ptr, int64_t n_mol, int64_t n_atoms,
                                      int64_t n_edges, int F, int n_rbf, float* q_out, float* mu_out,
                                      spk_stream_t stream) {
    if (n_atoms < 0 || n_edges < 0 || n_mol <= 0 || F <= 0 || n_rbf <= 0) return SPK_ERR_ARG;
    if ((F & 31) || F > 256 || n_rbf > 32) return SPK_ERR_UNSUPPORTED;
    if (n_atoms == 0) return SPK_OK;
    if (!x || !q || !rowptr || !wf || !bf || !mol_ptr || !q_out || !mu_out) return SPK_ERR_ARG;
    if (n_edges > 0 && (!phi || !geo || !slot_j)) return SPK_ERR_ARG;
    if (mu && mu == mu_out) return SPK_ERR_ARG;
    int cap, parts;
    plan(n_atoms, n_mol, F, (mu ? 2 : 1) * 3 * F * 4, &cap, &parts);
    cudaStream_t st = spk_st(stream);
    int rc = 0;
#define CALL_F(NW, NRB) \
    launch_fwd_sys<NW, NRB>(x, mu, q, phi, geo, rowptr, slot_j, wf, bf, mol_ptr, (int)n_mol, parts, cap, n_rbf, q_out, mu_out, st)
    DISPATCH_SYS(CALL_F);
#undef CALL_F
    if (rc) return rc;
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}

extern "C" int spk_painn_edge_bwd_sys(const float* x, const float* mu, const float* g_q, const float* g_mu,
                                      const float* phi, const float* dphi, const float* geo, const int32_t* sptr,
                                      const int32_t* pos_slot, const int32_t* pos_i, const int32_t* slot_eid,
                                      const float* wf, const float* bf, const int32_t* mol_ptr, int64_t n_mol,
                                      int64_t n_atoms, int64_t n_edges, int F, int n_rbf, float* g_x, float* g_mu_in,
                                      float* g_rij, int accumulate, spk_stream_t stream) {
    if (n_atoms < 0 || n_edges < 0 || n_mol <= 0 || F <= 0 || n_rbf <= 0) return SPK_ERR_ARG;
    if ((F & 31) || F > 256 || n_rbf > 32) return SPK_ERR_UNSUPPORTED;
    if (n_atoms == 0) return SPK_OK;
    if (!x || !g_q || !g_mu || !sptr || !wf || !bf || !mol_ptr || !g_x) return SPK_ERR_ARG;
    if (mu && !g_mu_in) return SPK_ERR_ARG;
    if (n_edges > 0 && (!phi || !dphi || !geo || !pos_slot || !pos_i || !slot_eid || !g_rij)) return SPK_ERR_ARG;
    int cap, parts;
    plan(n_atoms, n_mol, F, 4 * F * 4, &cap, &parts);
    cudaStream_t st = spk_st(stream);
    int rc = 0;
#define CALL_B(NW, NRB)                                                                                               \
    launch_bwd_sys<NW, NRB>(x, mu, g_q, g_mu, phi, dphi, geo, sptr, pos_slot, pos_i, slot_eid, wf, bf, mol_ptr, (int)n_mol, \
                            parts, cap, n_rbf, g_x, g_mu_in, g_rij, accumulate, st)
    DISPATCH_SYS(CALL_B);
#undef CALL_B
    if (rc) return rc;
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}
