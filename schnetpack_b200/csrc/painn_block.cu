// Block-level PaiNN interaction with a CALLER-SUPPLIED filter -- the reference's public block API
//     PaiNNInteraction.forward(q, mu, Wij, dir_ij, idx_i, idx_j, n_atoms)      representation/painn.py:31-67
// where Wij [E,1,3F] has been materialised by the caller (filter_net(phi) * fcut, painn.py:232-236) and dir_ij [E,3] are
// the unit vectors.  The fused model path (painn_tc.cu / painn.cu) never materialises Wij; this pair exists so that code
// which drives the blocks directly (custom representations) keeps working on the same reductions: receiver-grouped
// register accumulation forward, sender-grouped reverse, no atomics, deterministic.
//
//   forward :  q_out[i]  = q[i]  + sum_s Wij[e_s, 0:F]  * x[j_s, 0:F]
//              mu_out[i] = mu[i] + sum_s Wij[e_s, F:2F] * x[j_s, F:2F] (x) dir[e_s] + Wij[e_s, 2F:3F] * x[j_s, 2F:3F] * mu[j_s]
//   reverse :  g_x[j], g_mu_in[j] (= g_mu[j] + ...), g_W[e, 3F], g_dir[e, 3]
// Thread = feature channel (blockDim = F <= 256), CTAs take edge-balanced row ranges.
#include "common.cuh"

namespace {

constexpr int BCH = 8;   // edges per reduction chunk of the reverse kernel

__global__ void __launch_bounds__(256) k_painn_edge_wij_fwd(
    const float* __restrict__ x, const float* __restrict__ mu, const float* __restrict__ q,
    const float* __restrict__ Wij, const float* __restrict__ dir, const int* __restrict__ rowptr,
    const int* __restrict__ slot_j, const int* __restrict__ slot_eid, int n_atoms, int n_edges, int F,
    float* __restrict__ q_out, float* __restrict__ mu_out) {
    SPK_PDL_ENTER();
    const int c = threadIdx.x;
    const int row_lo = spk_block_row_begin(rowptr, n_atoms, n_edges, gridDim.x, blockIdx.x);
    const int row_hi = spk_block_row_begin(rowptr, n_atoms, n_edges, gridDim.x, blockIdx.x + 1);
    for (int i = row_lo; i < row_hi; ++i) {
        float dq = 0.f, d0 = 0.f, d1 = 0.f, d2 = 0.f;
        for (int s = rowptr[i]; s < rowptr[i + 1]; ++s) {
            const int j = slot_j[s];
            const int64_t e = slot_eid[s];
            const float* __restrict__ w = Wij + e * 3 * F + c;
            const float* __restrict__ xj = x + (int64_t)j * 3 * F + c;
            const float* __restrict__ mj = mu + (int64_t)j * 3 * F + c;
            const float ux = dir[e * 3 + 0], uy = dir[e * 3 + 1], uz = dir[e * 3 + 2];
            dq = fmaf(w[0], xj[0], dq);
            const float tb = w[F] * xj[F];
            const float tc = w[2 * F] * xj[2 * F];
            d0 = fmaf(tb, ux, d0);
            d1 = fmaf(tb, uy, d1);
            d2 = fmaf(tb, uz, d2);
            d0 = fmaf(tc, mj[0], d0);
            d1 = fmaf(tc, mj[F], d1);
            d2 = fmaf(tc, mj[2 * F], d2);
        }
        const int64_t o = (int64_t)i * F + c, om = (int64_t)i * 3 * F + c;
        q_out[o] = q[o] + dq;
        mu_out[om] = mu[om] + d0;
        mu_out[om + F] = mu[om + F] + d1;
        mu_out[om + 2 * F] = mu[om + 2 * F] + d2;
    }
}

__global__ void __launch_bounds__(256) k_painn_edge_wij_bwd(
    const float* __restrict__ x, const float* __restrict__ mu, const float* __restrict__ g_q,
    const float* __restrict__ g_mu, const float* __restrict__ Wij, const float* __restrict__ dir,
    const int* __restrict__ sptr, const int* __restrict__ pos_slot, const int* __restrict__ pos_i,
    const int* __restrict__ slot_eid, int n_atoms, int n_edges, int F, float* __restrict__ g_x,
    float* __restrict__ g_mu_in, float* __restrict__ g_W, float* __restrict__ g_dir) {
    SPK_PDL_ENTER();
    __shared__ float s_red[BCH][8][3];
    __shared__ int s_e[BCH];
    const int c = threadIdx.x, lane = c & 31, warp = c >> 5, nw = blockDim.x >> 5;
    const int j_lo = spk_block_row_begin(sptr, n_atoms, n_edges, gridDim.x, blockIdx.x);
    const int j_hi = spk_block_row_begin(sptr, n_atoms, n_edges, gridDim.x, blockIdx.x + 1);
    for (int j = j_lo; j < j_hi; ++j) {
        const float* __restrict__ xj = x + (int64_t)j * 3 * F + c;
        const float* __restrict__ mj = mu + (int64_t)j * 3 * F + c;
        const float xa = xj[0], xb = xj[F], xc = xj[2 * F];
        const float m0 = mj[0], m1 = mj[F], m2 = mj[2 * F];
        float gxa = 0.f, gxb = 0.f, gxc = 0.f, gm0 = 0.f, gm1 = 0.f, gm2 = 0.f;
        const int p_lo = sptr[j], p_hi = sptr[j + 1];
        for (int p0 = p_lo; p0 < p_hi; p0 += BCH) {
            const int n = min(BCH, p_hi - p0);
            __syncthreads();                                     // s_red / s_e of the previous chunk consumed
            for (int t = 0; t < n; ++t) {
                const int s = pos_slot[p0 + t];
                const int i = pos_i[p0 + t];
                const int64_t e = slot_eid[s];
                const float* __restrict__ w = Wij + e * 3 * F + c;
                const float wa = w[0], wb = w[F], wc = w[2 * F];
                const float gq = g_q[(int64_t)i * F + c];
                const float* __restrict__ gmi = g_mu + (int64_t)i * 3 * F + c;
                const float g0 = gmi[0], g1 = gmi[F], g2 = gmi[2 * F];
                const float ux = dir[e * 3 + 0], uy = dir[e * 3 + 1], uz = dir[e * 3 + 2];
                const float gu = g0 * ux + g1 * uy + g2 * uz;
                const float gm = g0 * m0 + g1 * m1 + g2 * m2;
                gxa = fmaf(wa, gq, gxa);
                gxb = fmaf(wb, gu, gxb);
                gxc = fmaf(wc, gm, gxc);
                const float wcx = wc * xc;
                gm0 = fmaf(wcx, g0, gm0);
                gm1 = fmaf(wcx, g1, gm1);
                gm2 = fmaf(wcx, g2, gm2);
                float* __restrict__ gw = g_W + e * 3 * F + c;
                gw[0] = gq * xa;
                gw[F] = gu * xb;
                gw[2 * F] = gm * xc;
                const float wbx = wb * xb;
                const float r0 = spk_warp_sum(g0 * wbx), r1 = spk_warp_sum(g1 * wbx), r2 = spk_warp_sum(g2 * wbx);
                if (lane == 0) {
                    s_red[t][warp][0] = r0;
                    s_red[t][warp][1] = r1;
                    s_red[t][warp][2] = r2;
                }
                if (c == 0) s_e[t] = (int)e;
            }
            __syncthreads();
            if (c < n * 3) {
                const int t = c / 3, d = c - t * 3;
                float acc = 0.f;
                for (int wv = 0; wv < nw; ++wv) acc += s_red[t][wv][d];
                g_dir[(int64_t)s_e[t] * 3 + d] = acc;
            }
        }
        const int64_t o = (int64_t)j * 3 * F + c;
        g_x[o] = gxa;
        g_x[o + F] = gxb;
        g_x[o + 2 * F] = gxc;
        g_mu_in[o] = g_mu[o] + gm0;
        g_mu_in[o + F] = g_mu[o + F] + gm1;
        g_mu_in[o + 2 * F] = g_mu[o + 2 * F] + gm2;
    }
}

inline unsigned block_count(int64_t n_atoms, int64_t n_edges) {
    int64_t nb = (int64_t)spk_num_sms() * 8;
    if (nb > spk_cdiv(n_edges, 8) + 1) nb = spk_cdiv(n_edges, 8) + 1;
    if (nb > n_atoms) nb = n_atoms;
    return (unsigned)(nb < 1 ? 1 : nb);
}

}  // namespace

extern "C" int spk_painn_edge_wij_fwd(const float* x, const float* mu, const float* q, const float* Wij,
                                      const float* dir, const int32_t* rowptr, const int32_t* slot_j,
                                      const int32_t* slot_eid, int64_t n_atoms, int64_t n_edges, int F, float* q_out,
                                      float* mu_out, spk_stream_t stream) {
    if (n_atoms < 0 || n_edges < 0 || F <= 0) return SPK_ERR_ARG;
    if ((F & 31) || F > 256) return SPK_ERR_UNSUPPORTED;
    if (n_atoms == 0) return SPK_OK;
    if (!x || !mu || !q || !rowptr || !q_out || !mu_out) return SPK_ERR_ARG;
    if (n_edges > 0 && (!Wij || !dir || !slot_j || !slot_eid)) return SPK_ERR_ARG;
    if (mu == mu_out) return SPK_ERR_ARG;
    spk_launch(k_painn_edge_wij_fwd, block_count(n_atoms, n_edges), F, 0, spk_st(stream), x, mu, q, Wij, dir, rowptr,
               slot_j, slot_eid, (int)n_atoms, (int)n_edges, F, q_out, mu_out);
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}

extern "C" int spk_painn_edge_wij_bwd(const float* x, const float* mu, const float* g_q, const float* g_mu,
                                      const float* Wij, const float* dir, const int32_t* sptr, const int32_t* pos_slot,
                                      const int32_t* pos_i, const int32_t* slot_eid, int64_t n_atoms, int64_t n_edges,
                                      int F, float* g_x, float* g_mu_in, float* g_W, float* g_dir,
                                      spk_stream_t stream) {
    if (n_atoms < 0 || n_edges < 0 || F <= 0) return SPK_ERR_ARG;
    if ((F & 31) || F > 256) return SPK_ERR_UNSUPPORTED;
    if (n_atoms == 0) return SPK_OK;
    if (!x || !mu || !g_q || !g_mu || !sptr || !g_x || !g_mu_in) return SPK_ERR_ARG;
    if (n_edges > 0 && (!Wij || !dir || !pos_slot || !pos_i || !slot_eid || !g_W || !g_dir)) return SPK_ERR_ARG;
    spk_launch(k_painn_edge_wij_bwd, block_count(n_atoms, n_edges), F, 0, spk_st(stream), x, mu, g_q, g_mu, Wij, dir,
               sptr, pos_slot, pos_i, slot_eid, (int)n_atoms, (int)n_edges, F, g_x, g_mu_in, g_W, g_dir);
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}
