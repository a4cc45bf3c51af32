/*
 * spk_b200.h -- C ABI of the B200-native (sm_100a) message-passing hot path for SchNetPack models.
 *
 * The reference (atomistic-machine-learning/schnetpack @ 4c967da) has NO native interface for this path: every op is
 * an ATen kernel dispatched from Python ("CUDA kernel inventory: empty").  The entry points below
 * are therefore what a ctypes binding inside the reference's own nn.Modules would bind (INTEGRATION.md shows the stub);
 * each one cites the reference Python it replaces (paths relative to /root/reference/src/schnetpack).
 *
 * Conventions:
 *   - every pointer is a DEVICE pointer to contiguous row-major data; floats are fp32, graph indices int32,
 *     user-facing neighbour indices (idx_i/idx_j/idx_m/Z) int64 exactly as the reference's tensors;
 *   - the caller owns every buffer (outputs and workspaces included); the library never allocates, never
 *     synchronises, keeps no global mutable state (it reads no environment
 *     variables; per-device launch attributes are cached) and only enqueues work on the given stream (graph-capturable);
 *   - return value: 0 = ok, SPK_ERR_ARG (-1) bad argument, SPK_ERR_UNSUPPORTED (-2) shape outside the compiled
 *     templates, -(1000 + cudaError_t) if a launch failed.  Nothing throws.
 *   - F = n_atom_basis (multiple of 32, <= 256), n_rbf <= 32, KP = SPK_KP(n_rbf) = n_rbf rounded up to 4.
 */
#ifndef SPK_B200_H
#define SPK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* spk_stream_t; /* cudaStream_t */

#define SPK_OK 0
#define SPK_ERR_ARG (-1)
#define SPK_ERR_UNSUPPORTED (-2)

#define SPK_ACT_NONE 0
#define SPK_ACT_SILU 1 /* torch.nn.functional.silu            (PaiNN, Atomwise)  */
#define SPK_ACT_SSP 2  /* nn/activations.py:9-22 shifted_softplus (SchNet)        */
#define SPK_ACT_GIVEN 3      /* a_act only: a_pre already holds act'(pre) (saved by the forward layer, see below)  */
#define SPK_SAVE_DERIV 0x10  /* OR-ed into `act`: y_pre receives act'(pre) instead of the pre-activation, so the input-
                              * gradient layer multiplies by it without re-evaluating exp() (a_act = SPK_ACT_GIVEN)  */

#define SPK_RBF_GAUSSIAN 0 /* nn/radial.py:11-48  */
#define SPK_RBF_BESSEL 1   /* nn/radial.py:82-110 */

#define SPK_GEO_STRIDE 8 /* floats per edge in the geometry record: ux uy uz d fc dfc/dd 1/d 0 */
#define SPK_NRB(n_rbf) ((n_rbf) <= 20 ? 20 : 32)          /* radial-basis capacity of the compiled edge kernels */

int spk_version(void);

/* ---------------------------------------------------------------------------------------------------------------
 * Graph structure.  Replaces the implicit ordering assumptions of index_select / index_add_ in
 * nn/scatter.py:26-34, representation/painn.py:55-62, representation/schnet.py:65-67.
 *   rowptr[N+1], slot_j[E], slot_eid[E] : edges grouped by receiver idx_i (CSR).  Slot s of row i holds the sender
 *       slot_j[s] and the position slot_eid[s] of that edge in the caller's idx_i/idx_j arrays.  If idx_i is already
 *       sorted (reference collate order) slots are the identity permutation; otherwise a stable grouping is built.
 *   sptr[N+1], pos_slot[E], pos_i[E]    : the same edges grouped by sender idx_j (ascending slot inside a group).
 *   status[4] (device int32): [0]=1 if idx_i was sorted, [1]=number of out-of-range indices (must be 0; otherwise
 *       rowptr and sptr are zeroed = an empty graph, so no consumer gathers through unwritten slots),
 *       [2]=max receiver degree, [3]=max sender degree.
 *   workspace: spk_graph_workspace_bytes(N, E) bytes.
 * ------------------------------------------------------------------------------------------------------------- */
size_t spk_graph_workspace_bytes(int64_t n_atoms, int64_t n_edges);
int spk_graph_build(const int64_t* idx_i, const int64_t* idx_j, int64_t n_atoms, int64_t n_edges, int32_t* rowptr,
                    int32_t* slot_j, int32_t* slot_eid, int32_t* sptr, int32_t* pos_slot, int32_t* pos_i,
                    int32_t* status, void* workspace, size_t workspace_bytes, spk_stream_t stream);

/* The same views over the ACTIVE edges only: an edge is kept iff |r_ij[e]| < cutoff, i.e. iff the cosine cutoff does not
 * zero its message (nn/cutoff.py:30-32) -- padded neighbour lists (slots at distance >= cutoff) then cost nothing in the
 * edge kernels.  rowptr[n_atoms] (device) is the number of kept edges; slot arrays are filled for [0, rowptr[n_atoms]) and
 * slot_eid still points into the caller's full edge arrays.  r_ij [E,3] fp32. */
int spk_graph_build_active(const int64_t* idx_i, const int64_t* idx_j, const float* r_ij, float cutoff, int64_t n_atoms,
                           int64_t n_edges, int32_t* rowptr, int32_t* slot_j, int32_t* slot_eid, int32_t* sptr,
                           int32_t* pos_slot, int32_t* pos_i, int32_t* status, void* workspace, size_t workspace_bytes,
                           spk_stream_t stream);

/* mol_ptr[B+1] from the sorted system index idx_m (atomistic/atomwise.py:79-81 uses idx_m with index_add). */
int spk_segment_ptr(const int64_t* idx_m, int64_t n_atoms, int64_t n_mol, int32_t* mol_ptr, spk_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Geometry.  atomistic/distances.py:14-26 (PairwiseDistances), representation/painn.py:227-230,
 * representation/schnet.py:156-158, nn/radial.py, nn/cutoff.py:14-33.
 * ------------------------------------------------------------------------------------------------------------- */
/* Rij[e] = R[idx_j[e]] - R[idx_i[e]] + offsets[e]; an edge with an index outside [0, n_atoms) yields NaN (no OOB read) */
int spk_pairwise_fwd(const float* R, const int64_t* idx_i, const int64_t* idx_j, const float* offsets, int64_t n_atoms,
                     int64_t n_edges, float* r_ij, spk_stream_t stream);
/* dE/dR[a] = sum_{e: j(e)=a} g[e] - sum_{e: i(e)=a} g[e]   (deterministic, no atomics); out = sign * that */
int spk_pairwise_bwd(const float* g_rij, const int32_t* rowptr, const int32_t* slot_eid, const int32_t* sptr,
                     const int32_t* pos_slot, int64_t n_atoms, float sign, float* g_R, spk_stream_t stream);
/* per CSR slot s (edge slot_eid[s]): phi[s,0:n_rbf] radial basis (zero padded to KP), dphi = d phi/dd,
 * geo[s] = (ux, uy, uz, d, fc, dfc/dd, 1/d, 0).  rbf_p0/p1 = offsets/widths (gaussian) or freqs/NULL (bessel).
 * n_active (nullable, device): number of slots that exist (rowptr[n_atoms] of spk_graph_build_active); slots past it are
 * left untouched. */
int spk_edge_geometry(const float* r_ij, const int32_t* slot_eid, int64_t n_edges, int rbf_kind, int n_rbf,
                      const float* rbf_p0, const float* rbf_p1, float cutoff, const int32_t* n_active, float* phi,
                      float* dphi, float* geo, spk_stream_t stream);
/* standalone radial basis / cutoff / activation (nn.GaussianRBF, nn.BesselRBF, nn.CosineCutoff, shifted_softplus
 * forward + derivative, used by the nn.* module mirrors).  d: [n]; out: [n, n_rbf]; dout nullable */
int spk_rbf_fwd(const float* d, int64_t n, int rbf_kind, int n_rbf, const float* rbf_p0, const float* rbf_p1,
                float* out, float* dout, spk_stream_t stream);
int spk_cosine_cutoff_fwd(const float* d, int64_t n, float cutoff, float* out, float* dout, spk_stream_t stream);
int spk_act_fwd(const float* x, int64_t n, int act, float* y, float* dy, spk_stream_t stream);
/* out[a, :] = table[Z[a], :]   (nn.Embedding in representation/painn.py:239, schnet.py:161); Z outside [0, n_rows)
 * (nn.Embedding raises) gives a NaN row */
int spk_embedding(const float* table, const int64_t* Z, int64_t n_atoms, int F, int n_rows, float* out,
                  spk_stream_t stream);
/* out[idx[r], :] += x[r, :] for sorted-or-not idx via a row-pointer (deterministic): generic nn/scatter.py:7-34.
 * rowptr/slot_eid from spk_graph_build-style grouping of idx.  out[n_out, C] is fully overwritten. */
int spk_segment_sum(const float* x, const int32_t* rowptr, const int32_t* slot_eid, int64_t n_out, int C, float* out,
                    spk_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Dense layers.  nn/base.py:52-55 (Dense.forward = activation(F.linear(x, W, b))).
 *   Y[M,N] = act( A[M,K] * B[K,N] + bias[N] ) + addend[M,N]
 * B is the TRANSPOSED torch weight (W^T, [in,out] row-major) for a forward layer, or the weight itself for the
 * input-gradient of a layer ( gX = (gY .* act'(pre)) W ).  If a_pre != NULL the A operand is multiplied on load by
 * act'(a_pre) with activation a_act (the backward prologue).  y_pre (nullable) receives the pre-activation.
 * ------------------------------------------------------------------------------------------------------------- */
int spk_dense(const float* A, int64_t M, int K, int64_t lda, const float* a_pre, int a_act, const float* B, int N,
              const float* bias, int act, const float* addend, int64_t ld_add, float* Y, int64_t ldy, float* y_pre,
              spk_stream_t stream);

/* Contract of every PACKED operand (spk_tc_pack_weight, spk_painn_pack_filter, spk_schnet_pack_filter): the packed buffer is
 * a static operand.  The tensor-core kernels fetch it by TMA BEFORE their programmatic-launch dependency wait, so it must be
 * written only by these pack entry points (whose kernels never release their dependents early) or be complete before the
 * preceding kernels of the stream were launched; do not overwrite it with other kernels while evaluations are in flight. */
/* Same layer on the tcgen05 tensor cores with 3xTF32 error compensation (fp32-grade results: split accumulators, K-tile
 * draining; see csrc/gemm_tc.cu).  The weight the A rows are contracted with -- the torch weight [N,K] itself for a
 * forward layer, its transpose for the input-gradient -- is packed once by spk_tc_pack_weight into per-(64-row, 16-column)
 * operand tiles [tf32_round(W) | W - tf32_round(W)] laid out exactly as the kernel's shared-memory UMMA operand, so one TMA
 * bulk copy stages a tile.  Requirements: K, N, lda, ldy, ld_add multiples of 4, 16 B-aligned pointers, and at most one of
 * (a_act, act) active; otherwise SPK_ERR_UNSUPPORTED is returned and the caller uses spk_dense. */
size_t spk_tc_packed_floats(int N, int K);
int spk_tc_pack_weight(const float* W, int N, int K, float* packed, spk_stream_t stream);
int spk_dense_tc(const float* A, int64_t M, int K, int64_t lda, const float* a_pre, int a_act, const float* W_packed,
                 int N, const float* bias, int act, const float* addend, int64_t ld_add, float* Y, int64_t ldy,
                 float* y_pre, spk_stream_t stream);

/* Two Dense layers as ONE launch (csrc/mlp2_tc.cu): H = act(A W0^T + b0) [M,128] stays in shared memory as the tensor-core
 * operand of Y = H W1^T + b1 [+ addend] [M,N2]; h_deriv (or NULL) receives act'(pre) of the hidden layer [M,128] (what
 * spk_dense_tc returns with SPK_SAVE_DERIV) for the reverse pass.  Replaces nn.Sequential(Dense(K1, 128, activation),
 * Dense(128, N2)) of /root/reference/src/schnetpack/representation/painn.py:39-44, 84-91 (PaiNN's context networks).
 * W0_packed / W1_packed = the 128-column-tile section of spk_tc_pack_weight's output (offset
 * spk_tc_packed_floats_tn(N, K, 64) floats) for W0 [128,K1] and W1 [N2,128].  Requirements: hidden width 128, N2 % 128 == 0,
 * K1 % 4 == 0, leading dimensions % 4 == 0, 16 B-aligned pointers; otherwise SPK_ERR_UNSUPPORTED (run two spk_dense_tc). */
int spk_mlp2_tc(const float* A, int64_t M, int K1, int64_t lda, const float* W0_packed, const float* b0, int act,
                const float* W1_packed, int N2, const float* b1, const float* addend, int64_t ld_add, float* Y, int64_t ldy,
                float* h_deriv, spk_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * PaiNN.  representation/painn.py:31-67 (PaiNNInteraction.forward) and :92-117 (PaiNNMixing.forward).
 * ------------------------------------------------------------------------------------------------------------- */
/* Fused per-edge kernel of one interaction block (painn.py:55-65 + the filter of :232-236 recomputed on the fly):
 *   W_s   = fc_s * (bf + phi_s Wf^T)                       [3F]   (never materialised)
 *   q_out[i]  = q[i]  + sum_{s in row i} W_s[0:F]   * x[j_s, 0:F]
 *   mu_out[i] = mu[i] + sum_{s in row i} W_s[F:2F]  * x[j_s, F:2F] (x) u_s + W_s[2F:3F] * x[j_s, 2F:3F] * mu[j_s]
 * x = interatomic_context_net(q) [N,3F]; mu [N,3,F] may be NULL (== zeros, first block); wf [3F, n_rbf] and bf [3F]
 * are this block's rows of filter_net.{weight,bias}.  q_out may alias q; mu_out must not alias mu. */
int spk_painn_edge_fwd(const float* x, const float* mu, const float* q, const float* phi, const float* geo,
                       const int32_t* rowptr, const int32_t* slot_j, const float* wf, const float* bf, int64_t n_atoms,
                       int64_t n_edges, int F, int n_rbf, float* q_out, float* mu_out, spk_stream_t stream);
/* Reverse of the above grouped by SENDER (no atomics): given g_q = dE/dq_out [N,F], g_mu = dE/dmu_out [N,3,F]:
 *   g_x[j]     = sum_{edges with sender j} ...                                   [N,3F]  (overwritten)
 *   g_mu_in[j] = g_mu[j] + sum ...                                               [N,3,F] (nullable when mu == NULL)
 *   g_rij[eid] (+)= dE/dr_ij through d (phi, fc) and u                           [E,3]   (accumulate != 0 -> +=)
 * The residual dE/dq_in = g_q is the caller's (identity).  */
int spk_painn_edge_bwd(const float* x, const float* mu, const float* g_q, const float* g_mu, const float* phi,
                       const float* dphi, const float* geo, const int32_t* sptr, const int32_t* pos_slot,
                       const int32_t* pos_i, const int32_t* slot_eid, const float* wf, const float* bf,
                       int64_t n_atoms, int64_t n_edges, int F, int n_rbf, float* g_x, float* g_mu_in, float* g_rij,
                       int accumulate, spk_stream_t stream);
/* Tensor-core variant of spk_painn_edge_fwd (csrc/painn_tc.cu): the filter W = fc * (phi . w^T + b) of a chunk of edges is
 * one 3xTF32 tcgen05 GEMM per filter third with the channels on the TMEM lanes, so the gather/accumulate threads read their
 * channel's filter values with tcgen05.ld instead of recomputing them from warp-broadcast shared-memory loads (the
 * streaming kernel is bound by the load/store unit).  wf_packed = spk_painn_pack_filter(wf, bf) (hi/lo operand tiles,
 * spk_painn_filter_packed_floats() floats).  Same results as spk_painn_edge_fwd to fp32 rounding.  Returns
 * SPK_ERR_UNSUPPORTED unless F == 128, n_rbf <= 31 and n_edges > 0 (the caller then uses spk_painn_edge_fwd). */
size_t spk_painn_filter_packed_floats(void);
int spk_painn_pack_filter(const float* wf, const float* bf, int F, int n_rbf, float* packed, spk_stream_t stream);
int spk_painn_edge_fwd_tc(const float* x, const float* mu, const float* q, const float* phi, const float* geo,
                          const int32_t* rowptr, const int32_t* slot_j, const float* wf_packed, int64_t n_atoms,
                          int64_t n_edges, int F, int n_rbf, float* q_out, float* mu_out, spk_stream_t stream);
/* Tensor-core variant of spk_painn_edge_bwd: one operand of 64 rows per chunk of 32 edges, rows 0..31 = fc [phi | 1],
 * rows 32..63 = dfc [phi | 1] + fc [dphi | 0], so the same 27 MMAs produce W and dW/dd for the 8 edges of each of the 4
 * sender groups.  Same outputs as spk_painn_edge_bwd to fp32 rounding; SPK_ERR_UNSUPPORTED under the conditions of
 * spk_painn_edge_fwd_tc. */
int spk_painn_edge_bwd_tc(const float* x, const float* mu, const float* g_q, const float* g_mu, const float* phi,
                          const float* dphi, const float* geo, const int32_t* sptr, const int32_t* pos_slot,
                          const int32_t* pos_i, const int32_t* slot_eid, const float* wf_packed, int64_t n_atoms,
                          int64_t n_edges, int F, int n_rbf, float* g_x, float* g_mu_in, float* g_rij, int accumulate,
                          spk_stream_t stream);
/* Block-level interaction with a caller-supplied, materialised filter -- PaiNNInteraction.forward(q, mu, Wij, dir_ij, idx_i,
 * idx_j, n_atoms), painn.py:31-67 (csrc/painn_block.cu).  Wij [E,3F] and dir [E,3] are in the CALLER's edge order (looked
 * up through slot_eid); x = interatomic_context_net(q) [N,3F]; mu [N,3,F] required.  Reverse: g_x [N,3F], g_mu_in [N,3,F]
 * (= g_mu + ...), g_W [E,3F], g_dir [E,3] in the caller's edge order, all overwritten. */
int spk_painn_edge_wij_fwd(const float* x, const float* mu, const float* q, const float* Wij, const float* dir,
                           const int32_t* rowptr, const int32_t* slot_j, const int32_t* slot_eid, int64_t n_atoms,
                           int64_t n_edges, int F, float* q_out, float* mu_out, spk_stream_t stream);
int spk_painn_edge_wij_bwd(const float* x, const float* mu, const float* g_q, const float* g_mu, const float* Wij,
                           const float* dir, const int32_t* sptr, const int32_t* pos_slot, const int32_t* pos_i,
                           const int32_t* slot_eid, int64_t n_atoms, int64_t n_edges, int F, float* g_x, float* g_mu_in,
                           float* g_W, float* g_dir, spk_stream_t stream);
/* painn.py:104-107: ctx[a] = [ q[a] | sqrt(sum_d V[a,d]^2 + eps) ], VW = mu_channel_mix(mu) [N,3,2F] */
int spk_painn_mix_ctx(const float* q, const float* VW, int64_t n_atoms, int F, float eps, float* ctx,
                      spk_stream_t stream);
/* painn.py:110-116: q_out = q + s1 + s3 * sum_d V_d W_d ; mu_out[d] = mu[d] + s2 * W_d ; s [N,3F] */
int spk_painn_mix_update(const float* q, const float* mu, const float* s, const float* VW, int64_t n_atoms, int F,
                         float* q_out, float* mu_out, spk_stream_t stream);
/* reverse of mix_update: g_s [N,3F]; g_VW [N,3,2F] (without the norm path) */
int spk_painn_mix_update_bwd(const float* g_q, const float* g_mu, const float* s, const float* VW, int64_t n_atoms,
                             int F, float* g_s, float* g_VW, spk_stream_t stream);
/* reverse of mix_ctx: g_q_out = g_q + g_ctx[:, :F]; g_VW[:, d, :F] += g_ctx[:, F:] * V_d / n */
int spk_painn_mix_ctx_bwd(const float* g_ctx, const float* g_q, const float* VW, int64_t n_atoms, int F, float eps,
                          float* g_q_out, float* g_VW, spk_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Persistent per-atom stage (csrc/atom_chain.cu): all per-atom work between two edge kernels -- PaiNNMixing
 * (painn.py:103-116) followed by the next block's interatomic_context_net (painn.py:54), or their reverses -- as ONE
 * launch.  The caller describes the stage as up to SPK_CHAIN_MAX_STEPS steps executed in order for every 128-atom tile
 * (step k of a tile starts when step k-1 of the SAME tile is complete; tiles are independent because every step is
 * row-local):
 *   SPK_CHAIN_GEMM   Y[rows, N] = act( (A[rows, K] .* a_pre) * W^T + bias ) + addend over rows_per_atom * n_atoms rows
 *                    (the spk_dense_tc operation; a_pre != NULL multiplies A by the saved act'(pre) = SPK_ACT_GIVEN; with an
 *                    activation, y_pre != NULL receives act'(pre)).  Wp = spk_tc_pack_weight output advanced to the
 *                    128-column tiles: Wp + spk_tc_packed_floats_tn(N, K, 64).  Requires N % 128 == 0, K % 16 == 0.
 *   SPK_CHAIN_MIX_CTX / MIX_UPDATE / MIX_UPDATE_BWD / MIX_CTX_BWD   the elementwise glue of spk_painn_mix_* with operands
 *                    (g0, g1, g2, g3) -> (o0, o1) as listed in csrc/atom_chain.cu.
 * workspace: spk_atom_chain_workspace_ints(n_steps, n_atoms) int32, zero before the FIRST use (the kernel leaves it zero);
 * one workspace per concurrently running stream.
 * ------------------------------------------------------------------------------------------------------------- */
#define SPK_CHAIN_MAX_STEPS 8
#define SPK_CHAIN_GEMM 0
#define SPK_CHAIN_MIX_CTX 1        /* g0 = q [N,F], g1 = VW [N,3,2F]                      -> o0 = ctx [N,2F]                  */
#define SPK_CHAIN_MIX_UPDATE 2     /* g0 = q, g1 = VW, g2 = mu [N,3,F], g3 = s [N,3F]     -> o0 = q', o1 = mu'               */
#define SPK_CHAIN_MIX_UPDATE_BWD 3 /* g0 = g_q, g1 = VW, g2 = g_mu, g3 = s               -> o0 = g_s [N,3F], o1 = g_VW      */
#define SPK_CHAIN_MIX_CTX_BWD 4    /* g0 = g_ctx [N,2F], g1 = VW, g2 = g_q               -> o0 = g_q', o1 = g_VW (V part +=) */
typedef struct {
    int32_t kind, rows_per_atom, K, N, act, F;
    float eps;
    int32_t reserved;
    int64_t lda, ldy, ld_add;
    const float *A, *a_pre, *Wp, *bias, *addend;
    float *Y, *y_pre;
    const float *g0, *g1, *g2, *g3;
    float *o0, *o1;
} spk_chain_step_t;
size_t spk_tc_packed_floats_tn(int N, int K, int tile_n); /* floats of the tile_n-wide (64 | 128) packing of W [N,K] */
size_t spk_atom_chain_workspace_ints(int n_steps, int64_t n_atoms);
#define SPK_CHAIN_FLAG_NFOLD 1 /* 2 MMAs per k-step: [W_hi ; W_lo] as one 256-row operand (see csrc/atom_chain.cu) */
int spk_atom_chain(const spk_chain_step_t* steps /* host array */, int n_steps, int64_t n_atoms, int32_t* workspace,
                   size_t workspace_ints, int flags, spk_stream_t stream);
/* development aid: the same launch writing 8 int64 stamps per work item (claimed, dependencies satisfied, K-loop done,
 * published in ns of %globaltimer; CTA; step; atom tile; unused) to trace[8 * items] (tools/chain_trace.py) */
int spk_atom_chain_debug(const spk_chain_step_t* steps, int n_steps, int64_t n_atoms, int32_t* workspace,
                         size_t workspace_ints, int flags, long long* trace, spk_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * SchNet.  representation/schnet.py:41-70 (SchNetInteraction.forward).
 * ------------------------------------------------------------------------------------------------------------- */
/* continuous-filter convolution (schnet.py:62-67): m[i] = sum_{s in row i} h[j_s] * Wraw[s] * fc_s ; Wraw [E,F] is
 * the filter-network output per CSR slot. */
int spk_cfconv_fwd(const float* h, const float* w_raw, const float* geo, const int32_t* rowptr, const int32_t* slot_j,
                   int64_t n_atoms, int64_t n_edges, int F, float* m, spk_stream_t stream);
/* The whole interaction-block forward edge pipeline fused, filter network on tcgen05 (csrc/schnet_tc.cu):
 *   m[i] = sum_{s in row i} h[j_s] * ( W1 act(W0 phi_s + b0) + b1 ) * fc_s          schnet.py:61-67
 * filter_packed = spk_schnet_pack_filter(filter_network.0.{weight,bias}, filter_network.1.weight) (hi|lo operand tiles,
 * spk_schnet_filter_packed_floats() floats); b1 = filter_network.1.bias; act = SPK_ACT_SSP | SPK_ACT_SILU | SPK_ACT_NONE.
 * rowptr may come from spk_graph_build_active (rowptr[n_atoms] slots exist; n_edges is only the capacity of the arrays).
 * F == n_filters == 128 and n_rbf <= 31, otherwise SPK_ERR_UNSUPPORTED (the caller runs the materialised pipeline). */
size_t spk_schnet_filter_packed_floats(void);
int spk_schnet_pack_filter(const float* w0, const float* b0, const float* w1, int F, int n_rbf, float* packed,
                           spk_stream_t stream);
int spk_schnet_cfconv_fwd_tc(const float* h, const float* phi, const float* geo, const int32_t* rowptr,
                             const int32_t* slot_j, const float* filter_packed, const float* b1, int act, int64_t n_atoms,
                             int64_t n_edges, int F, int n_rbf, float* m, spk_stream_t stream);
/* reverse grouped by sender: g_h[j] = sum W fc g_m[i];  g_wraw[s] = h[j] g_m[i] fc_s;
 * g_fc[s] = sum_c h[j,c] g_m[i,c] Wraw[s,c] */
int spk_cfconv_bwd(const float* h, const float* w_raw, const float* geo, const float* g_m, const int32_t* sptr,
                   const int32_t* pos_slot, const int32_t* pos_i, int64_t n_atoms, int64_t n_edges, int F, float* g_h,
                   float* g_wraw, float* g_fc, spk_stream_t stream);
/* g_rij[eid] (+)= ( sum_k g_phi[s,k] dphi[s,k] + g_fc[s] dfc_s ) * u_s      (d-only dependence of SchNet) */
int spk_radial_bwd(const float* g_phi, const float* g_fc, const float* dphi, const float* geo, const int32_t* slot_eid,
                   int64_t n_edges, int n_rbf, float* g_rij, int accumulate, spk_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Atomwise head.  atomistic/atomwise.py:69-88: y = outnet(q); E[m] = sum_{a in m} y[a].
 * ------------------------------------------------------------------------------------------------------------- */
/* y[a] = hid[a,:] . w1 + b1 ; energy[m] = sum_{a in mol m} y[a]   (hid = silu(Dense_0(q)) [N,H]); y nullable */
int spk_atomwise_out(const float* hid, const float* w1, const float* b1, const int32_t* mol_ptr, int64_t n_atoms,
                     int64_t n_mol, int H, float* y, float* energy, spk_stream_t stream);
/* g_hid[a,:] = g_energy[mol(a)] * w1   (g_energy nullable == ones) */
int spk_atomwise_out_bwd(const float* g_energy, const int64_t* idx_m, const float* w1, int64_t n_atoms, int H,
                         float* g_hid, spk_stream_t stream);

/* elementwise helpers for residual streams: out = a + b (b nullable -> copy), may alias */
int spk_add(const float* a, const float* b, int64_t n, float* out, spk_stream_t stream);

/* ---- "next" row f1: device-resident neighbour list (replaces the host round trip of md/neighborlist_md.py:129,213-232 and
 * the ASE / matscipy / vesin / torch builders of transform/neighborlist.py:213-286,428-553) ---------------------------------
 * Linked-cell search per system of a collated batch.  Listed: every (i, j, S) with |R[j] - R[i] + S @ cell| < cutoff, S
 * integer with S_a = 0 on non-periodic axes, except (i == j, S == 0); offsets = S @ cell (fp32); rows sorted by idx_i, a row
 * in traversal order (deterministic; compare after the canonical (i, j, S) sort as the reference's tests do).
 *   R [n_atoms,3] fp32, cell [n_sys,3,3] fp32 (rows = lattice vectors; ignored where pbc is all zero), pbc [n_sys,3] uint8,
 *   sys_ptr [n_sys+1] int32 = first atom of every system (spk_segment_ptr).
 *   idx_i, idx_j [capacity] int64, offsets [capacity,3] fp32, shifts [capacity,3] int32 (nullable).
 *   n_pairs [2] int64 on the DEVICE: [0] = pairs found, [1] = 1 if they exceeded `capacity` (then only the first
 *   `capacity` were written).  pad != 0 spreads the unused capacity over the rows as self pairs (i, i) at distance
 *   2 * cutoff appended to each row -- outside the cutoff, so they contribute nothing to energies or forces, idx_i stays
 *   sorted and no row grows long -- which lets a caller run a fixed-size edge list without ever reading n_pairs on the host
 *   (CUDA-graph capturable MD step).  Without pad the pairs are contiguous in [0, n_pairs).  capacity == 0 only counts.
 * Enqueue-only on `stream`; workspace of spk_neighbor_list_workspace_bytes(n_atoms, n_sys) bytes. */
size_t spk_neighbor_list_workspace_bytes(int64_t n_atoms, int64_t n_sys);
int spk_neighbor_list(const float* R, const float* cell, const uint8_t* pbc, const int32_t* sys_ptr, int64_t n_atoms,
                      int64_t n_sys, float cutoff, int64_t capacity, int pad, int64_t* idx_i, int64_t* idx_j,
                      float* offsets, int32_t* shifts, int64_t* n_pairs, void* workspace, size_t workspace_bytes,
                      spk_stream_t stream);

/* ---- row e: halo exchange of ghost-atom rows over NVLink peer memory (csrc/halo.cu).  peer_base[world] (device array) holds
 * every rank's symmetric buffer address as mapped into THIS process (CUDA IPC / torch symmetric memory); the caller has
 * enqueued a cross-rank barrier after the owners wrote their rows.
 *   spk_halo_pull      ghost[k, :] = peer[ghost_rank[k]][ghost_row[k], :]          rows of row_floats floats
 *   spk_halo_pull_add  g_rows[rows[k], :] += sum_{e in [entry_ptr[k], entry_ptr[k+1])} peer[entry_rank[e]][entry_pos[e], :]
 *                      (entries of a row in ascending peer order: deterministic reverse halo, no atomics) */
int spk_halo_pull(float* ghost, const uint64_t* peer_base, const int32_t* ghost_rank, const int32_t* ghost_row,
                  int64_t n_ghost, int row_floats, spk_stream_t stream);
int spk_halo_pull_add(float* g_rows, const uint64_t* peer_base, const int32_t* rows, const int32_t* entry_ptr,
                      const int32_t* entry_rank, const int32_t* entry_pos, int64_t n_listed, int row_floats,
                      spk_stream_t stream);

/* ---- "next" row f2: velocity-Verlet update of the device-resident MD state (md/integrators.py:59-70,97-110;
 * unit handling of md/calculators/base_calculator.py:96,120-152).  momenta [N,3] += 1/2 dt * forces * force_conversion; with
 * drift != 0 also positions [N,3] += dt * momenta / masses [N] and, if given, model_positions = positions *
 * position_conversion (the array the next force evaluation reads).  All in place, enqueue-only. */
int spk_md_velocity_verlet(float* momenta, float* positions, float* model_positions, const float* forces,
                           const float* masses, int64_t n_atoms, float dt, float force_conversion,
                           float position_conversion, int drift, spk_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SPK_B200_H */
