// Graph structure build: receiver-grouped (CSR) and sender-grouped views of the reference's (idx_i, idx_j) edge
// list, fully on device, no host synchronisation, deterministic output (groups ordered by edge position).
// Replaces the ordering that index_select/index_add_ leave implicit (nn/scatter.py:26-34, painn.py:55-62).
#include "common.cuh"

namespace {

// sorted_hint = 0 forces the general (cursor + stable rank) fill: used when edges are filtered, because the identity
// mapping slot == edge position of the sorted fast path no longer holds
__global__ void k_init_status(int* __restrict__ status, int sorted_hint) {
    SPK_PDL_ENTER();
    if (threadIdx.x < 4) status[threadIdx.x] = (threadIdx.x == 0) ? sorted_hint : 0;
}

// edge filter of spk_graph_build_active: keep the edges the cosine cutoff does not zero (nn/cutoff.py:30-32: d < rc), with
// the distance expression of k_edge_geometry
__device__ __forceinline__ bool edge_kept(const float* __restrict__ r_ij, int64_t e, float rc) {
    if (!r_ij) return true;
    const float x = r_ij[e * 3 + 0], y = r_ij[e * 3 + 1], z = r_ij[e * 3 + 2];
    return sqrtf(x * x + y * y + z * z) < rc;
}

__global__ void k_count(const int64_t* __restrict__ idx_i, const int64_t* __restrict__ idx_j, int64_t n_atoms,
                        int64_t n_edges, const float* __restrict__ r_ij, float rc, int* __restrict__ deg_i,
                        int* __restrict__ deg_j, int* __restrict__ status) {
    SPK_PDL_ENTER();
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    if (!edge_kept(r_ij, e, rc)) return;
    int64_t i = idx_i[e], j = idx_j[e];
    if (i < 0 || i >= n_atoms || j < 0 || j >= n_atoms) {
        atomicAdd(&status[1], 1);
        return;
    }
    atomicAdd(&deg_i[i], 1);
    atomicAdd(&deg_j[j], 1);
    if (e > 0 && idx_i[e - 1] > i) status[0] = 0;  // benign race: every writer stores 0
}

// exclusive scan of deg[0..n) -> ptr[0..n]; one block per array (blockIdx.x selects), 1024 threads.
__global__ void __launch_bounds__(1024) k_scan2(const int* __restrict__ deg_a, int* __restrict__ ptr_a,
                                                const int* __restrict__ deg_b, int* __restrict__ ptr_b, int n,
                                                int* __restrict__ status) {
    SPK_PDL_ENTER();
    const int* deg = blockIdx.x == 0 ? deg_a : deg_b;
    int* ptr = blockIdx.x == 0 ? ptr_a : ptr_b;
    __shared__ int s_sum[1024];
    __shared__ int s_max[32];
    int t = threadIdx.x;
    int per = (n + 1023) / 1024;
    int lo = min(n, t * per), hi = min(n, lo + per);
    int sum = 0, mx = 0;
    for (int k = lo; k < hi; ++k) {
        int d = deg[k];
        sum += d;
        mx = max(mx, d);
    }
    s_sum[t] = sum;
    __syncthreads();
    // inclusive Hillis-Steele scan over 1024 partial sums
    for (int off = 1; off < 1024; off <<= 1) {
        int v = (t >= off) ? s_sum[t - off] : 0;
        __syncthreads();
        s_sum[t] += v;
        __syncthreads();
    }
    int run = s_sum[t] - sum;  // exclusive prefix of this thread's chunk
    for (int k = lo; k < hi; ++k) {
        ptr[k] = run;
        run += deg[k];
    }
    if (t == 1023) ptr[n] = s_sum[1023];
    // max degree
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((t & 31) == 0) s_max[t >> 5] = mx;
    __syncthreads();
    if (t < 32) {
        mx = s_max[t];
        for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (t == 0) status[2 + blockIdx.x] = mx;
    }
}

// CSR fill.  Sorted input: identity.  Otherwise atomic cursors into tmp (ordered later by k_rank_rows).
__global__ void k_fill_csr(const int64_t* __restrict__ idx_i, const int64_t* __restrict__ idx_j, int64_t n_edges,
                           const float* __restrict__ r_ij, float rc, const int* __restrict__ rowptr,
                           int* __restrict__ cursor, const int* __restrict__ status, int* __restrict__ slot_j,
                           int* __restrict__ slot_eid, int* __restrict__ tmp_eid) {
    SPK_PDL_ENTER();
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    if (status[1] != 0) return;
    if (!edge_kept(r_ij, e, rc)) return;
    if (status[0] != 0) {
        slot_j[e] = (int)idx_j[e];
        slot_eid[e] = (int)e;
    } else {
        int i = (int)idx_i[e];
        int p = rowptr[i] + atomicAdd(&cursor[i], 1);
        tmp_eid[p] = (int)e;
    }
}

// Order every row of (key) ascending by key and carry up to one payload computed from the key.
// mode 0: CSR rows of unsorted input: key = eid -> slot_eid, slot_j = idx_j[eid]   (skipped when input was sorted)
// mode 1: sender rows: key = slot -> pos_slot, pos_i = idx_i[slot_eid[slot]]
__global__ void k_rank_rows(const int* __restrict__ ptr, int n_rows, const int* __restrict__ tmp_key,
                            const int64_t* __restrict__ idx_src, const int* __restrict__ slot_eid_in, int mode,
                            const int* __restrict__ status, int* __restrict__ out_key, int* __restrict__ out_val) {
    SPK_PDL_ENTER();
    if (status[1] != 0) return;
    if (mode == 0 && status[0] != 0) return;
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (warp >= n_rows) return;
    int base = ptr[warp], len = ptr[warp + 1] - base;
    for (int a = lane; a < len; a += 32) {
        int ka = tmp_key[base + a];
        int cnt = 0;
        for (int b = 0; b < len; ++b) cnt += (tmp_key[base + b] < ka) ? 1 : 0;
        out_key[base + cnt] = ka;
        int src = (mode == 0) ? ka : slot_eid_in[ka];
        out_val[base + cnt] = (int)idx_src[src];
    }
}

__global__ void k_fill_csc(const int* __restrict__ slot_j, int64_t n_edges, int n_atoms, const int* __restrict__ rowptr,
                           const int* __restrict__ sptr, int* __restrict__ cursor, const int* __restrict__ status,
                           int* __restrict__ tmp_slot) {
    SPK_PDL_ENTER();
    int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= n_edges) return;
    if (status[1] != 0) return;
    if (s >= rowptr[n_atoms]) return;            // filtered build: only the first rowptr[N] slots exist
    int j = slot_j[s];
    int p = sptr[j] + atomicAdd(&cursor[j], 1);
    tmp_slot[p] = (int)s;
}

// Bad input (status[1] != 0: some idx_i/idx_j outside [0, n_atoms)): the fill/rank kernels above returned early and left the
// slot arrays unwritten, so the row pointers are zeroed -- every consumer (edge, cfconv, pairwise kernels) then sees an
// EMPTY graph instead of gathering through garbage indices.  The host raises IndexError when it builds a new list outside
// stream capture (ops.EdgeGraph); inside a captured graph the caller can poll status[1].
__global__ void k_guard(int* __restrict__ rowptr, int* __restrict__ sptr, int n, const int* __restrict__ status) {
    SPK_PDL_ENTER();
    if (status[1] == 0) return;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k <= n; k += gridDim.x * blockDim.x) {
        rowptr[k] = 0;
        sptr[k] = 0;
    }
}

__global__ void k_segment_ptr(const int64_t* __restrict__ idx_m, int64_t n_atoms, int64_t n_mol,
                              int* __restrict__ mol_ptr) {
    SPK_PDL_ENTER();
    int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (a > n_atoms) return;
    // boundary a: systems in (prev, cur] start at atom a
    int64_t prev = (a == 0) ? -1 : idx_m[a - 1];
    int64_t cur = (a == n_atoms) ? n_mol : idx_m[a];
    if (cur > n_mol) cur = n_mol;
    for (int64_t m = prev + 1; m <= cur && m <= n_mol; ++m) mol_ptr[m] = (int)a;
}

}  // namespace

extern "C" int spk_version(void) { return 1; }

extern "C" size_t spk_graph_workspace_bytes(int64_t n_atoms, int64_t n_edges) {
    // deg_i[N+1] deg_j[N+1] cursor_i[N+1] cursor_j[N+1] tmp_a[E] tmp_b[E]
    size_t n = (size_t)(n_atoms + 1), e = (size_t)(n_edges > 0 ? n_edges : 1);
    return sizeof(int) * (4 * n + 2 * e) + 256;
}

static int graph_build_impl(const int64_t* idx_i, const int64_t* idx_j, const float* r_ij, float cutoff, int64_t n_atoms,
                            int64_t n_edges, int32_t* rowptr, int32_t* slot_j, int32_t* slot_eid, int32_t* sptr,
                            int32_t* pos_slot, int32_t* pos_i, int32_t* status, void* workspace, size_t workspace_bytes,
                            spk_stream_t stream) {
    if (n_atoms < 0 || n_edges < 0 || n_atoms >= (1ll << 31) - 1 || n_edges >= (1ll << 31) - 1) return SPK_ERR_ARG;
    if (!rowptr || !sptr || !status || !workspace) return SPK_ERR_ARG;
    if (workspace_bytes < spk_graph_workspace_bytes(n_atoms, n_edges)) return SPK_ERR_ARG;
    if (n_edges > 0 && (!idx_i || !idx_j || !slot_j || !slot_eid || !pos_slot || !pos_i)) return SPK_ERR_ARG;
    cudaStream_t st = spk_st(stream);
    size_t n = (size_t)(n_atoms + 1), e = (size_t)(n_edges > 0 ? n_edges : 1);
    int* deg_i = reinterpret_cast<int*>(workspace);
    int* deg_j = deg_i + n;
    int* cur_i = deg_j + n;
    int* cur_j = cur_i + n;
    int* tmp_a = cur_j + n;
    int* tmp_b = tmp_a + e;
    (void)tmp_b;
    cudaError_t err = cudaMemsetAsync(deg_i, 0, sizeof(int) * 4 * n, st);
    if (err != cudaSuccess) return SPK_CUDA_ERR(err);
    const int T = 256;
    int n_int = (int)n_atoms;
    spk_launch(k_init_status, 1, 32, 0, st, status, r_ij ? 0 : 1);
    if (n_edges > 0) {
        spk_launch(k_count, (unsigned)spk_cdiv(n_edges, T), T, 0, st, idx_i, idx_j, n_atoms, n_edges, r_ij, cutoff, deg_i,
                   deg_j, status);
    }
    spk_launch(k_scan2, 2, 1024, 0, st, deg_i, rowptr, deg_j, sptr, n_int, status);
    if (n_edges > 0) {
        unsigned ge = (unsigned)spk_cdiv(n_edges, T);
        unsigned gw = (unsigned)spk_cdiv(n_atoms * 32, T);
        spk_launch(k_fill_csr, ge, T, 0, st, idx_i, idx_j, n_edges, r_ij, cutoff, rowptr, cur_i, status, slot_j, slot_eid,
                   tmp_a);
        spk_launch(k_rank_rows, gw, T, 0, st, rowptr, n_int, tmp_a, idx_j, nullptr, 0, status, slot_eid, slot_j);
        spk_launch(k_fill_csc, ge, T, 0, st, slot_j, n_edges, n_int, rowptr, sptr, cur_j, status, tmp_a);
        spk_launch(k_rank_rows, gw, T, 0, st, sptr, n_int, tmp_a, idx_i, slot_eid, 1, status, pos_slot, pos_i);
        spk_launch(k_guard, (unsigned)(n_atoms / 256 + 1 < 64 ? n_atoms / 256 + 1 : 64), 256, 0, st, rowptr, sptr, n_int, status);
    }
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}

extern "C" int spk_graph_build(const int64_t* idx_i, const int64_t* idx_j, int64_t n_atoms, int64_t n_edges,
                               int32_t* rowptr, int32_t* slot_j, int32_t* slot_eid, int32_t* sptr, int32_t* pos_slot,
                               int32_t* pos_i, int32_t* status, void* workspace, size_t workspace_bytes,
                               spk_stream_t stream) {
    return graph_build_impl(idx_i, idx_j, nullptr, 0.f, n_atoms, n_edges, rowptr, slot_j, slot_eid, sptr, pos_slot, pos_i,
                            status, workspace, workspace_bytes, stream);
}

extern "C" int spk_graph_build_active(const int64_t* idx_i, const int64_t* idx_j, const float* r_ij, float cutoff,
                                      int64_t n_atoms, int64_t n_edges, int32_t* rowptr, int32_t* slot_j,
                                      int32_t* slot_eid, int32_t* sptr, int32_t* pos_slot, int32_t* pos_i, int32_t* status,
                                      void* workspace, size_t workspace_bytes, spk_stream_t stream) {
    if (n_edges > 0 && !r_ij) return SPK_ERR_ARG;
    return graph_build_impl(idx_i, idx_j, r_ij, cutoff, n_atoms, n_edges, rowptr, slot_j, slot_eid, sptr, pos_slot, pos_i,
                            status, workspace, workspace_bytes, stream);
}

extern "C" int spk_segment_ptr(const int64_t* idx_m, int64_t n_atoms, int64_t n_mol, int32_t* mol_ptr,
                               spk_stream_t stream) {
    if (n_atoms < 0 || n_mol < 0 || !mol_ptr || (n_atoms > 0 && !idx_m)) return SPK_ERR_ARG;
    const int T = 256;
    spk_launch(k_segment_ptr, (unsigned)spk_cdiv(n_atoms + 1, T), T, 0, spk_st(stream), idx_m, n_atoms, n_mol, mol_ptr);
    SPK_LAUNCH_CHECK();
    return SPK_OK;
}
